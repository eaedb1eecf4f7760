// strict build: compiled with -fmad=false, the reference's exact operation order
// (namespace mfc_strict); bit-comparable with the strict build of the CPU oracle
#define MFC_STRICT 1
#include "kernels_inst.inc"
namespace mfc { const Launchers &launchers_strict() { return mfc_strict::table; } }

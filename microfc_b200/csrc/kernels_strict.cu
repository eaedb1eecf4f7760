// strict build: compiled with -fmad=false, the reference's exact operation order
// (namespace mfc_strict); bit-comparable with oracle/liborc_strict.so
#define MFC_STRICT 1
#include "kernels_inst.inc"
namespace mfc { const Launchers &launchers_strict() { return mfc_strict::table; } }

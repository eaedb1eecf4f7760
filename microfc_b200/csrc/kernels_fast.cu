// fast build: default FMA contraction, one-division WENO weights (namespace mfc_fast)
#define MFC_STRICT 0
#include "kernels_inst.inc"
namespace mfc { const Launchers &launchers_fast() { return mfc_fast::table; } }

// fast build, sweep kernels of num_fluids = 3 (see kernels_inst.inc: MFC_NF3_UNIT)
#define MFC_STRICT 0
#define MFC_NF3_UNIT
#include "kernels_inst.inc"

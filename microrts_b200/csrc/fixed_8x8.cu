// fixed_8x8.cu -- the generic step kernel for 8x8 maps with 64 unit slots (maps/8x8/basesWorkers8x8*.xml); see fixed_generic.inc
#define MRTS_TU_W 8
#define MRTS_TU_H 8
#define MRTS_TU_CAP 64
#define MRTS_TU_NAME 8x8
#include "fixed_generic.inc"

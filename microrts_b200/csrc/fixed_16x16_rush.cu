// fixed_16x16_rush.cu -- the rush-only copy of the generic step kernel for 16x16 maps with 128 unit slots; see fixed_generic.inc
#define MRTS_TU_W 16
#define MRTS_TU_H 16
#define MRTS_TU_CAP 128
#define MRTS_TU_NAME 16x16_rush
#define MRTS_TU_RUSH_ONLY 1
#include "fixed_generic.inc"

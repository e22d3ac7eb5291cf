// fixed_24x24_rush.cu -- the rush-only copy of the generic step kernel for 24x24 maps with 128 unit slots; see fixed_generic.inc
#define MRTS_TU_W 24
#define MRTS_TU_H 24
#define MRTS_TU_CAP 128
#define MRTS_TU_NAME 24x24_rush
#define MRTS_TU_RUSH_ONLY 1
#include "fixed_generic.inc"

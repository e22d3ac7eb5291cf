// fixed_16x16.cu -- the generic step kernel for 16x16 maps with 128 unit slots (maps/16x16/basesWorkers16x16*.xml, the reference's most used size); see fixed_generic.inc
#define MRTS_TU_W 16
#define MRTS_TU_H 16
#define MRTS_TU_CAP 128
#define MRTS_TU_NAME 16x16
#include "fixed_generic.inc"

// fixed_24x24.cu -- the generic step kernel for 24x24 maps with 128 unit slots (BASELINE configs[2]: maps/24x24/basesWorkers24x24*.xml); see fixed_generic.inc
#define MRTS_TU_W 24
#define MRTS_TU_H 24
#define MRTS_TU_CAP 128
#define MRTS_TU_NAME 24x24
#include "fixed_generic.inc"

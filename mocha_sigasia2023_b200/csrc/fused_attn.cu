// placeholder until the fused attention kernel lands
#include "fused.cuh"

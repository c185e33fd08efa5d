// One translation unit per design width: nvcc -DFN_W=<1|2|3|4|6|8|12|16> -c fn_width.cu
// (fitnoise_b200/build.py compiles the widths in parallel and links them with fn_api.cu).
#define FN_WIDTH_ONLY
#include "fn_kernels.cuh"
#include "fn_width_inst.cuh"

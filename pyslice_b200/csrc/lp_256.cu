// line-pass kernels for N = 256 (E = 16 elements per thread; tile width rows 16 / cols 16)
#define PSB_LP_N 256
#define PSB_LP_E 16
#define PSB_LP_WR 16
#define PSB_LP_WC 16
#include "line_pass_inst.cuh"

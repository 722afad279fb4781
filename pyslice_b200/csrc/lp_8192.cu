// line-pass kernels for N = 8192 (E = 16 elements per thread; tile width rows 1 / cols 1)
#define PSB_LP_N 8192
#define PSB_LP_E 16
#define PSB_LP_WR 1
#define PSB_LP_WC 1
#include "line_pass_inst.cuh"

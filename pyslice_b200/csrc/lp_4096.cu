// line-pass kernels for N = 4096 (E = 16 elements per thread; tile width rows 1 / cols 2)
#define PSB_LP_N 4096
#define PSB_LP_E 16
#define PSB_LP_WR 1
#define PSB_LP_WC 2
#include "line_pass_inst.cuh"

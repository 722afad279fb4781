// line-pass kernels for N = 2048 (E = 16 elements per thread; tile width rows 2 / cols 4)
#define PSB_LP_N 2048
#define PSB_LP_E 16
#define PSB_LP_WR 2
#define PSB_LP_WC 4
#include "line_pass_inst.cuh"

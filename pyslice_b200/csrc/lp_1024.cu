// line-pass kernels for N = 1024 (E = 16 elements per thread; tile width rows 4 / cols 8)
#define PSB_LP_N 1024
#define PSB_LP_E 16
#define PSB_LP_WR 4
#define PSB_LP_WC 8
#include "line_pass_inst.cuh"

// line-pass kernels for N = 512 (E = 16 elements per thread; tile width rows 8 / cols 16)
#define PSB_LP_N 512
#define PSB_LP_E 16
#define PSB_LP_WR 8
#define PSB_LP_WC 16
#include "line_pass_inst.cuh"

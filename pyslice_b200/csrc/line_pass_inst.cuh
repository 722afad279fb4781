// Instantiates every pass kind for one power-of-two line size.  Included by lp_<N>.cu with
// PSB_LP_N / PSB_LP_E / PSB_LP_WR / PSB_LP_WC defined (size, elements per thread, tile width for
// row- and column-oriented passes).
#include "line_pass.cuh"
#include "psb_rt.h"

namespace psb {

#define PSB_LP_CONCAT2(a, b) a##b
#define PSB_LP_CONCAT(a, b) PSB_LP_CONCAT2(a, b)
#define PSB_LP_FN PSB_LP_CONCAT(launch_lp_, PSB_LP_N)

template <bool COLS, bool BLUE, int F1, int MID, int F2, int ST, bool MS>
static int lp_go(const PassParams& p, int n_img, cudaStream_t s) {
    constexpr int W = COLS ? PSB_LP_WC : PSB_LP_WR;
    using K = LinePass<PSB_LP_N, PSB_LP_E, W, COLS, BLUE, F1, MID, F2, ST, MS>;
    dim3 grid((p.nlines + W - 1) / W, n_img, 1);
    cudaError_t e = launch<K>(grid, K::kSmem, s, p);
    if (e != cudaSuccess) {
#ifndef PSB_EMU
        return fail(PSB_ERR_CUDA, std::string("line pass launch: ") + cudaGetErrorString(e));
#else
        return PSB_ERR_CUDA;
#endif
    }
    return PSB_OK;
}

template <bool BLUE>
static int lp_kind(int kind, const PassParams& p, int n_img, cudaStream_t s) {
    switch (kind) {
        case PASS_R1:       return lp_go<false, BLUE, F_NONE, M_FULL, F_FWD, S_PLAIN, false>(p, n_img, s);
        case PASS_R:        return lp_go<false, BLUE, F_INV, M_FULL, F_FWD, S_PLAIN, false>(p, n_img, s);
        case PASS_C:        return lp_go<true, BLUE, F_FWD, M_SEP, F_INV, S_PLAIN, false>(p, n_img, s);
        case PASS_CX:       return lp_go<true, BLUE, F_NONE, M_NONE, F_FWD, S_SHIFT, false>(p, n_img, s);
        case PASS_INV_ROWS: return lp_go<false, BLUE, F_NONE, M_NONE, F_INV, S_PLAIN, false>(p, n_img, s);
        case PASS_INV_COLS: return lp_go<true, BLUE, F_NONE, M_NONE, F_INV, S_PLAIN, false>(p, n_img, s);
        case PASS_CP:       return lp_go<true, BLUE, F_NONE, M_SEP, F_INV, S_PLAIN, false>(p, n_img, s);
        case PASS_FWD_ROWS: return lp_go<false, BLUE, F_NONE, M_NONE, F_FWD, S_PLAIN, false>(p, n_img, s);
        case PASS_FWD_COLS: return lp_go<true, BLUE, F_NONE, M_NONE, F_FWD, S_PLAIN, false>(p, n_img, s);
        case PASS_TW:       return lp_go<true, BLUE, F_NONE, M_NONE, F_FWD, S_ABS2, true>(p, n_img, s);
        case PASS_RI2:      return lp_go<false, BLUE, F_NONE, M_NONE, F_INV, S_TRANSMIT2, false>(p, n_img, s);
    }
    return fail(PSB_ERR_INVALID, "unknown pass kind");
}

int PSB_LP_FN(int kind, bool blue, const PassParams& p, int n_img, cudaStream_t s) {
    return blue ? lp_kind<true>(kind, p, n_img, s) : lp_kind<false>(kind, p, n_img, s);
}

}  // namespace psb

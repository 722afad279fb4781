#include "line_pass.cuh"
#include "tables.h"

namespace psb {

#define PSB_DECL(N) int launch_lp_##N(int kind, bool blue, const PassParams& p, int n_img, cudaStream_t s);
PSB_DECL(16) PSB_DECL(32) PSB_DECL(64) PSB_DECL(128) PSB_DECL(256) PSB_DECL(512) PSB_DECL(1024) PSB_DECL(2048) PSB_DECL(4096) PSB_DECL(8192)
#undef PSB_DECL

int launch_line_pass(int kind, PassParams p, int n_img, cudaStream_t stream) {
    if (n_img <= 0 || p.nlines <= 0) return PSB_OK;
    if (n_img > 65535) return fail(PSB_ERR_UNSUPPORTED, "more than 65535 images in one pass");
    int N = 0;
    bool blue = false;
    int rc = get_fft_tables(p.line_len, &p.tb, &N, &blue, stream);
    if (rc != PSB_OK) return rc;
    switch (N) {
        case 16: return launch_lp_16(kind, blue, p, n_img, stream);
        case 32: return launch_lp_32(kind, blue, p, n_img, stream);
        case 64: return launch_lp_64(kind, blue, p, n_img, stream);
        case 128: return launch_lp_128(kind, blue, p, n_img, stream);
        case 256: return launch_lp_256(kind, blue, p, n_img, stream);
        case 512: return launch_lp_512(kind, blue, p, n_img, stream);
        case 1024: return launch_lp_1024(kind, blue, p, n_img, stream);
        case 2048: return launch_lp_2048(kind, blue, p, n_img, stream);
        case 4096: return launch_lp_4096(kind, blue, p, n_img, stream);
        case 8192: return launch_lp_8192(kind, blue, p, n_img, stream);
    }
    return fail(PSB_ERR_UNSUPPORTED, "unsupported FFT size");
}

}  // namespace psb

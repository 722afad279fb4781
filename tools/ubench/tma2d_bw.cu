// Microbenchmark: throughput and completion latency of the column pass's input pattern on sm_100a -- a 2-D tiled TMA copy
// (cp.async.bulk.tensor.2d) of a box of W adjacent 8-byte columns x 256 rows out of an L2-resident (rows, 256) array --
// against a 1-D bulk copy of the same number of bytes.  Question (ncu of fast_cols_kernel: long-scoreboard = mbarrier waits
// are its top stall): how many bytes per clock does one SM get through the TMA unit when every box row is only 64 / 128 /
// 256 bytes long, and how long does one box take from issue to completion when nothing else is in flight?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma2d_bw tools/ubench/tma2d_bw.cu
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
            smem_u32(b)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tensor2d_g2s(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// one issuing thread per CTA, STAGES boxes in flight; mode 0: 2-D boxes of W columns x 256 rows, mode 1: 1-D bulk of the same bytes
template <int W, int STAGES>
__global__ void __launch_bounds__(128) k_tma(const __grid_constant__ CUtensorMap map, const char* base, int n_img, int ncols, int mode, int reps,
                                             long long* lat_out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t full[STAGES];
    constexpr int BOX = W * 256 * 8;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int tiles_per_img = ncols / W;
    const long long n_tiles = (long long)n_img * tiles_per_img;
    long long issued = 0, done = 0, lat = 0;
    for (int r = 0; r < reps; ++r)
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int s = (int)(issued % STAGES);
            if (issued >= STAGES) {
                mbar_wait(&full[s], (uint32_t)(((issued / STAGES) - 1) & 1));
                ++done;
            }
            mbar_expect_tx(&full[s], BOX);
            const long long c0 = clock64();
            if (mode == 0) tensor2d_g2s(sm + (size_t)s * BOX, &map, (int)(t % tiles_per_img) * W, (int)(t / tiles_per_img) * 256, &full[s]);
            else bulk_g2s(sm + (size_t)s * BOX, base + t * BOX, BOX, &full[s]);
            if (STAGES == 1) {          // latency mode: wait for this very box
                mbar_wait(&full[s], (uint32_t)(issued & 1));
                lat += clock64() - c0;
                ++done;
            }
            ++issued;
        }
    for (; done < issued; ++done) mbar_wait(&full[done % STAGES], (uint32_t)((done / STAGES) & 1));
    if (lat_out && blockIdx.x == 0) lat_out[0] = issued ? lat / issued : 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int W, int STAGES>
void run(EncodeTiledFn enc, char* buf, int n_img, int ncols, int ctas_per_sm, int sms, double clk_hz) {
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)ncols, (cuuint64_t)n_img * 256};
    const cuuint64_t gstride[1] = {(cuuint64_t)ncols * 8};
    const cuuint32_t box[2] = {(cuuint32_t)W, 256};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, buf, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        printf("encode failed\n");
        return;
    }
    constexpr int BOX = W * 256 * 8;
    auto kern = k_tma<W, STAGES>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, BOX * STAGES);
    long long* lat;
    cudaMalloc(&lat, 8);
    const int reps = 20;
    for (int mode = 0; mode < 2; ++mode) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        kern<<<sms * ctas_per_sm, 128, BOX * STAGES>>>(map, buf, n_img, ncols, mode, 2, lat);
        cudaEventRecord(a);
        kern<<<sms * ctas_per_sm, 128, BOX * STAGES>>>(map, buf, n_img, ncols, mode, reps, lat);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        long long hl = 0;
        cudaMemcpy(&hl, lat, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)n_img * 256 * ncols * 8 * reps;
        printf("box %2d cols x 256 rows (%3d-byte rows, %2d KB)  %s  %d CTA/SM x %d in flight: %8.1f GB/s = %5.1f B/clk/SM", W, W * 8, BOX / 1024,
               mode == 0 ? "2-D tensor" : "1-D bulk  ", ctas_per_sm, STAGES, bytes / ms * 1e-6, bytes / (ms * 1e-3) / sms / clk_hz);
        if (STAGES == 1) printf("   issue-to-complete %lld cycles", hl);
        printf("\n");
    }
    cudaFree(lat);
}

int main() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(ptr);
    const int n_img = 127, ncols = 256;          // 127 images of 256 x 256 complex64 = 66.6 MB: the slice step's L2-resident batch
    char* buf;
    cudaMalloc(&buf, (size_t)n_img * 256 * ncols * 8);
    cudaMemset(buf, 1, (size_t)n_img * 256 * ncols * 8);
    const double hz = clk * 1e3;
    printf("SMs %d, %.0f MHz; array: %d images of 256 x %d complex64 (L2-resident)\n", sms, clk / 1e3, n_img, ncols);
    run<8, 1>(enc, buf, n_img, ncols, 1, sms, hz);
    run<16, 1>(enc, buf, n_img, ncols, 1, sms, hz);
    run<32, 1>(enc, buf, n_img, ncols, 1, sms, hz);
    run<8, 2>(enc, buf, n_img, ncols, 2, sms, hz);
    run<16, 2>(enc, buf, n_img, ncols, 2, sms, hz);
    run<32, 2>(enc, buf, n_img, ncols, 2, sms, hz);
    run<16, 3>(enc, buf, n_img, ncols, 2, sms, hz);
    run<16, 1>(enc, buf, n_img, ncols, 2, sms, hz);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}

// Microbenchmark: L2 -> SM bandwidth on sm_100a for a working set that stays L2-resident.
//   mode 0: LDG.128 streaming read (grid-stride), mode 1: read + write copy (LDG.128 / STG.128),
//   mode 2: cp.async.bulk (TMA 1-D) global -> shared ring, one issuing thread per CTA, no compute,
//   mode 3: mode 2 plus a cp.async.bulk shared -> global store of every chunk (copy through smem).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/build/l2bw tools/ubench/l2bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_read(const float4* __restrict__ src, size_t n4, int reps, float* out) {
    float acc = 0.f;
    for (int r = 0; r < reps; ++r) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            float4 v = __ldcg(&src[i]);
            acc += v.x + v.y + v.z + v.w;
        }
    }
    if (acc == 1.2345f) out[0] = acc;
}

__global__ void __launch_bounds__(256) k_copy(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4, int reps) {
    for (int r = 0; r < reps; ++r) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
            dst[i] = __ldcg(&src[i]);
    }
}

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
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

template <int CHUNK, int STAGES, bool STORE>
__global__ void __launch_bounds__(128) k_bulk(const char* __restrict__ src, char* __restrict__ dst, size_t bytes, int reps) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t full[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t nchunks = bytes / CHUNK;
    long long issued = 0, done = 0;
    for (int r = 0; r < reps; ++r) {
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            const int s = (int)(issued % STAGES);
            if (issued >= STAGES) {   // wait for the chunk that used this stage
                mbar_wait(&full[s], (uint32_t)(((issued / STAGES) - 1) & 1));
                if (STORE) {
                    // chunk `done` landed in stage s: store it, then the stage may be reused once the store has read smem
                    bulk_s2g(dst + c * CHUNK, sm + (size_t)s * CHUNK, CHUNK);   // bandwidth only: contents do not matter
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                ++done;
            }
            mbar_expect_tx(&full[s], CHUNK);
            bulk_g2s(sm + (size_t)s * CHUNK, src + c * CHUNK, CHUNK, &full[s]);
            ++issued;
        }
    }
    for (; done < issued; ++done) {
        const int s = (int)(done % STAGES);
        mbar_wait(&full[s], (uint32_t)((done / STAGES) & 1));
    }
    if (STORE) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static float timeit(void (*fn)(void*), void* ctx) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fn(ctx);
    cudaEventRecord(e0);
    fn(ctx);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

struct Ctx { char* a; char* b; size_t bytes; int reps; int grid; float* out; };

int main() {
    const int reps = 40;
    float* out; cudaMalloc(&out, 4);
    for (size_t mb : {16, 32, 48, 96, 512}) {
        Ctx c; c.bytes = mb << 20; c.reps = mb > 100 ? 4 : reps; c.out = out;
        cudaMalloc(&c.a, c.bytes); cudaMalloc(&c.b, c.bytes);
        cudaMemset(c.a, 1, c.bytes); cudaMemset(c.b, 0, c.bytes);
        for (int occ : {2, 4, 8}) {
            c.grid = 148 * occ;
            float ms = timeit([](void* p) { Ctx* c = (Ctx*)p; k_read<<<c->grid, 256>>>((const float4*)c->a, c->bytes / 16, c->reps, c->out); }, &c);
            printf("%4zu MB  LDG.128 read      grid=148x%d  %8.1f GB/s\n", mb, occ, (double)c.bytes * c.reps / ms * 1e-6);
            ms = timeit([](void* p) { Ctx* c = (Ctx*)p; k_copy<<<c->grid, 256>>>((const float4*)c->a, (float4*)c->b, c->bytes / 16, c->reps); }, &c);
            printf("%4zu MB  LDG/STG copy      grid=148x%d  %8.1f GB/s (read+write)\n", mb, occ, 2.0 * c.bytes * c.reps / ms * 1e-6);
        }
        {
            constexpr int CH = 32768, ST = 6;
            cudaFuncSetAttribute(k_bulk<CH, ST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * ST);
            cudaFuncSetAttribute(k_bulk<CH, ST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * ST);
            c.grid = 148;
            float ms = timeit([](void* p) { Ctx* c = (Ctx*)p; k_bulk<CH, ST, false><<<c->grid, 128, CH * ST>>>(c->a, c->b, c->bytes, c->reps); }, &c);
            printf("%4zu MB  bulk g2s 32K x6   grid=148    %8.1f GB/s\n", mb, (double)c.bytes * c.reps / ms * 1e-6);
            ms = timeit([](void* p) { Ctx* c = (Ctx*)p; k_bulk<CH, ST, true><<<c->grid, 128, CH * ST>>>(c->a, c->b, c->bytes, c->reps); }, &c);
            printf("%4zu MB  bulk g2s+s2g      grid=148    %8.1f GB/s (read+write)\n", mb, 2.0 * c.bytes * c.reps / ms * 1e-6);
        }
        {
            constexpr int CH = 8192, ST = 12;
            cudaFuncSetAttribute(k_bulk<CH, ST, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH * ST);
            c.grid = 296;
            float ms = timeit([](void* p) { Ctx* c = (Ctx*)p; k_bulk<CH, ST, false><<<c->grid, 128, CH * ST>>>(c->a, c->b, c->bytes, c->reps); }, &c);
            printf("%4zu MB  bulk g2s 8K x12   grid=296    %8.1f GB/s\n", mb, (double)c.bytes * c.reps / ms * 1e-6);
        }
        cudaFree(c.a); cudaFree(c.b);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}

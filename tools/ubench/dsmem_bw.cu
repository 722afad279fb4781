// Microbenchmark: distributed-shared-memory bandwidth inside a thread-block cluster on sm_100a, against local shared
// memory with the same access loop.  Decides whether psi can live in the shared memory of a cluster across slices
// (row pass local, column pass = transpose through DSMEM) instead of making two L2 round trips per slice step.
//   mode 0: every CTA reads its OWN buffer (ld.shared, 16 B per thread per access)             -> local baseline
//   mode 1: every CTA reads the buffer of the next CTA of its cluster (ld.shared::cluster)     -> DSMEM read
//   mode 2: every CTA writes the buffer of the next CTA of its cluster (st.shared::cluster)    -> DSMEM write
//   mode 3: reads from ALL other CTAs of the cluster in turn (the transpose pattern)           -> DSMEM all-to-all read
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dsmem_bw tools/ubench/dsmem_bw.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

constexpr int kBufBytes = 64 * 1024;
constexpr int kThreads = 512;

__global__ void __launch_bounds__(kThreads, 1) k_dsmem(int mode, int reps, float* out) {
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank(), cs = cluster.num_blocks();
    float4* mine = reinterpret_cast<float4*>(smem);
    for (int i = threadIdx.x; i < kBufBytes / 16; i += kThreads) mine[i] = make_float4(rank, i, 1.f, 2.f);
    cluster.sync();
    float acc = 0.f;
    constexpr int n4 = kBufBytes / 16;
    for (int r = 0; r < reps; ++r) {
        unsigned peer = mode == 0 ? rank : (mode == 3 ? (rank + 1 + r % (cs > 1 ? cs - 1 : 1)) % cs : (rank + 1) % cs);
        float4* p = cluster.map_shared_rank(mine, peer);
        const int o = (r * 97) & (n4 - 1);          // the offset changes every repetition: nothing can be hoisted out of the loop
        if (mode == 2) {
#pragma unroll 8
            for (int i = threadIdx.x; i < n4; i += kThreads) p[(i + o) & (n4 - 1)] = make_float4(acc, r, i, 0.f);
        } else {
#pragma unroll 8
            for (int i = threadIdx.x; i < n4; i += kThreads) {
                const float4 v = p[(i + o) & (n4 - 1)];
                acc += v.x + v.y + v.z + v.w;
            }
        }
    }
    cluster.sync();
    if (acc == 1.2345f) out[0] = acc;
}

int main() {
    int sms = 0, clk = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out;
    cudaMalloc(&out, 4);
    cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeMaxDynamicSharedMemorySize, kBufBytes);
    cudaFuncSetAttribute(k_dsmem, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const char* names[4] = {"local read", "dsmem read (neighbour)", "dsmem write (neighbour)", "dsmem read (all peers in turn)"};
    printf("SMs %d, max SM clock %.0f MHz, buffer %d KB per CTA, %d threads, 16 B per thread and access\n", sms, clk / 1e3, kBufBytes / 1024,
           kThreads);
    for (int cs : {1, 2, 4, 8, 16}) {
        for (int mode = 0; mode < 4; ++mode) {
            if (cs == 1 && mode != 0) continue;
            cudaLaunchConfig_t cfg = {};
            int max_clusters = 0;
            cfg.gridDim = dim3(cs);
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = kBufBytes;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            if (cudaOccupancyMaxActiveClusters(&max_clusters, k_dsmem, &cfg) != cudaSuccess || max_clusters < 1) {
                printf("cluster %2d: not launchable (%s)\n", cs, cudaGetErrorString(cudaGetLastError()));
                break;
            }
            const int ctas = max_clusters * cs;          // one CTA per SM (launch bounds + 64 KB + 512 threads do not force it: occupancy API decides)
            cfg.gridDim = dim3(ctas < sms ? ctas : (sms / cs) * cs);
            const int reps = 2000;
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            cudaLaunchKernelEx(&cfg, k_dsmem, mode, 10, out);
            cudaDeviceSynchronize();
            cudaEventRecord(a);
            cudaError_t e = cudaLaunchKernelEx(&cfg, k_dsmem, mode, reps, out);
            cudaEventRecord(b);
            cudaDeviceSynchronize();
            if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("cluster %2d mode %d: launch failed\n", cs, mode); continue; }
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            const double bytes = (double)cfg.gridDim.x * reps * kBufBytes;
            printf("cluster %2d  %-32s grid %3d CTAs (max active clusters %3d): %8.1f GB/s total, %6.1f GB/s per SM, %5.1f B/clk/SM at max clock\n",
                   cs, names[mode], cfg.gridDim.x, max_clusters, bytes / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / 1e9 / cfg.gridDim.x,
                   bytes / (ms * 1e-3) / cfg.gridDim.x / (clk * 1e3));
        }
    }
    return 0;
}

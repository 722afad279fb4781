// CUDA-graph replay of the library's long launch sequences (one multislice batch = ~1000 back-to-back kernels).
//
// psb_propagate_ex and psb_build_transmission / psb_build_phase issue the same sequence of launches whenever they are
// called with the same arguments (same buffers, same sizes) -- which is what a run does batch after batch and step after
// step.  run_graphed() keys a sequence by the bytes of its argument block: the first call with a key runs eagerly (lazy
// table uploads and workspace growth happen there), the second captures the sequence on an internal stream into a graph,
// and from then on the call is a single cudaGraphLaunch into the caller's stream.  The GPU then schedules the whole batch
// itself: the host thread no longer has to win the driver lock a thousand times per batch, so pollers of that lock
// (nvidia-smi, NVML clock samplers of this or other processes) can no longer starve the GPU between launches, and the
// host cost of a batch drops from ~3 ms to ~30 us.  Capture failures fall back to eager launches for that key.
#pragma once
#include "psb_common.cuh"

#include <cstddef>
#include <functional>

namespace psb {

void graph_mode_set(int on);      // 0: always eager; 1 (default, PSB_GRAPHS=0 disables): replay
int graph_mode();
void graph_cache_release();       // destroys every cached graph (tables they point at are about to be freed)

// `eager(stream)` issues the launches on `stream` and returns a psb status.  `s` is the caller's stream.
int run_graphed(const void* key, size_t key_bytes, cudaStream_t s, const std::function<int(cudaStream_t)>& eager);

}  // namespace psb

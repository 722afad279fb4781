#include "graph_cache.h"

#include "psb_rt.h"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace psb {

namespace {

struct Entry {
    std::vector<unsigned char> key;
    cudaGraphExec_t exec = nullptr;
    long long kernels = 0;          // launches the graph stands for (psb_launch_count bookkeeping)
    bool failed = false;
};

constexpr size_t kMaxEntries = 96;
std::mutex g_mu;
std::list<Entry> g_entries;                         // most recently used first
std::map<int, cudaStream_t> g_capture_streams;      // per device
std::atomic<int> g_mode{-1};

cudaStream_t capture_stream() {                     // g_mu held
    const int dev = rt::device();
    auto it = g_capture_streams.find(dev);
    if (it != g_capture_streams.end()) return it->second;
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        s = nullptr;
    }
    g_capture_streams[dev] = s;
    return s;
}

}  // namespace

void graph_mode_set(int on) { g_mode.store(on ? 1 : 0); }

int graph_mode() {
    int m = g_mode.load();
    if (m < 0) {
        const char* e = std::getenv("PSB_GRAPHS");
        m = (e && e[0] == '0') ? 0 : 1;
        g_mode.store(m);
    }
    return m;
}

void graph_cache_release() {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& e : g_entries)
        if (e.exec) cudaGraphExecDestroy(e.exec);
    g_entries.clear();
}

int run_graphed(const void* key, size_t key_bytes, cudaStream_t s, const std::function<int(cudaStream_t)>& eager) {
    if (!graph_mode()) return eager(s);
    // inside somebody else's capture (a caller wrapping the library in its own graph): just contribute the launches
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess) {
        cudaGetLastError();
        return eager(s);
    }
    if (st != cudaStreamCaptureStatusNone) return eager(s);

    std::vector<unsigned char> k(sizeof(int) + key_bytes);
    const int dev = rt::device();
    std::memcpy(k.data(), &dev, sizeof(int));
    std::memcpy(k.data() + sizeof(int), key, key_bytes);

    std::unique_lock<std::mutex> lk(g_mu);
    auto it = g_entries.begin();
    for (; it != g_entries.end(); ++it)
        if (it->key == k) break;
    if (it == g_entries.end()) {                    // first sighting: eager, so that lazy set-up runs outside a capture
        g_entries.emplace_front();
        g_entries.front().key = std::move(k);
        if (g_entries.size() > kMaxEntries) {
            if (g_entries.back().exec) cudaGraphExecDestroy(g_entries.back().exec);
            g_entries.pop_back();
        }
        lk.unlock();
        return eager(s);
    }
    g_entries.splice(g_entries.begin(), g_entries, it);          // most recently used
    Entry& e = g_entries.front();
    if (e.failed) {
        lk.unlock();
        return eager(s);
    }
    if (!e.exec) {                                  // second sighting: capture on the internal stream (nothing executes)
        cudaStream_t cs = capture_stream();
        if (!cs || cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            e.failed = true;
            lk.unlock();
            return eager(s);
        }
        const long long before = launch_counter().load();
        const int rc = eager(cs);
        cudaGraph_t g = nullptr;
        cudaError_t ce = cudaStreamEndCapture(cs, &g);
        const long long kernels = launch_counter().load() - before;
        launch_counter().fetch_sub(kernels);        // nothing ran yet
        cudaGraphExec_t exec = nullptr;
        if (rc == PSB_OK && ce == cudaSuccess && g) ce = cudaGraphInstantiate(&exec, g, 0);
        if (g) cudaGraphDestroy(g);
        if (rc != PSB_OK || ce != cudaSuccess || !exec) {
            cudaGetLastError();
            e.failed = true;
            lk.unlock();
            return eager(s);
        }
        e.exec = exec;
        e.kernels = kernels;
    }
    cudaError_t le = cudaGraphLaunch(e.exec, s);
    const long long kernels = e.kernels;
    lk.unlock();
    if (le != cudaSuccess) return fail(PSB_ERR_CUDA, std::string("graph launch: ") + cudaGetErrorString(le));
    launch_counter().fetch_add(kernels);
    return PSB_OK;
}

}  // namespace psb

// twl_ctx.hpp — host-side context shared by the C-ABI translation units (twl_api.cu: batch DP; twl_level.cu: the
// device-resident level pipeline).
#pragma once
#include "../../include/twilight_b200.h"
#include "twl_device.cuh"

#include <algorithm>
#include <string>
#include <vector>

template <typename T>
struct DevBuf {
    T *ptr = nullptr;
    size_t cap = 0; // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = n + n / 4 + 256;
        cudaError_t e = cudaMalloc(&ptr, want * sizeof(T));
        if (e != cudaSuccess) { e = cudaMalloc(&ptr, n * sizeof(T)); want = n; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};

template <typename T>
struct PinBuf {
    T *ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        const size_t had = cap;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
        // pinning costs about a millisecond per MB: grow by doubling so that a buffer that keeps growing over the levels of a big
        // tree (cached msaFreq staging) is re-pinned O(log) times
        size_t want = std::max(n + n / 4 + 256, 2 * had);
        cudaError_t e = cudaMallocHost(&ptr, want * sizeof(T));
        if (e != cudaSuccess && want > n) { cudaGetLastError(); want = n; e = cudaMallocHost(&ptr, want * sizeof(T)); }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (ptr) cudaFreeHost(ptr); ptr = nullptr; cap = 0; }
};

struct TwlLevelState;   // twl_level.cu

struct twl_ctx {
    int device = 0;
    int smCount = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = true;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    std::string error;

    // scoring
    int M = 0, P = 0;
    float gapOpen = 0, gapExtend = 0, gapBoundary = 0;
    int marker = twl::kMaxMarker;
    DevBuf<float> dScore;
    std::vector<float> hScore;

    // staged batch
    int nPairs = 0;
    bool staged = false, ran = false;
    size_t profWords = 0, pathBytes = 0;
    std::vector<twl::DevPair> hPairs;
    std::vector<int> hOrder;
    int maxFLen = 0;
    PinBuf<float> hProf;
    PinBuf<int8_t> hPaths;
    PinBuf<twl::DevResult> hResults;
    DevBuf<float> dProf;
    DevBuf<twl::DevPair> dPairs;
    DevBuf<twl::DevResult> dResults;
    DevBuf<int8_t> dPaths;
    DevBuf<int> dOrder, dOverflow;
    DevBuf<int> dCounters; // per chain stage: [2s] queue cursor, [2s+1] work count
    DevBuf<uint8_t> dTb;
    DevBuf<float> dState;

    int latencyMode = -1;        // low-latency wavefront shape (one CTA per SM) for levels with <= smCount pairs: -1 auto, 0 off, 1 always
    int dpTrace = 0;
    int injectNomem = 0;         // fault injection for tests: the n-th twl_align_level call from now fails with TWL_E_NOMEM after its last
                                 // chunk has already flipped row buffers (exercises the rollback and the caller's spill-and-retry path)
    int maxCtasPerSm = 0;        // cap on resident CTAs per SM of the wavefront stages (0 = as many as fit); occupancy experiments
    int wideWorkers = 8;         // CTAs of the wide wavefront kernel that run next to the narrow one (0 = run the wide stage afterwards)
    PinBuf<int> hWatchdog;       // copy of the device flag a wide worker sets when its 20 ms watchdog fires: the two kernels were not
                                 // co-scheduled (MPS / time slicing, profiler, CUDA_LAUNCH_BLOCKING) -> co-running is switched off for
                                 // the rest of the context's life instead of paying the watchdog at every level
    cudaStream_t stream2 = nullptr;
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    int latencyShape = 2;        // 2: 512 threads x 2 rows; 3: 512 x 1 first, 512 x 2 for pairs whose band outgrows 512 rows
    bool forceGeneric = false;   // route nucleotide batches through the generic kernel (A/B parity + benchmarking)
    // Protein path (talco_sim.cu + wavefront kernel with SC = 1): the callers of twlLaunchDpChain leave, per pair index, upper
    // bounds of the two profile lengths and the host copy of the work list; the chain lays the similarity matrices out from those.
    int proteinSim = 1;          // 0: proteins run on the generic kernel only (A/B)
    size_t simBudgetBytes = static_cast<size_t>(16) << 30;   // larger batches fall back to the generic kernel
    std::vector<int> hSimRefUb, hSimQryUb, hChainOrder;
    std::vector<twl::DevSim> hSimInfo;
    DevBuf<float> dSim;
    DevBuf<twl::DevSim> dSimInfo;
    float lastMs = -1.0f;
    int lastLaunches = 0;
    bool timingPending = false;

    TwlLevelState *level = nullptr;
};

constexpr size_t kProfPadWords = 4096;   // 16 KB of slack on both ends of the device profile buffer (wavefront kernel over-reads)

int twlFail(twl_ctx *ctx, int code, const std::string &msg);
// Enqueues the DP kernel chain over ctx->dPairs / dProf / dOrder (n pairs, widest legal band wideCap); no timing events.
int twlLaunchDpChain(twl_ctx *ctx, int n, int wideCap);
void twlLevelDestroy(twl_ctx *ctx);

#define TWL_CUDA(ctx, call)                                                                             \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            return twlFail(ctx, (e_ == cudaErrorMemoryAllocation) ? TWL_E_NOMEM : TWL_E_CUDA,           \
                           std::string(#call) + ": " + cudaGetErrorString(e_));                         \
        }                                                                                               \
    } while (0)

// twl_api.cu — host side of the C ABI declared in include/twilight_b200.h: context lifecycle, batch staging in
// pinned memory, kernel launches on one stream, result download. No CPU alignment code lives here: if CUDA is not
// usable every entry point fails.
#include "twl_ctx.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

namespace twl {
size_t genericStateWords(int stateCap);
cudaError_t launchTalcoGeneric(int P, bool globalState, const TalcoArgs &args, int grid, size_t dynSmemBytes, cudaStream_t stream);
int genericThreads();
int wavefrontBandCapacity(int threads, int slots);
int wavefrontWindow(int threads, int slots);
int nucleotideMatrixClass(const float *score5x5);
cudaError_t launchTalcoWavefront(int threads, int slots, int matClass, const TalcoArgs &args, int grid, cudaStream_t stream);
cudaError_t launchCoRunGate(const int *arrived, int want, cudaStream_t stream);
int wavefrontMaxCtasPerSm(int threads, int slots, int matClass);
int simMatrixTiles(int refLen, int qryLen);
cudaError_t launchSimMatrixAa(const float *prof, const DevPair *pairs, const int *order, int nOrder, const DevSim *simInfo, float *sim, const float *score,
                              int maxTiles, cudaStream_t stream);
cudaError_t launchDivSelfTest(const float *num, const float *den, int n, int *mismatches, cudaStream_t stream);
} // namespace twl

namespace {

std::string g_initError;

int fail(twl_ctx *ctx, int code, const std::string &msg) { return twlFail(ctx, code, msg); }

constexpr int kSmemStateCap = 1020;   // band cells held in shared memory by the narrow generic variant
constexpr int kNarrowCtasPerSm = 3;

size_t tbBytesPerCta(int marker) {
    // diagonal k <= marker stores at most k+1 traceback bytes
    size_t n = static_cast<size_t>(marker + 1) * (marker + 2) / 2;
    return (n + 255) & ~static_cast<size_t>(255);
}

// words one side of `len` columns occupies in the packed profile buffer
size_t sideWords(int len, int P) {
    if (P == 6) return static_cast<size_t>((len + 3) / 4) * 32;   // 8 streams of ceil(len/4) float4
    return static_cast<size_t>(len) * (P + 2);
}

// returns true when every column of a nucleotide side is exactly one-hot (one of A,C,G,T,N = 1.0, everything else 0)
bool packColumns(float *dst, const float *freq, const float *gapOp, const float *gapEx, int len, int P) {
    const int PW = P + 2;
    if (P == 6) {   // de-interleaved nucleotide layout, see twl_device.cuh
        const int n4 = (len + 3) / 4;
        std::memset(dst, 0, sideWords(len, P) * sizeof(float));
        bool oneHot = true;
        for (int c = 0; c < len; ++c) {
            float *x = dst + twl::ntColIndex(c, n4) * 4;
            float *y = x + static_cast<size_t>(16) * n4;
            const float *f = freq + static_cast<size_t>(c) * 6;
            x[0] = f[0]; x[1] = f[1]; x[2] = f[2]; x[3] = f[3];
            y[0] = f[4]; y[1] = f[5]; y[2] = gapOp[c]; y[3] = gapEx[c];
            int ones = 0, zeros = 0;
            for (int v = 0; v < 5; ++v) { ones += (f[v] == 1.0f); zeros += (f[v] == 0.0f); }
            oneHot = oneHot && ones == 1 && zeros == 4 && f[5] == 0.0f;
        }
        return oneHot;
    }
    for (int c = 0; c < len; ++c) {
        float *d = dst + static_cast<size_t>(c) * PW;
        std::memcpy(d, freq + static_cast<size_t>(c) * P, sizeof(float) * P);
        d[P] = gapOp[c];
        d[P + 1] = gapEx[c];
    }
    return false;
}

} // namespace

int twlFail(twl_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->error = msg;
    else g_initError = msg;
    return code;
}

extern "C" {

const char *twl_version(void) { return "twilight_b200 0.1 (sm_100a)"; }

int twl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int twl_init(int device, twl_ctx **out) {
    if (!out) return fail(nullptr, TWL_E_ARG, "twl_init: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, TWL_E_NO_DEVICE, std::string("twl_init: no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU fallback");
    if (device < 0 || device >= n) return fail(nullptr, TWL_E_ARG, "twl_init: device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, TWL_E_CUDA, "twl_init: cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, TWL_E_NO_DEVICE, "twl_init: device is not sm_100 (kernels are built for sm_100a only)");
    twl_ctx *ctx = new twl_ctx();
    ctx->device = device;
    ctx->smCount = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->evStart) != cudaSuccess || cudaEventCreate(&ctx->evStop) != cudaSuccess) {
        delete ctx;
        return fail(nullptr, TWL_E_CUDA, "twl_init: stream/event creation failed");
    }
    // tuning switches for A/B runs of unmodified host programs: TWL_OPTIONS="name=value,name=value" (see twl_set_option)
    if (const char *env = std::getenv("TWL_OPTIONS")) {
        std::string all(env);
        size_t at = 0;
        while (at < all.size()) {
            const size_t end = std::min(all.find(',', at), all.size());
            const std::string item = all.substr(at, end - at);
            const size_t eq = item.find('=');
            if (eq != std::string::npos) twl_set_option(ctx, item.substr(0, eq).c_str(), std::atoi(item.c_str() + eq + 1));
            at = end + 1;
        }
    }
    *out = ctx;
    return TWL_OK;
}

void twl_destroy(twl_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->dScore.release(); ctx->dProf.release(); ctx->dPairs.release(); ctx->dResults.release(); ctx->dPaths.release();
    ctx->dOrder.release(); ctx->dOverflow.release(); ctx->dCounters.release(); ctx->dTb.release(); ctx->dState.release(); ctx->dSim.release(); ctx->dSimInfo.release();
    ctx->hProf.release(); ctx->hPaths.release(); ctx->hResults.release(); ctx->hWatchdog.release();
    if (ctx->evStart) cudaEventDestroy(ctx->evStart);
    if (ctx->evStop) cudaEventDestroy(ctx->evStop);
    if (ctx->evFork) cudaEventDestroy(ctx->evFork);
    if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    if (ctx->ownStream && ctx->stream) cudaStreamDestroy(ctx->stream);
    twlLevelDestroy(ctx);
    delete ctx;
}

const char *twl_last_error(const twl_ctx *ctx) { return ctx ? ctx->error.c_str() : g_initError.c_str(); }

int twl_set_params(twl_ctx *ctx, const float *score, int M, float gap_open, float gap_extend, float gap_boundary) {
    if (!ctx) return TWL_E_ARG;
    if (!score || (M != 5 && M != 21)) return fail(ctx, TWL_E_ARG, "twl_set_params: M must be 5 (nucleotide) or 21 (protein)");
    cudaSetDevice(ctx->device);
    TWL_CUDA(ctx, ctx->dScore.reserve(static_cast<size_t>(M) * M));
    TWL_CUDA(ctx, cudaMemcpyAsync(ctx->dScore.ptr, score, sizeof(float) * M * M, cudaMemcpyHostToDevice, ctx->stream));
    TWL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->hScore.assign(score, score + static_cast<size_t>(M) * M);
    ctx->M = M;
    ctx->P = M + 1;
    ctx->gapOpen = gap_open;
    ctx->gapExtend = gap_extend;
    ctx->gapBoundary = gap_boundary;
    ctx->staged = false;
    return TWL_OK;
}

int twl_set_marker(twl_ctx *ctx, int marker) {
    if (!ctx) return TWL_E_ARG;
    if (marker < 1 || marker > twl::kMaxMarker) return fail(ctx, TWL_E_ARG, "twl_set_marker: marker must be in [1, 1024]");
    ctx->marker = marker;
    return TWL_OK;
}

int twl_batch_stage(twl_ctx *ctx, const twl_profile_pair *pairs, int n_pairs) {
    if (!ctx) return TWL_E_ARG;
    if (ctx->P == 0) return fail(ctx, TWL_E_STATE, "twl_batch_stage: call twl_set_params first");
    if (n_pairs < 0 || (n_pairs > 0 && !pairs)) return fail(ctx, TWL_E_ARG, "twl_batch_stage: bad pair list");
    cudaSetDevice(ctx->device);
    ctx->staged = false;
    ctx->ran = false;
    ctx->nPairs = n_pairs;
    if (n_pairs == 0) { ctx->staged = true; return TWL_OK; }
    const int P = ctx->P;
    const int defXdrop = static_cast<int>(1000 * -1 * ctx->gapExtend);     // TALCO-XDrop.cpp:49
    ctx->hPairs.resize(n_pairs);
    size_t words = 0, bytes = 0;
    int maxF = 0;
    for (int p = 0; p < n_pairs; ++p) {
        const twl_profile_pair &in = pairs[p];
        if (in.ref_len < 1 || in.qry_len < 1 || !in.freq_ref || !in.freq_qry || !in.gap_open_ref || !in.gap_ext_ref ||
            !in.gap_open_qry || !in.gap_ext_qry)
            return fail(ctx, TWL_E_ARG, "twl_batch_stage: pair " + std::to_string(p) + " has empty profiles or null pointers");
        twl::DevPair &d = ctx->hPairs[p];
        d.refOff = static_cast<long long>(words); words += sideWords(in.ref_len, P);
        d.qryOff = static_cast<long long>(words); words += sideWords(in.qry_len, P);
        d.refN4 = (in.ref_len + 3) / 4; d.qryN4 = (in.qry_len + 3) / 4;
        d.alnOff = static_cast<long long>(bytes); bytes += (static_cast<size_t>(in.ref_len) + in.qry_len + 15) & ~static_cast<size_t>(15);
        d.refLen = in.ref_len; d.qryLen = in.qry_len;
        d.refNum = in.ref_num; d.qryNum = in.qry_num;
        d.gapChar = in.gap_char_score;
        d.xdrop = in.xdrop > 0 ? in.xdrop : defXdrop;
        d.fLen = in.flen > 0 ? in.flen : 4096;                             // TALCO-XDrop.cpp:50
        d.pad = 0;
        maxF = std::max(maxF, std::min(d.fLen, std::min(d.refLen, d.qryLen)));
    }
    ctx->profWords = words;
    ctx->pathBytes = bytes;
    ctx->maxFLen = maxF;
    TWL_CUDA(ctx, ctx->hProf.reserve(words));
    // pack into pinned memory (memory-bound; a few host threads for big batches)
    {
        const int nThreads = (words > (1u << 22)) ? std::min(8u, std::max(1u, std::thread::hardware_concurrency())) : 1;
        auto work = [&](int t) {
            for (int p = t; p < n_pairs; p += nThreads) {
                const twl_profile_pair &in = pairs[p];
                const bool r1 = packColumns(ctx->hProf.ptr + ctx->hPairs[p].refOff, in.freq_ref, in.gap_open_ref, in.gap_ext_ref, in.ref_len, P);
                const bool q1 = packColumns(ctx->hProf.ptr + ctx->hPairs[p].qryOff, in.freq_qry, in.gap_open_qry, in.gap_ext_qry, in.qry_len, P);
                ctx->hPairs[p].pad = (r1 ? twl::kRefOneHot : 0) | (q1 ? twl::kQryOneHot : 0);
            }
        };
        if (nThreads == 1) work(0);
        else {
            std::vector<std::thread> pool;
            for (int t = 1; t < nThreads; ++t) pool.emplace_back(work, t);
            work(0);
            for (auto &th : pool) th.join();
        }
    }
    // heaviest pairs first (anti-diagonal count is the serial length of a pair)
    ctx->hOrder.resize(n_pairs);
    std::iota(ctx->hOrder.begin(), ctx->hOrder.end(), 0);
    std::stable_sort(ctx->hOrder.begin(), ctx->hOrder.end(), [&](int x, int y) {
        return ctx->hPairs[x].refLen + ctx->hPairs[x].qryLen > ctx->hPairs[y].refLen + ctx->hPairs[y].qryLen;
    });

    {
        const size_t before = ctx->dProf.cap;
        TWL_CUDA(ctx, ctx->dProf.reserve(words + 2 * kProfPadWords));
        if (ctx->dProf.cap != before) TWL_CUDA(ctx, cudaMemsetAsync(ctx->dProf.ptr, 0, ctx->dProf.cap * sizeof(float), ctx->stream));
    }
    TWL_CUDA(ctx, ctx->dPairs.reserve(n_pairs));
    TWL_CUDA(ctx, ctx->dResults.reserve(n_pairs));
    TWL_CUDA(ctx, ctx->dPaths.reserve(bytes));
    TWL_CUDA(ctx, ctx->dOrder.reserve(n_pairs));
    TWL_CUDA(ctx, ctx->dOverflow.reserve(n_pairs));
    TWL_CUDA(ctx, ctx->dCounters.reserve(8));
    TWL_CUDA(ctx, cudaMemcpyAsync(ctx->dProf.ptr + kProfPadWords, ctx->hProf.ptr, words * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    TWL_CUDA(ctx, cudaMemcpyAsync(ctx->dPairs.ptr, ctx->hPairs.data(), n_pairs * sizeof(twl::DevPair), cudaMemcpyHostToDevice, ctx->stream));
    TWL_CUDA(ctx, cudaMemcpyAsync(ctx->dOrder.ptr, ctx->hOrder.data(), n_pairs * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    TWL_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // hPairs/hOrder are pageable
    ctx->staged = true;
    return TWL_OK;
}

} // extern "C"

// Kernel chain. Every stage reads its work list + count from device memory and appends the pairs whose band outgrew its
// capacity to the next stage's list, so the whole chain is enqueued without a host round trip.
//   nucleotide: wavefront<128> (band <= 509) -> wavefront<256> (band <= 1021) -> generic/global state (any band)
//   protein   : generic/shared state (band <= 1020)                            -> generic/global state
int twlLaunchDpChain(twl_ctx *ctx, int n, int wideCapIn) {
    const int marker = ctx->marker;
    if (ctx->hWatchdog.ptr && *ctx->hWatchdog.ptr && ctx->wideWorkers > 0) ctx->wideWorkers = 0;   // an earlier chain's wide workers timed out: stop co-running
    const int wideCap = std::max(wideCapIn, 8);                     // widest band any pair of the batch may legally reach
    const bool nucleotide = (ctx->P == 6) && !ctx->forceGeneric;
    struct Stage { int kind; int threads; int cap; int grid; size_t tbStride; int slots; };   // kind 0 wavefront (CTA per pair), 1 generic smem, 2 generic global, 3 warp per pair
    // Protein: similarity matrices first (talco_sim.cu), then the same register-resident wavefront stages reading them (matClass -1),
    // when the callers left the length bounds and the matrices fit the budget; otherwise the generic kernel scores on the fly.
    bool proteinSim = (ctx->P == 22) && ctx->proteinSim && !ctx->forceGeneric && static_cast<int>(ctx->hChainOrder.size()) == n && n > 0;
    int simMaxTiles = 0;
    if (proteinSim) {
        constexpr size_t kPad = 1024 + 16;                         // widest wavefront window + slack, in elements (even)
        size_t cur = 0;
        ctx->hSimInfo.assign(ctx->hSimRefUb.size(), twl::DevSim{0, 0, 0, 0});
        for (int x : ctx->hChainOrder) {
            if (x < 0 || x >= static_cast<int>(ctx->hSimRefUb.size()) || x >= static_cast<int>(ctx->hSimQryUb.size())) { proteinSim = false; break; }
            const size_t refUb = static_cast<size_t>(std::max(ctx->hSimRefUb[x], 1)), qryUb = static_cast<size_t>(std::max(ctx->hSimQryUb[x], 1));
            twl::DevSim &si = ctx->hSimInfo[x];
            si.stride = static_cast<int>((qryUb + 3) & ~static_cast<size_t>(3));
            si.gapOff = static_cast<long long>(cur + 2 * kPad);
            cur += 2 * (kPad + ((refUb + 1) & ~static_cast<size_t>(1)) + 16);
            si.simOff = static_cast<long long>(cur);
            cur += (refUb + qryUb) * static_cast<size_t>(si.stride) + kPad;
            simMaxTiles = std::max(simMaxTiles, twl::simMatrixTiles(static_cast<int>(refUb), static_cast<int>(qryUb)));
            if (cur * sizeof(float) > ctx->simBudgetBytes) { proteinSim = false; break; }
        }
        if (proteinSim) {
            TWL_CUDA(ctx, ctx->dSim.reserve(cur + 16));
            TWL_CUDA(ctx, ctx->dSimInfo.reserve(ctx->hSimInfo.size()));
            TWL_CUDA(ctx, cudaMemcpyAsync(ctx->dSimInfo.ptr, ctx->hSimInfo.data(), sizeof(twl::DevSim) * ctx->hSimInfo.size(), cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    const int matClass = nucleotide ? twl::nucleotideMatrixClass(ctx->hScore.data()) : (proteinSim ? -1 : 0);
    std::vector<Stage> stages;
    auto tbRows = [&](int w) { return (static_cast<size_t>(marker + 1) * w + 255) & ~static_cast<size_t>(255); };
    if (nucleotide || proteinSim) {
        // few pairs (every CTA has an SM to itself): 512 threads x 2 rows, one 1024-row window for any legal nucleotide band
        // (latency_shape 3: 512 x 1 first, 512 x 2 for pairs whose band outgrows 512 rows); otherwise 128 threads x 4 rows,
        // 5 CTAs per SM, with 512 x 2 as the wide stage
        const bool lowLatency = ctx->latencyMode != 0 && (ctx->latencyMode == 1 || n <= ctx->smCount);
        int plan[3][2] = {{128, 4}, {512, 2}, {0, 0}};
        if (lowLatency) {
            if (ctx->latencyShape == 3 && nucleotide) { plan[0][0] = 512; plan[0][1] = 1; }
            else { plan[0][0] = 512; plan[0][1] = 2; plan[1][0] = 0; }
        }
        for (int s = 0; s < 3 && plan[s][0]; ++s) {
            const int threads = plan[s][0], slots = plan[s][1];
            const int cap = twl::wavefrontBandCapacity(threads, slots);
            int perSm = std::max(1, twl::wavefrontMaxCtasPerSm(threads, slots, matClass));
            if (ctx->maxCtasPerSm > 0) perSm = std::min(perSm, ctx->maxCtasPerSm);
            stages.push_back({0, threads, cap, std::min(n, ctx->smCount * perSm), tbRows(twl::wavefrontWindow(threads, slots)), slots});
            if (wideCap <= cap) break;
        }
    } else {
        const int cap = std::min(kSmemStateCap, wideCap);
        stages.push_back({1, twl::genericThreads(), cap, std::min(n, ctx->smCount * kNarrowCtasPerSm), tbBytesPerCta(marker), 0});
    }
    if (wideCap > stages.back().cap) stages.push_back({2, twl::genericThreads(), wideCap, std::min(n, ctx->smCount), tbBytesPerCta(marker), 0});

    // Co-run: stage 0 (narrow window, 5 CTAs per SM) and stage 1 (wide window) execute at the same time; see TalcoArgs::coMode.
    // Only for levels of more than one wave of narrow CTAs (the regime it was built and measured for). With about one wave or less the
    // wide workers could only ever serve handed-over pairs and the gain is a few ms per level at best (10^5-leaf run: five such levels,
    // ~5 ms of wide stage each); those levels run the stages one after the other.
    int perSm0 = (nucleotide && !stages.empty() && stages[0].kind == 0) ? std::max(1, twl::wavefrontMaxCtasPerSm(stages[0].threads, stages[0].slots, matClass)) : 1;
    if (ctx->maxCtasPerSm > 0) perSm0 = std::min(perSm0, ctx->maxCtasPerSm);
    const bool coRun = nucleotide && ctx->wideWorkers > 0 && stages.size() >= 2 && stages[0].kind == 0 && stages[1].kind == 0 &&
                       stages[1].cap > stages[0].cap && n > ctx->smCount * perSm0;
    int wideGrid = 0;
    bool takeMain = false;
    if (coRun) {
        const int perSm = perSm0;
        // Many pairs per narrow CTA slot: a few wide workers that also eat from the main queue. About one wave or less: the
        // wide workers take the SMs the narrow kernel does not need and only serve handed-over pairs, which then restart at once.
        takeMain = n > ctx->smCount * perSm;
        if (takeMain) wideGrid = std::min(ctx->wideWorkers, ctx->smCount / 4);
        else wideGrid = std::min(ctx->smCount / 4, std::max(ctx->wideWorkers, (ctx->smCount * perSm - n) / perSm));
        stages[1].grid = wideGrid;
        stages[0].grid = std::min(n, (ctx->smCount - wideGrid) * perSm);
        if (!ctx->stream2) {
            int least = 0, greatest = 0;   // the wide workers' stream gets the highest priority: their CTAs are placed first when both kernels are pending
            TWL_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&least, &greatest));
            TWL_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->stream2, cudaStreamNonBlocking, greatest));
            TWL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->evFork, cudaEventDisableTiming));
            TWL_CUDA(ctx, cudaEventCreateWithFlags(&ctx->evJoin, cudaEventDisableTiming));
        }
    }
    // TWL_TRACE: stage times of the PREVIOUS chain of this thread (its events have completed by now: a level ends with a synchronisation)
    static const bool traceStages = std::getenv("TWL_TRACE") != nullptr;
    thread_local std::vector<cudaEvent_t> stageEv;
    thread_local int stageEvUsed = 0;
    if (traceStages) {
        if (stageEvUsed > 1) {
            std::fprintf(stderr, "[twl]   previous dp chain, ms per stage:");
            for (int e = 1; e < stageEvUsed; ++e) { float ms = 0.f; cudaEventElapsedTime(&ms, stageEv[e - 1], stageEv[e]); std::fprintf(stderr, " %.3f", ms); }
            std::fprintf(stderr, "\n");
        }
        stageEvUsed = 0;
    }
    auto markStage = [&]() {
        if (!traceStages) return;
        if (stageEvUsed == static_cast<int>(stageEv.size())) { cudaEvent_t e; cudaEventCreate(&e); stageEv.push_back(e); }
        cudaEventRecord(stageEv[stageEvUsed++], ctx->stream);
    };
    if (traceStages) {
        std::fprintf(stderr, "[twl]   dp chain: %d pairs, co-run %d, stages", n, coRun ? 1 : 0);
        for (const Stage &st : stages) std::fprintf(stderr, " [kind %d %dx%d cap %d grid %d]", st.kind, st.threads, st.slots, st.cap, st.grid);
        std::fprintf(stderr, "\n");
    }
    size_t tbBytes = 0, stateWords = 0;
    for (const Stage &st : stages) {
        tbBytes = std::max(tbBytes, st.tbStride * static_cast<size_t>(st.grid));
        if (st.kind == 2) stateWords = twl::genericStateWords(st.cap) * static_cast<size_t>(st.grid);
    }
    const size_t tbNarrow = coRun ? stages[0].tbStride * static_cast<size_t>(stages[0].grid) : 0;
    if (coRun) tbBytes = std::max({tbBytes, tbNarrow + stages[1].tbStride * static_cast<size_t>(stages[1].grid),
                                   stages[1].tbStride * static_cast<size_t>(std::min(n, ctx->smCount))});
    // Scratch is reserved for the largest plan a context of this alphabet can see, not for this chain: a buffer that regrows in the middle of
    // a run costs a cudaFree + cudaMalloc, which now and then stall for 100-700 ms (measured: the level of a 10^5-leaf run at which the
    // generic stage's band cap outgrew the 25 % slack of dState took 15 ms in most runs and up to 700 ms in others, tools/variance_levels.sh).
    {
        size_t tbCeil = static_cast<size_t>(ctx->smCount) * tbRows(1024);                                    // wide / low-latency wavefront, one CTA per SM
        tbCeil = std::max(tbCeil, static_cast<size_t>(ctx->smCount) * kNarrowCtasPerSm * tbBytesPerCta(marker));   // generic kernels
        if (nucleotide || proteinSim) {
            const int perSmN = std::max(1, twl::wavefrontMaxCtasPerSm(128, 4, matClass));
            tbCeil = std::max(tbCeil, static_cast<size_t>(ctx->smCount) * perSmN * tbRows(512) + static_cast<size_t>(ctx->smCount / 4) * tbRows(1024));   // narrow + co-running wide workers
        }
        TWL_CUDA(ctx, ctx->dTb.reserve(std::max(tbBytes, tbCeil)));
        if (stateWords) TWL_CUDA(ctx, ctx->dState.reserve(std::max(stateWords, twl::genericStateWords(4096) * static_cast<size_t>(ctx->smCount))));
    }
    const int nStages = static_cast<int>(stages.size());
    TWL_CUDA(ctx, ctx->dOverflow.reserve(static_cast<size_t>(n) * std::max(1, nStages - 1)));

    // counters: [2*s] = queue cursor of stage s, [2*s+1] = work count of stage s; co-run: [2] feed cursor, [3] feed count, [14] main pairs finished
    int counters[16] = {0};
    counters[1] = n;
    TWL_CUDA(ctx, ctx->dCounters.reserve(16));
    TWL_CUDA(ctx, cudaMemcpyAsync(ctx->dCounters.ptr, counters, sizeof(counters), cudaMemcpyHostToDevice, ctx->stream));
    if (coRun) TWL_CUDA(ctx, cudaMemsetAsync(ctx->dOverflow.ptr, 0xFF, sizeof(int) * static_cast<size_t>(n), ctx->stream));

    twl::TalcoArgs a{};
    a.prof = ctx->dProf.ptr + kProfPadWords;
    a.pairs = ctx->dPairs.ptr;
    a.results = ctx->dResults.ptr;
    a.paths = ctx->dPaths.ptr;
    a.marker = marker;
    a.gapOpen = ctx->gapOpen;
    a.gapExtend = ctx->gapExtend;
    a.score = ctx->dScore.ptr;
    if (ctx->P == 6) std::memcpy(a.scoreNt, ctx->hScore.data(), sizeof(a.scoreNt));
    a.tbScratch = ctx->dTb.ptr;
    if (proteinSim) {
        a.sim = ctx->dSim.ptr; a.simInfo = ctx->dSimInfo.ptr;
        TWL_CUDA(ctx, twl::launchSimMatrixAa(a.prof, a.pairs, ctx->dOrder.ptr, n, a.simInfo, ctx->dSim.ptr, a.score, simMaxTiles, ctx->stream));
        ctx->lastLaunches += (n + 65534) / 65535;
    }
    markStage();
    for (int s = 0; s < nStages; ++s) {
        const Stage &st = stages[s];
        const bool hasNext = (s + 1 < nStages);
        a.order = (s == 0) ? ctx->dOrder.ptr : ctx->dOverflow.ptr + static_cast<size_t>(n) * (s - 1);
        a.queue = ctx->dCounters.ptr + 2 * s;
        a.nWorkPtr = ctx->dCounters.ptr + 2 * s + 1;
        a.overflowList = hasNext ? ctx->dOverflow.ptr + static_cast<size_t>(n) * s : nullptr;
        a.overflowCount = hasNext ? ctx->dCounters.ptr + 2 * (s + 1) + 1 : nullptr;
        a.tbScratch = ctx->dTb.ptr;
        a.tbStride = st.tbStride;
        a.stateCap = st.cap;
        a.resume = (s > 0) ? 1 : 0;
        a.stateScratch = (st.kind == 2) ? ctx->dState.ptr : nullptr;
        a.stateStride = (st.kind == 2) ? twl::genericStateWords(st.cap) : 0;
        a.coMode = 0;
        if (coRun && s == 0) {
            // the wide workers go first so that they hold their SMs before the narrow CTAs fill the machine
            twl::TalcoArgs w = a;
            const Stage &sw = stages[1];
            const bool wideHasNext = (2 < nStages);
            w.coMode = 2;
            w.coTakeBelow = takeMain ? std::max(0, n - stages[0].grid) : 0;
            w.resume = 0;
            w.mainDone = ctx->dCounters.ptr + 14;
            w.heartbeat = ctx->dCounters.ptr + 15;
            w.watchdog = ctx->dCounters.ptr + 13;
            w.arrived = ctx->dCounters.ptr + 12;
            w.feedList = ctx->dOverflow.ptr;
            w.feedCount = ctx->dCounters.ptr + 3;
            w.feedCursor = ctx->dCounters.ptr + 2;
            w.overflowList = wideHasNext ? ctx->dOverflow.ptr + static_cast<size_t>(n) : nullptr;
            w.overflowCount = wideHasNext ? ctx->dCounters.ptr + 5 : nullptr;
            w.tbScratch = ctx->dTb.ptr + tbNarrow;
            w.tbStride = sw.tbStride;
            w.stateCap = sw.cap;
            unsigned long long *dTrace = nullptr;
            if (ctx->dpTrace) {
                TWL_CUDA(ctx, cudaMalloc(&dTrace, sizeof(unsigned long long) * 4 * (static_cast<size_t>(n) + 1)));
                TWL_CUDA(ctx, cudaMemsetAsync(dTrace, 0, sizeof(unsigned long long) * 4 * (static_cast<size_t>(n) + 1), ctx->stream));
                w.coTrace = dTrace; a.coTrace = dTrace;
            }
            TWL_CUDA(ctx, cudaEventRecord(ctx->evFork, ctx->stream));
            TWL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->evFork, 0));
            TWL_CUDA(ctx, twl::launchTalcoWavefront(sw.threads, sw.slots, matClass, w, sw.grid, ctx->stream2));
            TWL_CUDA(ctx, cudaEventRecord(ctx->evJoin, ctx->stream2));
            TWL_CUDA(ctx, twl::launchCoRunGate(ctx->dCounters.ptr + 12, sw.grid, ctx->stream));   // the narrow kernel starts once the wide workers are resident
            a.coMode = 1;
            a.mainDone = w.mainDone; a.heartbeat = w.heartbeat; a.feedList = w.feedList; a.feedCount = w.feedCount; a.feedCursor = w.feedCursor;
            a.overflowList = nullptr; a.overflowCount = nullptr;
            TWL_CUDA(ctx, twl::launchTalcoWavefront(st.threads, st.slots, matClass, a, st.grid, ctx->stream));
            TWL_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->evJoin, 0));
            ctx->lastLaunches += 3;   // wide workers, gate, narrow kernel
            if (dTrace) {
                cudaStreamSynchronize(ctx->stream);
                std::vector<unsigned long long> t(4 * (static_cast<size_t>(n) + 1));
                cudaMemcpy(t.data(), dTrace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
                cudaFree(dTrace);
                a.coTrace = nullptr;
                const unsigned long long endNarrow = t[4 * n], endWide = t[4 * n + 1];
                unsigned long long t0 = ~0ull;
                int mainByWide = 0;
                for (int p = 0; p < n; ++p) { if (t[4 * p]) t0 = std::min(t0, t[4 * p]); if (t[4 * p + 3] == 1) ++mainByWide; }
                std::fprintf(stderr, "[twl co-run] narrow grid %d, wide grid %d, main pairs done by wide workers %d; block 0 exits: narrow %+.3f ms, wide %+.3f ms after the first hand-over\n",
                             st.grid, sw.grid, mainByWide, (double)(long long)(endNarrow - t0) * 1e-6, (double)(long long)(endWide - t0) * 1e-6);
                for (int p = 0; p < n; ++p)
                    if (t[4 * p])
                        std::fprintf(stderr, "[twl co-run]   pair %5d handed over %+8.3f ms, taken after %7.3f ms, ran %7.3f ms\n", p, (double)(long long)(t[4 * p] - t0) * 1e-6,
                                     (double)(long long)(t[4 * p + 1] - t[4 * p]) * 1e-6, (double)(long long)(t[4 * p + 2] - t[4 * p + 1]) * 1e-6);
            }
            // clean-up: pairs handed over after the wide workers had left (only when the two kernels did not overlap); the
            // launch continues the feed cursor and exits at once when nothing is pending
            {
                twl::TalcoArgs c = w;
                c.coMode = 0; c.coTakeBelow = 0; c.coTrace = nullptr;
                c.order = w.feedList; c.queue = w.feedCursor; c.nWorkPtr = w.feedCount;
                c.resume = 1;
                c.tbScratch = ctx->dTb.ptr;
                TWL_CUDA(ctx, twl::launchTalcoWavefront(sw.threads, sw.slots, matClass, c, std::min(n, ctx->smCount), ctx->stream));
                ctx->lastLaunches += 1;
            }
            // did a wide worker give up waiting? read at the next chain (pinned copy, in stream order after the kernels)
            if (!ctx->hWatchdog.ptr) { TWL_CUDA(ctx, ctx->hWatchdog.reserve(1)); *ctx->hWatchdog.ptr = 0; }
            TWL_CUDA(ctx, cudaMemcpyAsync(ctx->hWatchdog.ptr, ctx->dCounters.ptr + 13, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            ++s;   // stage 1 ran alongside
            continue;
        }
        if (st.kind == 0) TWL_CUDA(ctx, twl::launchTalcoWavefront(st.threads, st.slots, matClass, a, st.grid, ctx->stream));
        else TWL_CUDA(ctx, twl::launchTalcoGeneric(ctx->P, st.kind == 2, a, st.grid, st.kind == 1 ? twl::genericStateWords(st.cap) * sizeof(float) : 0, ctx->stream));
        ctx->lastLaunches += 1;
        markStage();
        if (ctx->dpTrace) {   // diagnosis only: serialises the chain and prints each stage's time and work count
            thread_local cudaEvent_t e0 = nullptr, e1 = nullptr;   // one pair per host thread = per device (the adapter drives every device from its own thread)
            thread_local bool e0Recorded = false;
            if (!e0) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
            const bool firstReported = (s == 0) || !e0Recorded;     // the co-run branch above reports stages 0 + 1 itself and continues past this block
            if (firstReported) { cudaEventRecord(e0, ctx->stream); e0Recorded = true; }   // includes nothing before the first launch's completion
            cudaEventRecord(e1, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            int host[16];
            cudaMemcpy(host, ctx->dCounters.ptr, sizeof(host), cudaMemcpyDeviceToHost);
            float ms = 0.f;
            if (!firstReported) cudaEventElapsedTime(&ms, e0, e1);
            if (s + 1 == nStages) e0Recorded = false;
            std::fprintf(stderr, "[twl dp] stage %d kind %d %dx%d grid %d: work %d, +%.3f ms since stage 0 ended\n", s, st.kind, st.threads, st.slots, st.grid,
                         host[2 * s + 1], ms);
        }
    }
    return TWL_OK;
}

extern "C" {

int twl_batch_run(twl_ctx *ctx) {
    if (!ctx) return TWL_E_ARG;
    if (!ctx->staged) return fail(ctx, TWL_E_STATE, "twl_batch_run: no staged batch");
    cudaSetDevice(ctx->device);
    ctx->lastLaunches = 0;
    ctx->lastMs = 0.0f;
    ctx->timingPending = false;
    if (ctx->nPairs == 0) { ctx->ran = true; return TWL_OK; }
    TWL_CUDA(ctx, cudaEventRecord(ctx->evStart, ctx->stream));
    if (ctx->P == 22) {   // protein path: exact lengths are known here
        ctx->hSimRefUb.resize(ctx->nPairs); ctx->hSimQryUb.resize(ctx->nPairs);
        for (int p = 0; p < ctx->nPairs; ++p) { ctx->hSimRefUb[p] = ctx->hPairs[p].refLen; ctx->hSimQryUb[p] = ctx->hPairs[p].qryLen; }
        ctx->hChainOrder = ctx->hOrder;
    }
    const int rc = twlLaunchDpChain(ctx, ctx->nPairs, ctx->maxFLen);
    if (rc != TWL_OK) return rc;
    TWL_CUDA(ctx, cudaEventRecord(ctx->evStop, ctx->stream));
    ctx->timingPending = true;
    ctx->ran = true;
    return TWL_OK;
}

int twl_batch_fetch(twl_ctx *ctx, int8_t *const *paths, twl_pair_result *results) {
    if (!ctx) return TWL_E_ARG;
    if (!ctx->ran) return fail(ctx, TWL_E_STATE, "twl_batch_fetch: no batch has been run");
    cudaSetDevice(ctx->device);
    const int n = ctx->nPairs;
    if (n == 0) return TWL_OK;
    TWL_CUDA(ctx, ctx->hResults.reserve(n));
    TWL_CUDA(ctx, cudaMemcpyAsync(ctx->hResults.ptr, ctx->dResults.ptr, n * sizeof(twl::DevResult), cudaMemcpyDeviceToHost, ctx->stream));
    if (paths) {
        TWL_CUDA(ctx, ctx->hPaths.reserve(ctx->pathBytes));
        TWL_CUDA(ctx, cudaMemcpyAsync(ctx->hPaths.ptr, ctx->dPaths.ptr, ctx->pathBytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    TWL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < n; ++p) {
        const twl::DevResult &r = ctx->hResults.ptr[p];
        if (results) {
            results[p].status = r.status;
            results[p].path_len = r.pathLen;
            results[p].tiles = r.tiles;
            results[p].reserved = 0;
            results[p].cells = r.cells;
            results[p].diagonals = r.diagonals;
        }
        if (paths && paths[p] && r.status == 0 && r.pathLen > 0)
            std::memcpy(paths[p], ctx->hPaths.ptr + ctx->hPairs[p].alnOff, static_cast<size_t>(r.pathLen));
    }
    return TWL_OK;
}

int twl_align_profiles(twl_ctx *ctx, const twl_profile_pair *pairs, int n_pairs, int8_t *const *paths, twl_pair_result *results) {
    int rc = twl_batch_stage(ctx, pairs, n_pairs);
    if (rc != TWL_OK) return rc;
    rc = twl_batch_run(ctx);
    if (rc != TWL_OK) return rc;
    return twl_batch_fetch(ctx, paths, results);
}

int twl_set_option(twl_ctx *ctx, const char *name, int value) {
    if (!ctx || !name) return TWL_E_ARG;
    if (std::strcmp(name, "force_generic") == 0) { ctx->forceGeneric = value != 0; return TWL_OK; }
    if (std::strcmp(name, "wide_workers") == 0) { ctx->wideWorkers = std::max(0, value); return TWL_OK; }
    if (std::strcmp(name, "dp_trace") == 0) { ctx->dpTrace = value; return TWL_OK; }
    if (std::strcmp(name, "inject_nomem") == 0) { ctx->injectNomem = std::max(0, value); return TWL_OK; }
    if (std::strcmp(name, "max_ctas_per_sm") == 0) { ctx->maxCtasPerSm = std::max(0, value); return TWL_OK; }   // occupancy experiments (0 = what fits)
    if (std::strcmp(name, "latency_shape") == 0) { if (value != 2 && value != 3) return TWL_E_ARG; ctx->latencyShape = value; return TWL_OK; }   // 2: 512x2, 3: 512x1 then 512x2
    if (std::strcmp(name, "protein_sim") == 0) { ctx->proteinSim = value; return TWL_OK; }       // 0: proteins on the generic kernel only
    if (std::strcmp(name, "sim_budget_mb") == 0) { ctx->simBudgetBytes = static_cast<size_t>(std::max(1, value)) << 20; return TWL_OK; }
    if (std::strcmp(name, "latency_mode") == 0) { ctx->latencyMode = value; return TWL_OK; }     // -1 auto, 0 off, 1 always
    return fail(ctx, TWL_E_ARG, std::string("twl_set_option: unknown option ") + name);
}

int twl_selftest_division(twl_ctx *ctx, const float *num, const float *den, int n, int *mismatches) {
    if (!ctx || !num || !den || !mismatches || n < 0) return TWL_E_ARG;
    cudaSetDevice(ctx->device);
    float *dn = nullptr, *dd = nullptr;
    int *dm = nullptr;
    TWL_CUDA(ctx, cudaMalloc(&dn, sizeof(float) * std::max(n, 1)));
    TWL_CUDA(ctx, cudaMalloc(&dd, sizeof(float) * std::max(n, 1)));
    TWL_CUDA(ctx, cudaMalloc(&dm, sizeof(int)));
    TWL_CUDA(ctx, cudaMemcpyAsync(dn, num, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    TWL_CUDA(ctx, cudaMemcpyAsync(dd, den, sizeof(float) * n, cudaMemcpyHostToDevice, ctx->stream));
    TWL_CUDA(ctx, cudaMemsetAsync(dm, 0, sizeof(int), ctx->stream));
    TWL_CUDA(ctx, twl::launchDivSelfTest(dn, dd, n, dm, ctx->stream));
    TWL_CUDA(ctx, cudaMemcpyAsync(mismatches, dm, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    TWL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(dn); cudaFree(dd); cudaFree(dm);
    return TWL_OK;
}

float twl_last_kernel_ms(const twl_ctx *ctx) {
    if (!ctx) return -1.0f;
    twl_ctx *c = const_cast<twl_ctx *>(ctx);
    if (c->timingPending) {
        cudaSetDevice(c->device);
        float ms = 0.0f;
        if (cudaEventSynchronize(c->evStop) == cudaSuccess && cudaEventElapsedTime(&ms, c->evStart, c->evStop) == cudaSuccess) c->lastMs = ms;
        c->timingPending = false;
    }
    return c->lastMs;
}

int twl_last_launch_count(const twl_ctx *ctx) { return ctx ? ctx->lastLaunches : 0; }

} // extern "C"

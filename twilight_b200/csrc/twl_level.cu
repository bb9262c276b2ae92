// twl_level.cu — device-resident row store and the per-level pipeline (twl_rows_*, twl_align_level, twl_level_fetch).
// Host orchestration only; the kernels are in level_kernels.cuh and talco_*.cu.
#include "level_kernels.cuh"
#include "twl_ctx.hpp"

#include <algorithm>
#include <map>
#include <cstring>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <thread>

namespace {

struct RowSlot {
    char *buf[2] = {nullptr, nullptr};   // alnStorage[2]; the buffer a level writes (the one that is not live) is allocated when first needed
    int cap[2] = {0, 0};                 // and regrown on its own: a row that is never rewritten again never pays for a second buffer
    int len = 0, storage = 0;
    float weight = 0.f;
    bool present = false;
};

struct PairKeep {   // what twl_level_fetch needs from the last level
    long long rawOff[2] = {-1, -1}, consOff[2] = {-1, -1}, runsOff[2] = {-1, -1}, profOff[2] = {-1, -1};
    int alnLen[2] = {0, 0}, newLen[2] = {0, 0}, nRuns[2] = {0, 0};
    long long pathWoOff = -1;
    int pathWoLen = 0;
    std::vector<float> freq[2], merged;
    int chunk = 0;
};

} // namespace

struct TwlLevelState {
    std::vector<RowSlot> rows;
    std::vector<void *> pools;            // 256 MB chunks (or one oversized request each); kept across twl_rows_clear
    std::vector<size_t> poolBytes;
    size_t poolNext = 0;                  // first chunk the arena has not handed out yet
    char *poolCur = nullptr;
    size_t poolLeft = 0;
    std::map<size_t, std::vector<char *>> freeBufs;   // buffers given up by rows that outgrew them (memCheck), by size class

    DevBuf<twl::DevSide> dSides;
    DevBuf<const char *> dRowIn;
    DevBuf<char *> dRowOut;
    DevBuf<float> dRowW, dRaw, dFreq, dMerged;
    DevBuf<char> dCons;
    DevBuf<float> dGap;
    DevBuf<int> dRuns, dChunkCounts;
    DevBuf<twl::DevUpdate> dUps;
    DevBuf<int8_t> dFinalPaths;
    DevBuf<signed char> dAaLut;
    DevBuf<twl::DevUpdate> dUps2;
    DevBuf<int> dUpdPair, dWhich;
    DevBuf<long long> dNeed;
    DevBuf<char> dLargeScratch;
    DevBuf<twl::RestoreJob> dJobs;        // consensus alignments put off by the restore walks of a level chunk (level_kernels.cuh)
    DevBuf<int> dJobCount;
    DevBuf<const char *> dUpdIn;
    PinBuf<twl::DevResult> hRes;
    PinBuf<twl::DevUpdate> hUps;
    PinBuf<long long> hNeed;
    PinBuf<int8_t> hFinal;
    PinBuf<twl::DevSide> hSides;
    PinBuf<float> hFreqPin, hMergedPin;
    int largeRestores = 0;       // pairs whose gappy-column restore needed the global-scratch pass since the context was created
    DevBuf<twl::RowCopy> dCopies;
    DevBuf<char> dStage;
    PinBuf<char> hStage;
    bool lutReady = false;
    cudaEvent_t stageFree = nullptr;      // recorded after the last asynchronous use of hStage / dStage / dCopies
    bool stagePending = false;
    std::vector<cudaEvent_t> sliceEv;
    cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float phaseMs[4] = {0, 0, 0, 0};
    float restoreMs = 0.f;       // the gappy-column restore share of phaseMs[3]

    std::vector<PairKeep> keep;
    int lastChunks = 0;
    int P = 0;
};

namespace {

constexpr size_t kPoolChunkBytes = static_cast<size_t>(256) << 20;

TwlLevelState *levelOf(twl_ctx *ctx) {
    if (!ctx->level) {
        cudaSetDevice(ctx->device);           // the events below belong to the context's device, whatever device the caller had current
        ctx->level = new TwlLevelState();
        for (auto &e : ctx->level->ev) cudaEventCreate(&e);
        cudaEventCreateWithFlags(&ctx->level->stageFree, cudaEventDisableTiming);
    }
    return ctx->level;
}

// Host threads for staging copies: up to 8, but the host's cores are shared by the ranks of a multi-GPU job (TWL_HOST_THREADS overrides).
int hostCopyThreads() {
    static const int n = [] {
        if (const char *e = std::getenv("TWL_HOST_THREADS")) return std::max(1, std::atoi(e));
        int gpus = 1;   // processes sharing this host: torchrun / mpirun export it
        for (const char *name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "MPI_LOCALNRANKS"})
            if (const char *e = std::getenv(name)) { gpus = std::max(1, std::atoi(e)); break; }
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        return static_cast<int>(std::min<unsigned>(8u, std::max(1u, hw / static_cast<unsigned>(gpus))));
    }();
    return n;
}

// splits [0, n) over a few host threads (staging copies of many small rows are memory-latency bound on one core)
template <typename F>
void parallelRows(int n, size_t bytes, const F &body) {
    const int nThreads = (bytes > (static_cast<size_t>(4) << 20)) ? hostCopyThreads() : 1;
    if (nThreads == 1) { body(0, n); return; }
    std::vector<std::thread> pool;
    const int per = (n + nThreads - 1) / nThreads;
    for (int t = 1; t < nThreads; ++t) pool.emplace_back([&, t] { body(std::min(n, t * per), std::min(n, (t + 1) * per)); });
    body(0, std::min(n, per));
    for (auto &th : pool) th.join();
}

// cuts a transfer list into slices of about 8 MB (row boundaries): returns the first row of every slice plus n
std::vector<int> sliceRows(const std::vector<twl::RowCopy> &list, int n) {
    constexpr long long kSlice = 8 << 20;
    std::vector<int> cuts{0};
    long long start = 0;
    for (int i = 0; i < n; ++i)
        if (list[i].stageOff - start >= kSlice) { cuts.push_back(i); start = list[i].stageOff; }
    cuts.push_back(n);
    return cuts;
}

// Row buffers come in size classes (3 mantissa bits: at most 12.5 % slack) so that a buffer a row has outgrown can serve
// another row later instead of being abandoned in the arena.
size_t rowSizeClass(size_t bytes) {
    bytes = std::max<size_t>((bytes + 15) & ~static_cast<size_t>(15), 64);
    int top = 63 - __builtin_clzll(bytes);
    const size_t step = static_cast<size_t>(1) << std::max(top - 3, 4);
    return (bytes + step - 1) & ~(step - 1);
}

cudaError_t poolAlloc(TwlLevelState *L, size_t bytes, char **out) {
    bytes = rowSizeClass(bytes + 16);   // 16 bytes of slack: the level kernels read rows in aligned 32-bit words, up to 7 bytes past the text
    auto it = L->freeBufs.find(bytes);
    if (it != L->freeBufs.end() && !it->second.empty()) {
        *out = it->second.back();
        it->second.pop_back();
        return cudaSuccess;
    }
    while (bytes > L->poolLeft) {
        if (L->poolNext < L->pools.size()) {           // a chunk kept from before twl_rows_clear
            L->poolCur = static_cast<char *>(L->pools[L->poolNext]);
            L->poolLeft = L->poolBytes[L->poolNext];
            ++L->poolNext;
            continue;
        }
        // chunks double the arena (up to 8 GB each): cudaMalloc costs milliseconds per CALL whatever the size (measured on the B200:
        // 32 x 256 MB = 120-200 ms, 8 x 1 GB = 9 ms, 1 x 8 GB = 3 ms), and a 10^5-leaf run grows the arena to tens of GB
        size_t have = 0;
        for (size_t b : L->poolBytes) have += b;
        size_t sz = std::max(bytes, std::min(std::max(kPoolChunkBytes, have), static_cast<size_t>(8) << 30));
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, sz);
        if (e == cudaErrorMemoryAllocation && sz > std::max(kPoolChunkBytes, bytes)) {   // not that much left: the smallest chunk that serves the request
            cudaGetLastError();
            sz = std::max(kPoolChunkBytes, bytes);
            e = cudaMalloc(&p, sz);
        }
        if (e != cudaSuccess) return e;
        L->pools.push_back(p);
        L->poolBytes.push_back(sz);
        L->poolNext = L->pools.size();
        L->poolCur = static_cast<char *>(p);
        L->poolLeft = sz;
    }
    *out = L->poolCur;
    L->poolCur += bytes;
    L->poolLeft -= bytes;
    return cudaSuccess;
}

void poolRecycle(TwlLevelState *L, char *buf, size_t cap) {
    if (buf) L->freeBufs[rowSizeClass(cap + 16)].push_back(buf);
}

// letterIdx (src/scoring-matrix.cpp:26-79) on the host: only used to build the 256-entry protein lookup table the kernels read
int letterIndexHost(char type, char c) {
    if (c >= 'a' && c <= 'z') c = static_cast<char>(c - 32);
    if (type == 'p') {
        static const char *aa = "ACDEFGHIKLMNPQRSTVWY";
        if (c == '-' || c == '.') return 21;
        const char *hit = c ? std::strchr(aa, c) : nullptr;
        return hit ? static_cast<int>(hit - aa) : 20;
    }
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': case 'U': return 3;
    case '-': case '.': return 5;
    default: return 4;
    }
}

size_t sideWordsL(int len, int P) { return (P == 6) ? static_cast<size_t>((len + 3) / 4) * 32 : static_cast<size_t>(len) * (P + 2); }

} // namespace

void twlLevelDestroy(twl_ctx *ctx) {
    TwlLevelState *L = ctx->level;
    if (!L) return;
    for (void *p : L->pools) cudaFree(p);
    L->dSides.release(); L->dRowIn.release(); L->dRowOut.release(); L->dRowW.release(); L->dRaw.release(); L->dFreq.release();
    L->dMerged.release(); L->dCons.release(); L->dGap.release(); L->dRuns.release(); L->dChunkCounts.release(); L->dUps.release();
    L->dFinalPaths.release(); L->dAaLut.release(); L->dCopies.release(); L->dStage.release(); L->hStage.release();
    L->hRes.release(); L->hUps.release(); L->hNeed.release(); L->hFinal.release(); L->hSides.release(); L->hFreqPin.release(); L->hMergedPin.release();
    L->dUps2.release(); L->dUpdPair.release(); L->dNeed.release(); L->dWhich.release(); L->dLargeScratch.release(); L->dUpdIn.release(); L->dJobs.release(); L->dJobCount.release();
    for (auto &e : L->ev) if (e) cudaEventDestroy(e);
    for (auto &e : L->sliceEv) cudaEventDestroy(e);
    if (L->stageFree) cudaEventDestroy(L->stageFree);
    delete L;
    ctx->level = nullptr;
}

extern "C" {

int twl_rows_clear(twl_ctx *ctx) {
    if (!ctx) return TWL_E_ARG;
    TwlLevelState *L = levelOf(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    // the chunks stay allocated and are handed out again from the start (a divide-and-conquer run clears once per subtree);
    // twl_destroy frees them. TWL_E_NOMEM recovery (host adapter) relies on the row store being empty afterwards.
    L->poolNext = 0;
    L->poolCur = nullptr;
    L->poolLeft = 0;
    L->freeBufs.clear();
    L->rows.clear();
    return TWL_OK;
}

int twl_rows_upload(twl_ctx *ctx, int n, const int32_t *ids, const char *const *rows, const int32_t *lens, const float *weights) {
    if (!ctx) return TWL_E_ARG;
    if (n < 0 || (n > 0 && (!ids || !rows || !lens || !weights))) return twlFail(ctx, TWL_E_ARG, "twl_rows_upload: null argument");
    TwlLevelState *L = levelOf(ctx);
    cudaSetDevice(ctx->device);
    if (n == 0) return TWL_OK;
    std::vector<twl::RowCopy> list(n);
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
        if (ids[i] < 0 || lens[i] < 0) return twlFail(ctx, TWL_E_ARG, "twl_rows_upload: negative id or length");
        const int id = ids[i];
        if (static_cast<size_t>(id) >= L->rows.size()) L->rows.resize(id + 1);
        RowSlot &r = L->rows[id];
        const int cap = std::max(16, 2 * lens[i]);                        // timesBigger = 2, sequencedb.cpp:40
        if (!r.present || r.cap[0] < lens[i]) {
            if (r.present) { poolRecycle(L, r.buf[0], r.cap[0]); poolRecycle(L, r.buf[1], r.cap[1]); r.present = false; }
            TWL_CUDA(ctx, poolAlloc(L, cap, &r.buf[0]));
            r.cap[0] = cap; r.buf[1] = nullptr; r.cap[1] = 0;               // the second buffer comes with the first level that rewrites the row
        }
        r.len = lens[i]; r.storage = 0; r.weight = weights[i]; r.present = true;
        list[i].dev = r.buf[0]; list[i].stageOff = static_cast<long long>(total); list[i].len = lens[i]; list[i].pad = 0;
        total += (static_cast<size_t>(lens[i]) + 15) & ~static_cast<size_t>(15);
    }
    if (L->stagePending) { TWL_CUDA(ctx, cudaEventSynchronize(L->stageFree)); L->stagePending = false; }   // an earlier upload may still read the staging buffers
    TWL_CUDA(ctx, L->hStage.reserve(std::max<size_t>(total, 16)));
    TWL_CUDA(ctx, L->dStage.reserve(std::max<size_t>(total, 16)));
    TWL_CUDA(ctx, L->dCopies.reserve(n));
    // staged in slices: the H2D copy of one slice runs while the host fills the next; the call returns with the copies and
    // the scatter kernel enqueued (stream order makes later calls see the rows), the staging buffer is fenced by an event
    const std::vector<int> cuts = sliceRows(list, n);
    for (size_t c = 0; c + 1 < cuts.size(); ++c) {
        const int b0 = cuts[c], e0 = cuts[c + 1];
        const size_t from = static_cast<size_t>(list[b0].stageOff), to = (e0 < n) ? static_cast<size_t>(list[e0].stageOff) : total;
        parallelRows(e0 - b0, to - from, [&](int b, int e) { for (int i = b0 + b; i < b0 + e; ++i) std::memcpy(L->hStage.ptr + list[i].stageOff, rows[i], lens[i]); });
        TWL_CUDA(ctx, cudaMemcpyAsync(L->dStage.ptr + from, L->hStage.ptr + from, to - from, cudaMemcpyHostToDevice, ctx->stream));
    }
    TWL_CUDA(ctx, cudaMemcpyAsync(L->dCopies.ptr, list.data(), sizeof(twl::RowCopy) * n, cudaMemcpyHostToDevice, ctx->stream));
    twl::rowTransferKernel<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(L->dCopies.ptr, n, L->dStage.ptr, 1);
    TWL_CUDA(ctx, cudaGetLastError());
    TWL_CUDA(ctx, cudaEventRecord(L->stageFree, ctx->stream));
    L->stagePending = true;
    return TWL_OK;
}

int twl_rows_length(twl_ctx *ctx, int32_t id) {
    if (!ctx || !ctx->level || id < 0 || static_cast<size_t>(id) >= ctx->level->rows.size() || !ctx->level->rows[id].present) return -1;
    return ctx->level->rows[id].len;
}

int twl_rows_export(twl_ctx *ctx, int n, const int32_t *ids, void *dev_dst, size_t cap_bytes, int32_t *lens, int64_t *offsets) {
    if (!ctx) return TWL_E_ARG;
    if (n < 0 || (n > 0 && (!ids || !dev_dst || !lens || !offsets))) return twlFail(ctx, TWL_E_ARG, "twl_rows_export: null argument");
    TwlLevelState *L = levelOf(ctx);
    cudaSetDevice(ctx->device);
    if (n == 0) return TWL_OK;
    std::vector<twl::RowCopy> list(n);
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
        if (ids[i] < 0 || static_cast<size_t>(ids[i]) >= L->rows.size() || !L->rows[ids[i]].present)
            return twlFail(ctx, TWL_E_ARG, "twl_rows_export: unknown row id " + std::to_string(ids[i]));
        const RowSlot &r = L->rows[ids[i]];
        list[i].dev = r.buf[r.storage]; list[i].stageOff = static_cast<long long>(total); list[i].len = r.len; list[i].pad = 0;
        lens[i] = r.len; offsets[i] = static_cast<int64_t>(total);
        total += (static_cast<size_t>(r.len) + 15) & ~static_cast<size_t>(15);
    }
    if (total > cap_bytes) return twlFail(ctx, TWL_E_ARG, "twl_rows_export: destination too small");
    if (L->stagePending) { TWL_CUDA(ctx, cudaEventSynchronize(L->stageFree)); L->stagePending = false; }
    TWL_CUDA(ctx, L->dCopies.reserve(n));
    TWL_CUDA(ctx, cudaMemcpyAsync(L->dCopies.ptr, list.data(), sizeof(twl::RowCopy) * n, cudaMemcpyHostToDevice, ctx->stream));
    twl::rowTransferKernel<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(L->dCopies.ptr, n, static_cast<char *>(dev_dst), 0);
    TWL_CUDA(ctx, cudaGetLastError());
    TWL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TWL_OK;
}

int twl_rows_import(twl_ctx *ctx, int n, const int32_t *ids, const int32_t *lens, const float *weights, const void *dev_src,
                    const int64_t *offsets) {
    if (!ctx) return TWL_E_ARG;
    if (n < 0 || (n > 0 && (!ids || !lens || !weights || !dev_src || !offsets))) return twlFail(ctx, TWL_E_ARG, "twl_rows_import: null argument");
    TwlLevelState *L = levelOf(ctx);
    cudaSetDevice(ctx->device);
    if (n == 0) return TWL_OK;
    std::vector<twl::RowCopy> list(n);
    for (int i = 0; i < n; ++i) {
        if (ids[i] < 0 || lens[i] < 0 || offsets[i] < 0) return twlFail(ctx, TWL_E_ARG, "twl_rows_import: negative id, length or offset");
        const int id = ids[i];
        if (static_cast<size_t>(id) >= L->rows.size()) L->rows.resize(id + 1);
        RowSlot &r = L->rows[id];
        const int cap = std::max(16, 2 * lens[i]);
        if (!r.present || r.cap[0] < lens[i]) {
            if (r.present) { poolRecycle(L, r.buf[0], r.cap[0]); poolRecycle(L, r.buf[1], r.cap[1]); r.present = false; }
            TWL_CUDA(ctx, poolAlloc(L, cap, &r.buf[0]));
            r.cap[0] = cap; r.buf[1] = nullptr; r.cap[1] = 0;               // the second buffer comes with the first level that rewrites the row
        }
        r.len = lens[i]; r.storage = 0; r.weight = weights[i]; r.present = true;
        list[i].dev = r.buf[0]; list[i].stageOff = static_cast<long long>(offsets[i]); list[i].len = lens[i]; list[i].pad = 0;
    }
    if (L->stagePending) { TWL_CUDA(ctx, cudaEventSynchronize(L->stageFree)); L->stagePending = false; }
    TWL_CUDA(ctx, L->dCopies.reserve(n));
    TWL_CUDA(ctx, cudaMemcpyAsync(L->dCopies.ptr, list.data(), sizeof(twl::RowCopy) * n, cudaMemcpyHostToDevice, ctx->stream));
    twl::rowTransferKernel<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(L->dCopies.ptr, n, const_cast<char *>(static_cast<const char *>(dev_src)), 1);
    TWL_CUDA(ctx, cudaGetLastError());
    TWL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return TWL_OK;
}

int twl_rows_drop(twl_ctx *ctx, int n, const int32_t *ids) {
    if (!ctx) return TWL_E_ARG;
    if (n < 0 || (n > 0 && !ids)) return twlFail(ctx, TWL_E_ARG, "twl_rows_drop: null argument");
    TwlLevelState *L = levelOf(ctx);
    for (int i = 0; i < n; ++i) {
        if (ids[i] < 0 || static_cast<size_t>(ids[i]) >= L->rows.size() || !L->rows[ids[i]].present) continue;
        RowSlot &r = L->rows[ids[i]];
        poolRecycle(L, r.buf[0], r.cap[0]);    // everything that still reads the buffers is stream-ordered before their next use
        poolRecycle(L, r.buf[1], r.cap[1]);
        r = RowSlot();
    }
    return TWL_OK;
}

int twl_rows_migrate(twl_ctx *src, twl_ctx *dst, int n, const int32_t *ids) {
    if (!src || !dst) return TWL_E_ARG;
    if (n < 0 || (n > 0 && !ids)) return twlFail(src, TWL_E_ARG, "twl_rows_migrate: null argument");
    if (n == 0 || src == dst) return TWL_OK;
    TwlLevelState *LS = levelOf(src), *LD = levelOf(dst);
    // pack on the source device
    std::vector<twl::RowCopy> list(n);
    std::vector<int32_t> lens(n);
    std::vector<float> weights(n);
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
        if (ids[i] < 0 || static_cast<size_t>(ids[i]) >= LS->rows.size() || !LS->rows[ids[i]].present)
            return twlFail(src, TWL_E_ARG, "twl_rows_migrate: unknown row id " + std::to_string(ids[i]));
        const RowSlot &r = LS->rows[ids[i]];
        list[i].dev = r.buf[r.storage]; list[i].stageOff = static_cast<long long>(total); list[i].len = r.len; list[i].pad = 0;
        lens[i] = r.len; weights[i] = r.weight;
        total += (static_cast<size_t>(r.len) + 15) & ~static_cast<size_t>(15);
    }
    cudaSetDevice(src->device);
    if (LS->stagePending) { TWL_CUDA(src, cudaEventSynchronize(LS->stageFree)); LS->stagePending = false; }
    TWL_CUDA(src, LS->dStage.reserve(std::max<size_t>(total, 16)));
    TWL_CUDA(src, LS->dCopies.reserve(n));
    TWL_CUDA(src, cudaMemcpyAsync(LS->dCopies.ptr, list.data(), sizeof(twl::RowCopy) * n, cudaMemcpyHostToDevice, src->stream));
    twl::rowTransferKernel<<<(n * 32 + 255) / 256, 256, 0, src->stream>>>(LS->dCopies.ptr, n, LS->dStage.ptr, 0);
    TWL_CUDA(src, cudaGetLastError());
    TWL_CUDA(src, cudaStreamSynchronize(src->stream));
    // device to device (NVLink / NVSwitch when peer access is possible; the runtime stages through the host otherwise)
    cudaSetDevice(dst->device);
    if (LD->stagePending) { TWL_CUDA(dst, cudaEventSynchronize(LD->stageFree)); LD->stagePending = false; }
    TWL_CUDA(dst, LD->dStage.reserve(std::max<size_t>(total, 16)));
    if (src->device != dst->device) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, dst->device, src->device) == cudaSuccess && can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return twlFail(dst, TWL_E_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
        TWL_CUDA(dst, cudaMemcpyPeerAsync(LD->dStage.ptr, dst->device, LS->dStage.ptr, src->device, total, dst->stream));
    } else {
        TWL_CUDA(dst, cudaMemcpyAsync(LD->dStage.ptr, LS->dStage.ptr, total, cudaMemcpyDeviceToDevice, dst->stream));
    }
    // unpack on the destination device
    for (int i = 0; i < n; ++i) {
        const int id = ids[i];
        if (static_cast<size_t>(id) >= LD->rows.size()) LD->rows.resize(id + 1);
        RowSlot &r = LD->rows[id];
        const int cap = std::max(16, 2 * lens[i]);
        if (!r.present || r.cap[0] < lens[i]) {
            if (r.present) { poolRecycle(LD, r.buf[0], r.cap[0]); poolRecycle(LD, r.buf[1], r.cap[1]); r.present = false; }
            TWL_CUDA(dst, poolAlloc(LD, cap, &r.buf[0]));
            r.cap[0] = cap; r.buf[1] = nullptr; r.cap[1] = 0;
        }
        r.len = lens[i]; r.storage = 0; r.weight = weights[i]; r.present = true;
        list[i].dev = r.buf[0];
    }
    TWL_CUDA(dst, LD->dCopies.reserve(n));
    TWL_CUDA(dst, cudaMemcpyAsync(LD->dCopies.ptr, list.data(), sizeof(twl::RowCopy) * n, cudaMemcpyHostToDevice, dst->stream));
    twl::rowTransferKernel<<<(n * 32 + 255) / 256, 256, 0, dst->stream>>>(LD->dCopies.ptr, n, LD->dStage.ptr, 1);
    TWL_CUDA(dst, cudaGetLastError());
    TWL_CUDA(dst, cudaStreamSynchronize(dst->stream));     // `list` is read by the copy above; the source staging buffer is free again
    return twl_rows_drop(src, n, ids);
}

int twl_rows_lengths(twl_ctx *ctx, int n, const int32_t *ids, int32_t *lens) {
    if (!ctx || n < 0 || (n > 0 && (!ids || !lens))) return TWL_E_ARG;
    for (int i = 0; i < n; ++i) lens[i] = twl_rows_length(ctx, ids[i]);
    return TWL_OK;
}

int twl_rows_download(twl_ctx *ctx, int n, const int32_t *ids, char *const *dst, int32_t *lens) {
    if (!ctx) return TWL_E_ARG;
    if (n < 0 || (n > 0 && (!ids || !dst))) return twlFail(ctx, TWL_E_ARG, "twl_rows_download: null argument");
    TwlLevelState *L = levelOf(ctx);
    cudaSetDevice(ctx->device);
    if (n == 0) return TWL_OK;
    std::vector<twl::RowCopy> list(n);
    size_t total = 0;
    for (int i = 0; i < n; ++i) {
        if (ids[i] < 0 || static_cast<size_t>(ids[i]) >= L->rows.size() || !L->rows[ids[i]].present)
            return twlFail(ctx, TWL_E_ARG, "twl_rows_download: unknown row id " + std::to_string(ids[i]));
        const RowSlot &r = L->rows[ids[i]];
        list[i].dev = r.buf[r.storage]; list[i].stageOff = static_cast<long long>(total); list[i].len = r.len; list[i].pad = 0;
        total += (static_cast<size_t>(r.len) + 15) & ~static_cast<size_t>(15);
    }
    if (L->stagePending) { TWL_CUDA(ctx, cudaEventSynchronize(L->stageFree)); L->stagePending = false; }   // an earlier upload may still read the staging buffers
    TWL_CUDA(ctx, L->hStage.reserve(std::max<size_t>(total, 16)));
    TWL_CUDA(ctx, L->dStage.reserve(std::max<size_t>(total, 16)));
    TWL_CUDA(ctx, L->dCopies.reserve(n));
    TWL_CUDA(ctx, cudaMemcpyAsync(L->dCopies.ptr, list.data(), sizeof(twl::RowCopy) * n, cudaMemcpyHostToDevice, ctx->stream));
    twl::rowTransferKernel<<<(n * 32 + 255) / 256, 256, 0, ctx->stream>>>(L->dCopies.ptr, n, L->dStage.ptr, 0);
    TWL_CUDA(ctx, cudaGetLastError());
    // slices: the host unpacks one slice while the D2H copy of the next is in flight
    const std::vector<int> cuts = sliceRows(list, n);
    const size_t nSlices = cuts.size() - 1;
    while (L->sliceEv.size() < nSlices) { cudaEvent_t e; TWL_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); L->sliceEv.push_back(e); }
    for (size_t c = 0; c < nSlices; ++c) {
        const int b0 = cuts[c], e0 = cuts[c + 1];
        const size_t from = static_cast<size_t>(list[b0].stageOff), to = (e0 < n) ? static_cast<size_t>(list[e0].stageOff) : total;
        TWL_CUDA(ctx, cudaMemcpyAsync(L->hStage.ptr + from, L->dStage.ptr + from, to - from, cudaMemcpyDeviceToHost, ctx->stream));
        TWL_CUDA(ctx, cudaEventRecord(L->sliceEv[c], ctx->stream));
    }
    for (size_t c = 0; c < nSlices; ++c) {
        const int b0 = cuts[c], e0 = cuts[c + 1];
        const size_t from = static_cast<size_t>(list[b0].stageOff), to = (e0 < n) ? static_cast<size_t>(list[e0].stageOff) : total;
        TWL_CUDA(ctx, cudaEventSynchronize(L->sliceEv[c]));
        parallelRows(e0 - b0, to - from, [&](int b, int e) {
            for (int i = b0 + b; i < b0 + e; ++i) {
                std::memcpy(dst[i], L->hStage.ptr + list[i].stageOff, list[i].len);
                if (lens) lens[i] = list[i].len;
            }
        });
    }
    return TWL_OK;
}

int twl_level_update_split_ms(twl_ctx *ctx, float out[2]) {
    if (!ctx || !out) return TWL_E_ARG;
    out[0] = ctx->level ? ctx->level->restoreMs : 0.f;
    out[1] = ctx->level ? ctx->level->phaseMs[3] - ctx->level->restoreMs : 0.f;
    return TWL_OK;
}

int twl_level_phase_ms(twl_ctx *ctx, float out[4]) {
    if (!ctx || !out) return TWL_E_ARG;
    for (int i = 0; i < 4; ++i) out[i] = ctx->level ? ctx->level->phaseMs[i] : 0.f;
    return TWL_OK;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// twl_align_level
// ---------------------------------------------------------------------------------------------------------------
namespace {

// What runLevelChunk changed in the row store before its kernels ran: restored when the chunk fails (a CUDA error, out of
// memory), so that a caller who handles the error code still finds every row where it was.
struct RowUndo { int id; char *buf[2]; int cap[2], storage, len; };

// TWL_TRACE=1: wall-clock of the host-side steps of a level chunk on stderr
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    Trace() : on(std::getenv("TWL_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[twl] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

template <int P>
int runLevelChunkImpl(twl_ctx *ctx, TwlLevelState *L, const twl_level_pair *pairs, int begin, int end, int task, float threshold,
                      int cacheTh, int8_t *const *paths, twl_level_result *results, int chunkNo, std::vector<RowUndo> &journal) {
    using namespace twl;
    Trace tr;
    const int n = end - begin;
    const int nSides = 2 * n;
    const int defXdrop = static_cast<int>(1000 * -1 * ctx->gapExtend);
    std::vector<DevSide> sides(nSides);
    std::vector<const char *> rowIn;
    std::vector<float> rowW;
    std::vector<DevPair> dp(n);
    std::vector<float> freqStage;
    size_t rawWords = 0, consBytes = 0, runInts = 0, profWords = 0, pathBytes = 0, freqWords = 0;
    int maxLen = 0, maxF = 0;

    for (int p = 0; p < n; ++p) {
        const twl_level_pair &in = pairs[begin + p];
        PairKeep &kp = L->keep[begin + p];
        kp = PairKeep();
        kp.chunk = chunkNo;
        const twl_node_side *sd[2] = {&in.ref, &in.qry};
        const bool store = (in.ref.aln_num >= cacheTh || in.qry.aln_num >= cacheTh) || (in.ref.msa_freq || in.qry.msa_freq);   // helper.cpp:14
        for (int s = 0; s < 2; ++s) {
            const twl_node_side &nd = *sd[s];
            if (nd.aln_len < 0 || nd.aln_num < 1 || nd.n_ids < 0 || (nd.n_ids > 0 && !nd.seq_ids))
                return twlFail(ctx, TWL_E_ARG, "twl_align_level: malformed node in pair " + std::to_string(begin + p));
            if (!nd.msa_freq && nd.n_ids == 0)
                return twlFail(ctx, TWL_E_ARG, "twl_align_level: node without rows and without msa_freq in pair " + std::to_string(begin + p));
            DevSide &d = sides[2 * p + s];
            d.memberOff = static_cast<long long>(rowIn.size());
            for (int m = 0; m < nd.n_ids; ++m) {
                const int id = nd.seq_ids[m];
                if (id < 0 || static_cast<size_t>(id) >= L->rows.size() || !L->rows[id].present)
                    return twlFail(ctx, TWL_E_ARG, "twl_align_level: row " + std::to_string(id) + " is not resident");
                const RowSlot &r = L->rows[id];
                if (r.len != nd.aln_len) return twlFail(ctx, TWL_E_ARG, "twl_align_level: row " + std::to_string(id) + " has length " + std::to_string(r.len) + ", node says " + std::to_string(nd.aln_len));
                rowIn.push_back(r.buf[r.storage]);
                rowW.push_back(r.weight);
            }
            d.rawOff = static_cast<long long>(rawWords); rawWords += (static_cast<size_t>(nd.aln_len) * P + 3) & ~static_cast<size_t>(3);   // 16-byte aligned sides
            d.consOff = static_cast<long long>(consBytes); consBytes += (static_cast<size_t>(nd.aln_len) + 15) & ~static_cast<size_t>(15);
            d.runsOff = static_cast<long long>(runInts); runInts += 2 * (static_cast<size_t>(nd.aln_len) / 2 + 2);
            d.profOff = static_cast<long long>(profWords); profWords += sideWordsL(nd.aln_len, P);
            d.freqInOff = -1; d.freqOutOff = -1;
            if (nd.msa_freq) {
                d.freqInOff = static_cast<long long>(freqWords);
                freqStage.insert(freqStage.end(), nd.msa_freq, nd.msa_freq + static_cast<size_t>(nd.aln_len) * P);
                freqWords += static_cast<size_t>(nd.aln_len) * P;
            } else if (store) {
                d.freqOutOff = static_cast<long long>(freqWords);
                freqStage.resize(freqStage.size() + static_cast<size_t>(nd.aln_len) * P, 0.f);
                freqWords += static_cast<size_t>(nd.aln_len) * P;
            }
            d.nRows = nd.msa_freq ? 0 : nd.n_ids;
            d.alnLen = nd.aln_len; d.alnNum = nd.aln_num; d.nodeWeight = nd.aln_weight;
            d.pairIdx = p; d.isQry = s; d.newLen = 0; d.nRuns = 0;
            kp.rawOff[s] = d.rawOff; kp.consOff[s] = d.consOff; kp.runsOff[s] = d.runsOff; kp.profOff[s] = d.profOff; kp.alnLen[s] = nd.aln_len;
            maxLen = std::max(maxLen, nd.aln_len);
        }
        DevPair &q = dp[p];
        q.refOff = sides[2 * p].profOff; q.qryOff = sides[2 * p + 1].profOff;
        q.alnOff = static_cast<long long>(pathBytes); pathBytes += (static_cast<size_t>(in.ref.aln_len) + in.qry.aln_len + 15) & ~static_cast<size_t>(15);
        q.refLen = 0; q.qryLen = 0; q.refN4 = 0; q.qryN4 = 0;
        q.refNum = static_cast<float>(in.ref.aln_num); q.qryNum = static_cast<float>(in.qry.aln_num);
        q.gapChar = (task == 1 || task == 2 || in.ref.aln_num > 10000 || in.qry.aln_num > 10000) ? 0.0f : ctx->gapExtend;   // alignment-cpu.cpp:88
        q.xdrop = defXdrop; q.fLen = 4096; q.pad = 0;
        kp.pathWoOff = q.alnOff;
        maxF = std::max(maxF, std::min(q.fLen, std::min(in.ref.aln_len, in.qry.aln_len)));
    }

    tr.mark("describe level (host)");
    // ---- upload the level description
    TWL_CUDA(ctx, L->dSides.reserve(nSides));
    TWL_CUDA(ctx, L->dRowIn.reserve(std::max<size_t>(rowIn.size(), 1)));
    TWL_CUDA(ctx, L->dRowW.reserve(std::max<size_t>(rowW.size(), 1)));
    TWL_CUDA(ctx, L->dRaw.reserve(std::max<size_t>(rawWords, 1)));
    TWL_CUDA(ctx, L->dCons.reserve(std::max<size_t>(consBytes, 16)));
    if (P == 6) TWL_CUDA(ctx, L->dGap.reserve(std::max<size_t>(consBytes, 16)));   // per-column gap counts (floats, laid out like the consensus)
    TWL_CUDA(ctx, L->dRuns.reserve(std::max<size_t>(runInts, 2)));
    TWL_CUDA(ctx, L->dFreq.reserve(std::max<size_t>(freqWords, 1)));
    {
        const size_t before = ctx->dProf.cap;
        TWL_CUDA(ctx, ctx->dProf.reserve(profWords + 2 * kProfPadWords));
        if (ctx->dProf.cap != before) TWL_CUDA(ctx, cudaMemsetAsync(ctx->dProf.ptr, 0, ctx->dProf.cap * sizeof(float), ctx->stream));
    }
    tr.mark("  reserve level scratch");
    TWL_CUDA(ctx, ctx->dPairs.reserve(n));
    TWL_CUDA(ctx, ctx->dResults.reserve(n));
    TWL_CUDA(ctx, ctx->dPaths.reserve(std::max<size_t>(pathBytes, 16)));
    TWL_CUDA(ctx, ctx->dOrder.reserve(n));
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        return pairs[begin + x].ref.aln_len + pairs[begin + x].qry.aln_len > pairs[begin + y].ref.aln_len + pairs[begin + y].qry.aln_len;
    });
    if (P == 22 && !L->lutReady) {
        signed char lut[256];
        for (int c = 0; c < 256; ++c) lut[c] = static_cast<signed char>(letterIndexHost('p', static_cast<char>(c)));
        TWL_CUDA(ctx, L->dAaLut.reserve(256));
        TWL_CUDA(ctx, cudaMemcpyAsync(L->dAaLut.ptr, lut, 256, cudaMemcpyHostToDevice, ctx->stream));
        TWL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        L->lutReady = true;
    }
    TWL_CUDA(ctx, cudaMemcpyAsync(L->dSides.ptr, sides.data(), sizeof(DevSide) * nSides, cudaMemcpyHostToDevice, ctx->stream));
    if (!rowIn.empty()) {
        TWL_CUDA(ctx, cudaMemcpyAsync(L->dRowIn.ptr, rowIn.data(), sizeof(char *) * rowIn.size(), cudaMemcpyHostToDevice, ctx->stream));
        TWL_CUDA(ctx, cudaMemcpyAsync(L->dRowW.ptr, rowW.data(), sizeof(float) * rowW.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (freqWords) TWL_CUDA(ctx, cudaMemcpyAsync(L->dFreq.ptr, freqStage.data(), sizeof(float) * freqWords, cudaMemcpyHostToDevice, ctx->stream));
    TWL_CUDA(ctx, cudaMemcpyAsync(ctx->dPairs.ptr, dp.data(), sizeof(DevPair) * n, cudaMemcpyHostToDevice, ctx->stream));
    TWL_CUDA(ctx, cudaMemcpyAsync(ctx->dOrder.ptr, order.data(), sizeof(int) * n, cudaMemcpyHostToDevice, ctx->stream));

    tr.mark("  H2D level description");
    // ---- the update list is laid out before anything runs, with every path sized by its upper bound (ref + qry columns),
    // so the whole level — profiles, DP, gappy-column restore, row rewrite — is enqueued without a host round trip
    std::vector<int> work;
    for (int x : order) if (!(pairs[begin + x].flags & TWL_PAIR_PROFILE_ONLY)) work.push_back(x);
    std::vector<DevUpdate> ups;
    std::vector<const char *> updIn;
    std::vector<char *> updOut;
    std::vector<int> updPair, upOfPair(n, -1);
    size_t chunkInts = 0, mergedWords = 0, finalBytes = 0;
    int maxUb = 0, maxRows = 1;
    for (int p = 0; p < n; ++p) {
        const twl_level_pair &in = pairs[begin + p];
        if (in.flags & TWL_PAIR_PROFILE_ONLY) continue;
        const DevSide &sr = sides[2 * p], &sq = sides[2 * p + 1];
        const int ub = in.ref.aln_len + in.qry.aln_len;
        DevUpdate u;
        u.pathOff = static_cast<long long>(finalBytes);
        finalBytes += (static_cast<size_t>(ub) + 16) & ~static_cast<size_t>(15);
        u.chunkOff = static_cast<long long>(chunkInts);
        chunkInts += 2 * (static_cast<size_t>(ub + kPathChunk - 1) / kPathChunk + 1);
        u.memberOff = static_cast<long long>(updIn.size());
        u.pathLen = 0;                                                       // written by gappyRestoreKernel
        const bool touchRows = (task != 2) && !(in.flags & TWL_PAIR_NO_ROW_UPDATE);   // currentTask 2 only composes subtree paths (helper.cpp:384)
        u.nRef = touchRows ? in.ref.n_ids : 0;
        u.nQry = touchRows ? in.qry.n_ids : 0;
        u.refWeight = in.ref.aln_weight; u.qryWeight = in.qry.aln_weight; u.pad = 0;
        const twl_node_side *sd2[2] = {&in.ref, &in.qry};
        for (int s = 0; s < 2; ++s) {
            const int cnt = (s == 0) ? u.nRef : u.nQry;
            for (int m = 0; m < cnt; ++m) {
                RowSlot &r = L->rows[sd2[s]->seq_ids[m]];
                journal.push_back({sd2[s]->seq_ids[m], {r.buf[0], r.buf[1]}, {r.cap[0], r.cap[1]}, r.storage, r.len});
                updIn.push_back(r.buf[r.storage]);
                const int t = 1 - r.storage;                                 // the buffer this level writes
                if (r.cap[t] < ub) {                                         // SequenceInfo::memCheck, sequencedb.cpp:57-76 (timesBigger = 2)
                    char *fresh;
                    TWL_CUDA(ctx, poolAlloc(L, 2 * static_cast<size_t>(ub), &fresh));
                    r.buf[t] = fresh; r.cap[t] = 2 * ub;                     // the buffer it replaces (if any) is recycled once the level has succeeded
                }
                updOut.push_back(r.buf[t]);
                r.storage = 1 - r.storage;                                   // changeStorage(); undone below if the pair fails
            }
        }
        // updateFrequency applies when both nodes carry msaFreq after calculateProfile (helper.cpp:508)
        const bool bothFreq = (sr.freqInOff >= 0 || sr.freqOutOff >= 0) && (sq.freqInOff >= 0 || sq.freqOutOff >= 0);
        u.freqRefOff = u.freqQryOff = u.mergedOff = -1;
        if (bothFreq) {
            u.freqRefOff = (sr.freqInOff >= 0) ? sr.freqInOff : sr.freqOutOff;
            u.freqQryOff = (sq.freqInOff >= 0) ? sq.freqInOff : sq.freqOutOff;
            u.mergedOff = static_cast<long long>(mergedWords);
            mergedWords += static_cast<size_t>(ub) * P;
        }
        maxUb = std::max(maxUb, ub);
        maxRows = std::max(maxRows, u.nRef + u.nQry);
        upOfPair[p] = static_cast<int>(ups.size());
        ups.push_back(u);
        updPair.push_back(p);
    }
    const int nu = static_cast<int>(ups.size());
    tr.mark("  update list + row buffers");
    TWL_CUDA(ctx, L->hRes.reserve(n));
    TWL_CUDA(ctx, L->hUps.reserve(std::max(nu, 1)));
    TWL_CUDA(ctx, L->hNeed.reserve(2 * static_cast<size_t>(std::max(nu, 1))));
    TWL_CUDA(ctx, L->hSides.reserve(nSides));
    TWL_CUDA(ctx, L->hFreqPin.reserve(std::max<size_t>(freqWords, 1)));
    TWL_CUDA(ctx, L->hMergedPin.reserve(std::max<size_t>(mergedWords, 1)));
    tr.mark("  reserve pinned");
    if (nu) {
        TWL_CUDA(ctx, L->dUps.reserve(nu));
        TWL_CUDA(ctx, L->dUpdPair.reserve(nu));
        TWL_CUDA(ctx, L->dNeed.reserve(2 * static_cast<size_t>(nu)));
        TWL_CUDA(ctx, L->dFinalPaths.reserve(std::max<size_t>(finalBytes, 16)));
        TWL_CUDA(ctx, L->dChunkCounts.reserve(std::max<size_t>(chunkInts, 2)));
        TWL_CUDA(ctx, L->dUpdIn.reserve(std::max<size_t>(updIn.size(), 1)));
        TWL_CUDA(ctx, L->dRowOut.reserve(std::max<size_t>(updOut.size(), 1)));
        TWL_CUDA(ctx, L->dMerged.reserve(std::max<size_t>(mergedWords, 1)));
        TWL_CUDA(ctx, cudaMemcpyAsync(L->dUps.ptr, ups.data(), sizeof(DevUpdate) * nu, cudaMemcpyHostToDevice, ctx->stream));
        TWL_CUDA(ctx, cudaMemcpyAsync(L->dUpdPair.ptr, updPair.data(), sizeof(int) * nu, cudaMemcpyHostToDevice, ctx->stream));
        if (!updIn.empty()) {
            TWL_CUDA(ctx, cudaMemcpyAsync(L->dUpdIn.ptr, updIn.data(), sizeof(char *) * updIn.size(), cudaMemcpyHostToDevice, ctx->stream));
            TWL_CUDA(ctx, cudaMemcpyAsync(L->dRowOut.ptr, updOut.data(), sizeof(char *) * updOut.size(), cudaMemcpyHostToDevice, ctx->stream));
        }
    }
    // final paths travel to the host only when the caller gave a destination for at least one pair of the chunk
    bool wantPaths = false;
    if (paths) for (int p = 0; p < n && !wantPaths; ++p) wantPaths = paths[begin + p] != nullptr && upOfPair[p] >= 0;
    if (wantPaths) TWL_CUDA(ctx, L->hFinal.reserve(std::max<size_t>(finalBytes, 16)));
    DevResult *res = L->hRes.ptr;
    for (int p = 0; p < n; ++p) { res[p].status = 0; res[p].pathLen = 0; res[p].tiles = 0; res[p].pad = 0; res[p].cells = 0; res[p].diagonals = 0; }

    tr.mark("reserve + H2D description");
    // ---- phase 1: profiles + consensus (+ msaFreq cache); phase 2: gappy-column compaction + PSGP + DP packing
    TWL_CUDA(ctx, cudaEventRecord(L->ev[0], ctx->stream));
    if (P == 6) {
        // many sides: one block per side walks its column tiles; few sides: one block per tile for parallelism
        const int tiles = std::max(1, (maxLen + kProfThreads * kProfCols - 1) / (kProfThreads * kProfCols));
        dim3 grid(nSides, (nSides >= 8 * ctx->smCount) ? 1 : tiles);
        profileBuildNtKernel<<<grid, kProfThreads, 0, ctx->stream>>>(L->dSides.ptr, L->dRowIn.ptr, L->dRowW.ptr, L->dRaw.ptr, L->dCons.ptr,
                                                                    L->dFreq.ptr, L->dFreq.ptr, L->dGap.ptr);
        TWL_CUDA(ctx, cudaGetLastError());
        TWL_CUDA(ctx, cudaEventRecord(L->ev[1], ctx->stream));
        gappyCompactNtKernel<<<nSides, kLvlThreads, 0, ctx->stream>>>(L->dSides.ptr, L->dRaw.ptr, L->dGap.ptr, ctx->dProf.ptr + kProfPadWords, L->dRuns.ptr,
                                                                     ctx->dPairs.ptr, threshold, ctx->gapOpen, ctx->gapExtend);
        TWL_CUDA(ctx, cudaGetLastError());
    } else {
        dim3 grid(nSides, std::max(1, (maxLen + kLvlThreads - 1) / kLvlThreads));
        profileBuildKernel<P><<<grid, kLvlThreads, 0, ctx->stream>>>(L->dSides.ptr, L->dRowIn.ptr, L->dRowW.ptr, L->dRaw.ptr, L->dCons.ptr,
                                                                 L->dFreq.ptr, L->dFreq.ptr, L->dAaLut.ptr);
        TWL_CUDA(ctx, cudaGetLastError());
        TWL_CUDA(ctx, cudaEventRecord(L->ev[1], ctx->stream));
        gappyCompactKernel<P><<<nSides, kLvlThreads, 0, ctx->stream>>>(L->dSides.ptr, L->dRaw.ptr, ctx->dProf.ptr + kProfPadWords, L->dRuns.ptr,
                                                                       ctx->dPairs.ptr, threshold, ctx->gapOpen, ctx->gapExtend);
        TWL_CUDA(ctx, cudaGetLastError());
    }
    TWL_CUDA(ctx, cudaEventRecord(L->ev[2], ctx->stream));
    ctx->lastLaunches += 2;

    tr.mark("  launch profile kernels");
    // ---- phase 3: DP chain (pairs flagged profile-only are simply not listed). Task 0 reports failed pairs to the caller;
    // tasks 1 and 2 retry them with wider limits (alignment-cpu.cpp:116-129), which needs the statuses on the host.
    bool first = true;
    while (!work.empty()) {
        const int nw = static_cast<int>(work.size());
        TWL_CUDA(ctx, cudaMemcpyAsync(ctx->dOrder.ptr, work.data(), sizeof(int) * nw, cudaMemcpyHostToDevice, ctx->stream));
        if (P == 22) {   // protein path: profile lengths before gappy-column removal bound the similarity matrices
            ctx->hSimRefUb.resize(n); ctx->hSimQryUb.resize(n);
            for (int p = 0; p < n; ++p) { ctx->hSimRefUb[p] = pairs[begin + p].ref.aln_len; ctx->hSimQryUb[p] = pairs[begin + p].qry.aln_len; }
            ctx->hChainOrder = work;
        }
        int rc = twlLaunchDpChain(ctx, nw, maxF);
        if (rc != TWL_OK) return rc;
        if (first) TWL_CUDA(ctx, cudaEventRecord(L->ev[3], ctx->stream));
        first = false;
        if (task == 0) break;
        TWL_CUDA(ctx, cudaMemcpyAsync(res, ctx->dResults.ptr, sizeof(DevResult) * n, cudaMemcpyDeviceToHost, ctx->stream));
        TWL_CUDA(ctx, cudaMemcpyAsync(dp.data(), ctx->dPairs.ptr, sizeof(DevPair) * n, cudaMemcpyDeviceToHost, ctx->stream));
        TWL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        std::vector<int> again;
        for (int x : work) {
            const DevResult &r = res[x];
            if (r.status == 0 || r.status == kStatusEmptySide || r.status == 3) continue;
            const int minLen = std::min(dp[x].refLen, dp[x].qryLen);
            if (r.status == 2) dp[x].fLen = std::min(static_cast<int>(dp[x].fLen * 1.2) << 1, minLen);
            else { dp[x].xdrop = dp[x].xdrop * 2; dp[x].fLen = std::min(static_cast<int>(dp[x].xdrop * 4) << 1, minLen); }
            maxF = std::max(maxF, std::min(dp[x].fLen, minLen));
            again.push_back(x);
        }
        if (!again.empty()) TWL_CUDA(ctx, cudaMemcpyAsync(ctx->dPairs.ptr, dp.data(), sizeof(DevPair) * n, cudaMemcpyHostToDevice, ctx->stream));
        work.swap(again);
    }
    if (first) TWL_CUDA(ctx, cudaEventRecord(L->ev[3], ctx->stream));   // nothing to align in this chunk
    tr.mark("  launch dp chain");

    // ---- phase 4: gappy columns back (helper.cpp:324-375), row rewrite + frequency merge (helper.cpp:377-448, 506-539)
    TWL_CUDA(ctx, cudaEventRecord(L->ev[5], ctx->stream));
    auto launchUpdate = [&](const DevUpdate *dUps, int count, int maxPath, int rowsMax) -> int {
        pathChunkKernel<<<(count * 32 + 255) / 256, 256, 0, ctx->stream>>>(dUps, count, L->dFinalPaths.ptr, L->dChunkCounts.ptr);
        TWL_CUDA(ctx, cudaGetLastError());
        const int chunks = std::max(1, (maxPath + kPathChunk - 1) / kPathChunk);
        // enough blocks to fill the GPU when a level has few pairs with many member rows each
        const int wantZ = std::max(1, (4 * ctx->smCount) / std::max(1, count * chunks));
        dim3 grid(count, chunks, std::min(std::min(rowsMax, wantZ), 64));
        rowUpdateKernel<P><<<grid, kLvlThreads, 0, ctx->stream>>>(dUps, L->dFinalPaths.ptr, L->dChunkCounts.ptr, L->dUpdIn.ptr, L->dRowOut.ptr,
                                                              L->dFreq.ptr, L->dMerged.ptr);
        TWL_CUDA(ctx, cudaGetLastError());
        ctx->lastLaunches += 2;
        return TWL_OK;
    };
    if (nu) {
        // walk (one warp per pair, consensus alignments of coinciding runs put off) -> all put-off alignments in parallel -> compaction
        const char *jobEnv = std::getenv("TWL_RESTORE_JOBS");           // capacity of the job list; 0: align in line (A/B); a full list also aligns in line
        const int jobCap = jobEnv ? std::max(0, std::atoi(jobEnv)) : (1 << 20);
        TWL_CUDA(ctx, L->dJobs.reserve(std::max(jobCap, 1)));
        TWL_CUDA(ctx, L->dJobCount.reserve(1));
        TWL_CUDA(ctx, cudaMemsetAsync(L->dJobCount.ptr, 0, sizeof(int), ctx->stream));
        gappyRestoreKernel<false><<<std::min(nu, ctx->smCount * 32), 32, 0, ctx->stream>>>(
            L->dUps.ptr, L->dUpdPair.ptr, nullptr, nu, ctx->dPairs.ptr, ctx->dResults.ptr, L->dSides.ptr, L->dRuns.ptr, L->dCons.ptr, ctx->dPaths.ptr,
            L->dFinalPaths.ptr, ctx->dScore.ptr, ctx->M, P == 6 ? 0 : 1, L->dAaLut.ptr, ctx->gapOpen, ctx->gapExtend, L->dNeed.ptr, nullptr, 0, 0,
            jobCap ? L->dJobs.ptr : nullptr, L->dJobCount.ptr, jobCap);
        TWL_CUDA(ctx, cudaGetLastError());
        ctx->lastLaunches += 1;
        if (jobCap) {
            consensusJobsKernel<<<ctx->smCount * 8, 32, 0, ctx->stream>>>(L->dJobs.ptr, L->dJobCount.ptr, jobCap, L->dCons.ptr, L->dFinalPaths.ptr, ctx->dScore.ptr,
                                                                     ctx->M, P == 6 ? 0 : 1, L->dAaLut.ptr, ctx->gapOpen, ctx->gapExtend);
            TWL_CUDA(ctx, cudaGetLastError());
            pathCompactKernel<<<nu, kLvlThreads, 0, ctx->stream>>>(L->dUps.ptr, nu, L->dFinalPaths.ptr);
            TWL_CUDA(ctx, cudaGetLastError());
            ctx->lastLaunches += 2;
        }
        TWL_CUDA(ctx, cudaEventRecord(L->ev[6], ctx->stream));
        const int rc = launchUpdate(L->dUps.ptr, nu, maxUb, maxRows);
        if (rc != TWL_OK) return rc;
    }
    TWL_CUDA(ctx, cudaEventRecord(L->ev[4], ctx->stream));
    tr.mark("  launch restore + update");

    // ---- results back to the host: one synchronisation per chunk
    DevSide *hs = L->hSides.ptr;
    TWL_CUDA(ctx, cudaMemcpyAsync(res, ctx->dResults.ptr, sizeof(DevResult) * n, cudaMemcpyDeviceToHost, ctx->stream));
    TWL_CUDA(ctx, cudaMemcpyAsync(hs, L->dSides.ptr, sizeof(DevSide) * nSides, cudaMemcpyDeviceToHost, ctx->stream));
    if (nu) {
        TWL_CUDA(ctx, cudaMemcpyAsync(L->hUps.ptr, L->dUps.ptr, sizeof(DevUpdate) * nu, cudaMemcpyDeviceToHost, ctx->stream));
        TWL_CUDA(ctx, cudaMemcpyAsync(L->hNeed.ptr, L->dNeed.ptr, sizeof(long long) * 2 * nu, cudaMemcpyDeviceToHost, ctx->stream));
        if (wantPaths) TWL_CUDA(ctx, cudaMemcpyAsync(L->hFinal.ptr, L->dFinalPaths.ptr, finalBytes, cudaMemcpyDeviceToHost, ctx->stream));
    }
    float *hFreq = L->hFreqPin.ptr, *hMerged = L->hMergedPin.ptr;
    if (freqWords) TWL_CUDA(ctx, cudaMemcpyAsync(hFreq, L->dFreq.ptr, sizeof(float) * freqWords, cudaMemcpyDeviceToHost, ctx->stream));
    if (mergedWords) TWL_CUDA(ctx, cudaMemcpyAsync(hMerged, L->dMerged.ptr, sizeof(float) * mergedWords, cudaMemcpyDeviceToHost, ctx->stream));
    TWL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    tr.mark("kernels + D2H results");

    // ---- pairs with a coinciding run pair too large for the kernel's shared memory: the same kernel again with its matrices in
    // a global scratch buffer sized from what the first pass asked for, then the row update of those pairs
    std::vector<int> redo;
    long long wantCells = kRestoreTbCells, wantCols = kRestoreRowCap;
    for (int k = 0; k < nu; ++k)
        if (L->hNeed.ptr[2 * k] > 0) {
            redo.push_back(k);
            wantCells = std::max(wantCells, L->hNeed.ptr[2 * k]);
            wantCols = std::max(wantCols, L->hNeed.ptr[2 * k + 1]);
        }
    if (!redo.empty()) {
        const int nr = static_cast<int>(redo.size());
        wantCells = (wantCells + 15) & ~15ll;
        wantCols = (wantCols + 15) & ~15ll;
        const size_t stride = static_cast<size_t>(wantCells) + 25ull * static_cast<size_t>(wantCols) + 64;
        const int blocks = static_cast<int>(std::max<size_t>(1, std::min<size_t>(nr, (static_cast<size_t>(1) << 30) / stride)));
        TWL_CUDA(ctx, L->dLargeScratch.reserve(stride * blocks));
        TWL_CUDA(ctx, L->dWhich.reserve(nr));
        TWL_CUDA(ctx, cudaMemcpyAsync(L->dWhich.ptr, redo.data(), sizeof(int) * nr, cudaMemcpyHostToDevice, ctx->stream));
        gappyRestoreKernel<true><<<blocks, 32, 0, ctx->stream>>>(
            L->dUps.ptr, L->dUpdPair.ptr, L->dWhich.ptr, nr, ctx->dPairs.ptr, ctx->dResults.ptr, L->dSides.ptr, L->dRuns.ptr, L->dCons.ptr, ctx->dPaths.ptr,
            L->dFinalPaths.ptr, ctx->dScore.ptr, ctx->M, P == 6 ? 0 : 1, L->dAaLut.ptr, ctx->gapOpen, ctx->gapExtend, L->dNeed.ptr, L->dLargeScratch.ptr,
            wantCells, static_cast<int>(wantCols), nullptr, nullptr, 0);
        TWL_CUDA(ctx, cudaGetLastError());
        ctx->lastLaunches += 1;
        TWL_CUDA(ctx, cudaMemcpyAsync(L->hUps.ptr, L->dUps.ptr, sizeof(DevUpdate) * nu, cudaMemcpyDeviceToHost, ctx->stream));
        TWL_CUDA(ctx, cudaMemcpyAsync(L->hNeed.ptr, L->dNeed.ptr, sizeof(long long) * 2 * nu, cudaMemcpyDeviceToHost, ctx->stream));
        TWL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        std::vector<DevUpdate> again;
        int maxPath = 0, rowsMax = 1;
        for (int k : redo) {
            if (L->hNeed.ptr[2 * k] > 0) return twlFail(ctx, TWL_E_STATE, "twl_align_level: gappy-column restore did not fit its scratch");
            const DevUpdate &u = L->hUps.ptr[k];
            again.push_back(u);
            maxPath = std::max(maxPath, u.pathLen);
            rowsMax = std::max(rowsMax, u.nRef + u.nQry);
        }
        TWL_CUDA(ctx, L->dUps2.reserve(again.size()));
        TWL_CUDA(ctx, cudaMemcpyAsync(L->dUps2.ptr, again.data(), sizeof(DevUpdate) * again.size(), cudaMemcpyHostToDevice, ctx->stream));
        const int rc = launchUpdate(L->dUps2.ptr, nr, maxPath, rowsMax);
        if (rc != TWL_OK) return rc;
        if (wantPaths) TWL_CUDA(ctx, cudaMemcpyAsync(L->hFinal.ptr, L->dFinalPaths.ptr, finalBytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (mergedWords) TWL_CUDA(ctx, cudaMemcpyAsync(hMerged, L->dMerged.ptr, sizeof(float) * mergedWords, cudaMemcpyDeviceToHost, ctx->stream));
        TWL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        L->largeRestores += nr;
    }

    // ---- per pair: results, final path, row bookkeeping
    for (int p = 0; p < n; ++p) {
        const twl_level_pair &in = pairs[begin + p];
        twl_level_result &out = results[begin + p];
        PairKeep &kp = L->keep[begin + p];
        const DevSide &sr = hs[2 * p], &sq = hs[2 * p + 1];
        for (int s = 0; s < 2; ++s) {
            const DevSide &d = hs[2 * p + s];
            kp.newLen[s] = d.newLen; kp.nRuns[s] = d.nRuns;
            if (d.freqOutOff >= 0) kp.freq[s].assign(hFreq + d.freqOutOff, hFreq + d.freqOutOff + static_cast<size_t>(d.alnLen) * P);
        }
        out.status = res[p].status; out.tiles = res[p].tiles; out.cells = res[p].cells; out.diagonals = res[p].diagonals;
        out.ref_len_dp = sr.newLen; out.qry_len_dp = sq.newLen; out.path_len = 0;
        out.cached = (sr.freqOutOff >= 0 ? 1 : 0) | (sq.freqOutOff >= 0 ? 2 : 0);
        if (in.flags & TWL_PAIR_PROFILE_ONLY) { out.status = 0; continue; }
        const int k = upOfPair[p];
        const DevUpdate &u = L->hUps.ptr[k];
        kp.pathWoLen = (out.status == 0) ? res[p].pathLen : 0;
        const twl_node_side *sd2[2] = {&in.ref, &in.qry};
        long long at = u.memberOff;
        for (int s = 0; s < 2; ++s) {
            const int cnt = (s == 0) ? u.nRef : u.nQry;
            for (int m = 0; m < cnt; ++m, ++at) {
                RowSlot &r = L->rows[sd2[s]->seq_ids[m]];
                if (out.status == 0) { r.len = u.pathLen; continue; }
                r.storage = 1 - r.storage;                                   // the pair failed: the row keeps its old content (the live buffer was not touched)
            }
        }
        if (out.status != 0) continue;
        out.path_len = u.pathLen;
        if (u.mergedOff >= 0) {
            out.cached |= 4;
            kp.merged.assign(hMerged + u.mergedOff, hMerged + u.mergedOff + static_cast<size_t>(u.pathLen) * P);
        }
    }
    if (wantPaths && nu)
        parallelRows(n, finalBytes, [&](int b, int e) {
            for (int p = b; p < e; ++p) {
                const twl_level_result &out = results[begin + p];
                if (upOfPair[p] < 0 || out.status != 0 || !paths[begin + p]) continue;
                std::memcpy(paths[begin + p], L->hFinal.ptr + L->hUps.ptr[upOfPair[p]].pathOff, static_cast<size_t>(out.path_len));
            }
        });
    tr.mark("results to caller (host)");
    const bool updateTimed = nu > 0;
    for (int i = 0; i < 3; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, L->ev[i], L->ev[i + 1]);
        L->phaseMs[i] += ms;
    }
    if (updateTimed) {   // gappy-column restore + path chunk counts + row rewrite / frequency merge
        float ms = 0.f;
        cudaEventElapsedTime(&ms, L->ev[5], L->ev[4]);
        L->phaseMs[3] += ms;
        cudaEventElapsedTime(&ms, L->ev[5], L->ev[6]);
        L->restoreMs += ms;
    }
    return TWL_OK;
}

template <int P>
int runLevelChunk(twl_ctx *ctx, TwlLevelState *L, const twl_level_pair *pairs, int begin, int end, int task, float threshold,
                  int cacheTh, int8_t *const *paths, twl_level_result *results, int chunkNo, std::vector<RowUndo> &journal) {
    return runLevelChunkImpl<P>(ctx, L, pairs, begin, end, task, threshold, cacheTh, paths, results, chunkNo, journal);
}

} // namespace

extern "C" {

int twl_align_level(twl_ctx *ctx, const twl_level_pair *pairs, int n_pairs, int current_task, float gappy_threshold, int32_t cache_threshold,
                    int8_t *const *paths, twl_level_result *results) {
    if (!ctx) return TWL_E_ARG;
    if (ctx->P == 0) return twlFail(ctx, TWL_E_STATE, "twl_align_level: call twl_set_params first");
    if (n_pairs < 0 || (n_pairs > 0 && (!pairs || !results))) return twlFail(ctx, TWL_E_ARG, "twl_align_level: bad arguments");
    TwlLevelState *L = levelOf(ctx);
    cudaSetDevice(ctx->device);
    ctx->lastLaunches = 0;
    ctx->staged = false;   // the batch buffers are reused
    ctx->ran = false;
    L->keep.assign(n_pairs, PairKeep());
    L->P = ctx->P;
    for (float &m : L->phaseMs) m = 0.f;
    L->restoreMs = 0.f;
    if (cache_threshold <= 0) cache_threshold = 1000;
    // chunk the level so that the scratch (raw profiles dominate: 4*P bytes per column and side) stays bounded
    size_t budget = static_cast<size_t>(12) << 30;
    if (const char *e = std::getenv("TWL_LEVEL_BUDGET_MB")) budget = static_cast<size_t>(std::max(1, std::atoi(e))) << 20;   // tests force small chunks
    // The call is all or nothing for the row store: `journal` remembers every row a chunk touched; when any chunk fails
    // (CUDA error, out of memory) every row points at its old buffers again (a chunk only ever writes the OTHER buffer of a
    // row, or freshly allocated ones), so a caller that handles the error code finds the rows as they were before the call.
    std::vector<RowUndo> journal;
    int begin = 0, chunkNo = 0;
    while (begin < n_pairs) {
        int end = begin;
        size_t bytes = 0;
        while (end < n_pairs) {
            size_t need = (static_cast<size_t>(pairs[end].ref.aln_len) + pairs[end].qry.aln_len) * (ctx->P * 4 + (ctx->P + 2) * 4 + 16) +
                          (static_cast<size_t>(pairs[end].ref.n_ids) + pairs[end].qry.n_ids) * 24;
            if (ctx->P == 22 && ctx->proteinSim)   // the pair's similarity matrix (anti-diagonal-major: (ref + qry) x qry floats)
                need += (static_cast<size_t>(pairs[end].ref.aln_len) + pairs[end].qry.aln_len) * (static_cast<size_t>(pairs[end].qry.aln_len) + 4) * 4;
            if (end > begin && bytes + need > budget) break;
            bytes += need;
            ++end;
        }
        const int rc = (ctx->P == 6) ? runLevelChunk<6>(ctx, L, pairs, begin, end, current_task, gappy_threshold, cache_threshold, paths, results, chunkNo, journal)
                                     : runLevelChunk<22>(ctx, L, pairs, begin, end, current_task, gappy_threshold, cache_threshold, paths, results, chunkNo, journal);
        int rcEff = rc;
        if (rc == TWL_OK && end >= n_pairs && ctx->injectNomem > 0 && --ctx->injectNomem == 0)
            rcEff = twlFail(ctx, TWL_E_NOMEM, "twl_align_level: injected out-of-memory failure (twl_set_option inject_nomem)");
        if (rcEff != TWL_OK) {
            const int rc = rcEff;
            const std::string why = ctx->error;
            cudaStreamSynchronize(ctx->stream);
            for (auto it = journal.rbegin(); it != journal.rend(); ++it) {
                RowSlot &r = L->rows[it->id];
                r.buf[0] = it->buf[0]; r.buf[1] = it->buf[1]; r.cap[0] = it->cap[0]; r.cap[1] = it->cap[1]; r.storage = it->storage; r.len = it->len;
            }
            cudaGetLastError();
            ctx->error = why;
            return rc;
        }
        begin = end;
        ++chunkNo;
    }
    // buffers that regrown rows left behind serve other rows from now on (everything that read them is stream-ordered before)
    for (const RowUndo &u : journal) {
        const RowSlot &r = L->rows[u.id];
        for (int b = 0; b < 2; ++b)
            if (u.buf[b] && r.buf[0] != u.buf[b] && r.buf[1] != u.buf[b]) poolRecycle(L, u.buf[b], u.cap[b]);
    }
    L->lastChunks = chunkNo;
    float total = 0.f;
    for (float m : L->phaseMs) total += m;
    ctx->lastMs = total;
    ctx->timingPending = false;
    return TWL_OK;
}

int twl_level_large_restores(const twl_ctx *ctx) { return (ctx && ctx->level) ? ctx->level->largeRestores : 0; }

int twl_level_fetch(twl_ctx *ctx, int pair, int what, void *dst, size_t cap_bytes, size_t *out_bytes) {
    if (!ctx || !ctx->level) return TWL_E_ARG;
    TwlLevelState *L = ctx->level;
    if (pair < 0 || static_cast<size_t>(pair) >= L->keep.size()) return twlFail(ctx, TWL_E_ARG, "twl_level_fetch: pair index out of range");
    cudaSetDevice(ctx->device);
    const PairKeep &kp = L->keep[pair];
    const int P = L->P;
    const void *src = nullptr;
    size_t bytes = 0;
    std::vector<float> tmp;
    const int s = what & 1;
    switch (what) {
    case TWL_F_PROFILE_RAW_REF: case TWL_F_PROFILE_RAW_QRY: {
        if (kp.chunk != L->lastChunks - 1) return twlFail(ctx, TWL_E_STATE, "twl_level_fetch: raw profiles are only kept for the last chunk of a level");
        tmp.resize(static_cast<size_t>(kp.alnLen[s]) * P);
        TWL_CUDA(ctx, cudaMemcpy(tmp.data(), L->dRaw.ptr + kp.rawOff[s], tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
        src = tmp.data(); bytes = tmp.size() * sizeof(float);
        break;
    }
    case TWL_F_DP_PROFILE_REF: case TWL_F_DP_PROFILE_QRY: {
        if (kp.chunk != L->lastChunks - 1) return twlFail(ctx, TWL_E_STATE, "twl_level_fetch: packed profiles are only kept for the last chunk of a level");
        const int len = kp.newLen[s], PW = P + 2;
        const size_t words = (P == 6) ? static_cast<size_t>((len + 3) / 4) * 32 : static_cast<size_t>(len) * PW;
        std::vector<float> packed(std::max<size_t>(words, 1));
        TWL_CUDA(ctx, cudaMemcpy(packed.data(), ctx->dProf.ptr + kProfPadWords + kp.profOff[s], words * sizeof(float), cudaMemcpyDeviceToHost));
        tmp.resize(static_cast<size_t>(len) * PW);
        if (P == 6) {
            const int n4 = (len + 3) / 4;
            for (int c = 0; c < len; ++c) {
                const float *x = packed.data() + twl::ntColIndex(c, n4) * 4, *y = x + static_cast<size_t>(16) * n4;
                float *d = tmp.data() + static_cast<size_t>(c) * PW;
                d[0] = x[0]; d[1] = x[1]; d[2] = x[2]; d[3] = x[3]; d[4] = y[0]; d[5] = y[1]; d[6] = y[2]; d[7] = y[3];
            }
        } else std::memcpy(tmp.data(), packed.data(), tmp.size() * sizeof(float));
        src = tmp.data(); bytes = tmp.size() * sizeof(float);
        break;
    }
    case TWL_F_CONSENSUS_REF: case TWL_F_CONSENSUS_QRY: case TWL_F_RUNS_REF: case TWL_F_RUNS_QRY: case TWL_F_PATH_WO: {
        // intermediates stay on the device (nothing on the production path reads them); fetched on demand
        if (kp.chunk != L->lastChunks - 1) return twlFail(ctx, TWL_E_STATE, "twl_level_fetch: intermediates are only kept for the last chunk of a level");
        const void *dev = nullptr;
        if (what == TWL_F_PATH_WO) { bytes = static_cast<size_t>(kp.pathWoLen); dev = ctx->dPaths.ptr + kp.pathWoOff; }
        else if (what == TWL_F_CONSENSUS_REF || what == TWL_F_CONSENSUS_QRY) { bytes = static_cast<size_t>(kp.alnLen[s]); dev = L->dCons.ptr + kp.consOff[s]; }
        else { bytes = static_cast<size_t>(kp.nRuns[s]) * 2 * sizeof(int32_t); dev = L->dRuns.ptr + kp.runsOff[s]; }
        tmp.resize((bytes + 3) / 4 + 1);
        if (bytes) TWL_CUDA(ctx, cudaMemcpy(tmp.data(), dev, bytes, cudaMemcpyDeviceToHost));
        src = tmp.data();
        break;
    }
    case TWL_F_FREQ_REF: src = kp.freq[0].data(); bytes = kp.freq[0].size() * sizeof(float); break;
    case TWL_F_FREQ_QRY: src = kp.freq[1].data(); bytes = kp.freq[1].size() * sizeof(float); break;
    case TWL_F_FREQ_MERGED: src = kp.merged.data(); bytes = kp.merged.size() * sizeof(float); break;
    default: return twlFail(ctx, TWL_E_ARG, "twl_level_fetch: unknown selector");
    }
    if (out_bytes) *out_bytes = bytes;
    if (dst) {
        if (bytes > cap_bytes) return twlFail(ctx, TWL_E_ARG, "twl_level_fetch: destination too small");
        if (bytes) std::memcpy(dst, src, bytes);
    }
    return TWL_OK;
}

} // extern "C"


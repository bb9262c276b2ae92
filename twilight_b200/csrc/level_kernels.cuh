// level_kernels.cuh — the HBM-bound kernels around the DP of one guide-tree level:
//   profileBuildKernel   calculateProfile + getConsensus + msaFreq cache   (src/alignment-helper.cpp:8-72, 221-241)
//   gappyCompactKernel   removeGappyColumns + calculatePSGP, fused, writing the DP kernels' packed layout directly
//                        (src/alignment-helper.cpp:74-166, 168-219)
//   pathChunkKernel      per-1024-op prefix counts of a final alignment path
//   rowUpdateKernel      alignment_helper::updateAlignment row rewrite + updateFrequency merge
//                        (src/alignment-helper.cpp:377-448, 506-539)
// All arithmetic that reaches a float result follows the reference's operation order and types.
#pragma once
#include "twl_device.cuh"

namespace twl {

constexpr int kLvlThreads = 256;
constexpr int kPathChunk = 2048;   // path ops per block of rowUpdateKernel (8 per thread): most pairs of a ~1.5 kb level are one block

// One side (node) of a pair as the level kernels see it.
struct DevSide {
    long long memberOff;    // first entry of this side in the member arrays (row pointers, weights)
    long long rawOff;       // floats: raw profile [alnLen][P] (calculateProfile output)
    long long consOff;      // bytes : consensus [alnLen]
    long long freqInOff;    // floats: cached msaFreq uploaded by the caller, or -1
    long long freqOutOff;   // floats: where the msaFreq cache is written when storeFreq, or -1
    long long profOff;      // floats: packed DP-layout columns of this side inside the DP `prof` buffer
    long long runsOff;      // ints  : (start,len) pairs of removed column runs
    int nRows, alnLen, alnNum;
    float nodeWeight;
    int pairIdx, isQry;     // which DevPair field this side fills
    int newLen, nRuns;      // outputs of gappyCompactKernel
};

struct DevUpdate {          // one pair of rowUpdateKernel
    long long pathOff;      // bytes: final path (with gappy columns) of this pair
    long long chunkOff;     // ints : per-chunk prefix counts (2 per chunk: ref-consuming, qry-consuming)
    long long memberOff;    // first member (ref members, then qry members) in rowIn/rowOut
    long long freqRefOff, freqQryOff, mergedOff;   // floats, -1 when the frequency merge does not apply
    int pathLen, nRef, nQry;
    float refWeight, qryWeight;
    int pad;
};

__device__ __forceinline__ int letterIndexNt(unsigned char c) {   // letterIdx after toupper, scoring-matrix.cpp:55-79
    if (c >= 'a' && c <= 'z') c -= 32;
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': case 'U': return 3;
    case '-': case '.': return 5;
    default: return 4;
    }
}

__device__ __forceinline__ int letterIndexAa(unsigned char c, const signed char *lut) { return lut[c]; }

// ---------------------------------------------------------------------------------------------------------------
// calculateProfile: thread = one column; member rows are visited in seqsIncluded order so the float accumulation
// order equals the reference's. Accumulators live in shared memory ([letter][thread], conflict free) because the
// letter index is data dependent.
// ---------------------------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(kLvlThreads) profileBuildKernel(const DevSide *sides, const char *const *rowPtr, const float *rowWeight,
                                                                  float *raw, char *cons, const float *freqIn, float *freqOut,
                                                                  const signed char *aaLut) {
    __shared__ float acc[P][kLvlThreads];
    // grid.x = side (the dimension with the 2^31 limit), grid.y = block of kLvlThreads columns
    const DevSide sd = sides[blockIdx.x];
    const int t = blockIdx.y * kLvlThreads + threadIdx.x;
    if (blockIdx.y * kLvlThreads >= sd.alnLen) return;
    const bool live = t < sd.alnLen;
    float col[P];
    if (sd.freqInOff >= 0) {                                               // helper.cpp:16-21
        if (live) {
#pragma unroll
            for (int v = 0; v < P; ++v) col[v] = __fmul_rn(__fdiv_rn(freqIn[sd.freqInOff + static_cast<long long>(t) * P + v], sd.nodeWeight), static_cast<float>(sd.alnNum));
        }
    } else {                                                               // helper.cpp:23-34
#pragma unroll
        for (int v = 0; v < P; ++v) acc[v][threadIdx.x] = 0.0f;
        const char *const *rows = rowPtr + sd.memberOff;
        const float *wts = rowWeight + sd.memberOff;
        for (int s = 0; s < sd.nRows; ++s) {
            const float w = __fmul_rn(__fdiv_rn(wts[s], sd.nodeWeight), static_cast<float>(sd.alnNum));
            if (live) {
                const unsigned char ch = static_cast<unsigned char>(rows[s][t]);
                const int letter = (P == 6) ? letterIndexNt(ch) : letterIndexAa(ch, aaLut);
                acc[letter][threadIdx.x] = __fadd_rn(acc[letter][threadIdx.x], w);
            }
        }
#pragma unroll
        for (int v = 0; v < P; ++v) col[v] = acc[v][threadIdx.x];
    }
    if (!live) return;
    float *dst = raw + sd.rawOff + static_cast<long long>(t) * P;
#pragma unroll
    for (int v = 0; v < P; ++v) dst[v] = col[v];
    if (sd.freqOutOff >= 0) {                                              // helper.cpp:35-40
        float *f = freqOut + sd.freqOutOff + static_cast<long long>(t) * P;
#pragma unroll
        for (int v = 0; v < P; ++v) f[v] = __fmul_rn(__fdiv_rn(col[v], static_cast<float>(sd.alnNum)), sd.nodeWeight);
    }
    // getConsensus, helper.cpp:221-241: strict > from 0 over the first P-2 letters, default = the ambiguity letter
    int best = P - 2;
    float top = 0.0f;
#pragma unroll
    for (int v = 0; v < P - 2; ++v)
        if (col[v] > top) { top = col[v]; best = v; }
    const char *lut = (P == 6) ? "ACGTN" : "ACDEFGHIKLMNPQRSTVWYX";
    cons[sd.consOff + t] = lut[best];
}

// Nucleotide version (P = 6) of the same computation for the HBM roofline: a thread owns FOUR consecutive columns, reads
// each member row with one 32-bit load (rows are 16-byte aligned and over-allocated, so the word that holds the last
// columns may be read whole), and writes its 24 profile floats with six 16-byte stores (the side's raw offset is a multiple
// of four floats). Accumulation per column is still member by member in seqsIncluded order; the letter index comes from a
// 256-entry table in shared memory. Also emits the per-column gap counts as a compact float array for the compaction pass.
constexpr int kProfThreads = 128;
constexpr int kProfCols = 4;
__global__ void __launch_bounds__(kProfThreads) profileBuildNtKernel(const DevSide *sides, const char *const *rowPtr, const float *rowWeight,
                                                                      float *raw, char *cons, const float *freqIn, float *freqOut, float *gapCount) {
    constexpr int P = 6;
    __shared__ float acc[P][kProfCols][kProfThreads];
    __shared__ unsigned char lut[256];
    const DevSide sd = sides[blockIdx.x];
    if (blockIdx.y * kProfThreads * kProfCols >= sd.alnLen) return;
    for (int c = threadIdx.x; c < 256; c += kProfThreads) lut[c] = static_cast<unsigned char>(letterIndexNt(static_cast<unsigned char>(c)));
    // grid.y column tiles of 512 columns; with many sides the host launches grid.y = 1 and the block walks the tiles of its side
    // (the side descriptor and the member pointers are fetched once, fewer and fatter blocks)
    for (int tile = blockIdx.y; tile * kProfThreads * kProfCols < sd.alnLen; tile += gridDim.y) {
    const int t0 = (tile * kProfThreads + threadIdx.x) * kProfCols;           // first of this thread's columns
    const int nHere = min(kProfCols, sd.alnLen - t0);                          // <= 0: nothing to do for this thread
    float col[kProfCols][P];
    __syncthreads();                                                       // the previous tile's accumulators have been read
    if (sd.freqInOff >= 0) {                                               // helper.cpp:16-21
#pragma unroll
        for (int c = 0; c < kProfCols; ++c)
#pragma unroll
            for (int v = 0; v < P; ++v)
                col[c][v] = (c < nHere) ? __fmul_rn(__fdiv_rn(freqIn[sd.freqInOff + static_cast<long long>(t0 + c) * P + v], sd.nodeWeight), static_cast<float>(sd.alnNum)) : 0.0f;
        __syncthreads();
    } else {                                                               // helper.cpp:23-34
#pragma unroll
        for (int v = 0; v < P; ++v)
#pragma unroll
            for (int c = 0; c < kProfCols; ++c) acc[v][c][threadIdx.x] = 0.0f;
        __syncthreads();
        const char *const *rows = rowPtr + sd.memberOff;
        const float *wts = rowWeight + sd.memberOff;
        const float numF = static_cast<float>(sd.alnNum);
        if (nHere > 0) {
            int s = 0;
            for (; s + 4 <= sd.nRows; s += 4) {                            // four member rows in flight
                unsigned w4[4];
                float ww[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    w4[u] = __ldg(reinterpret_cast<const unsigned *>(rows[s + u] + t0));
                    ww[u] = __fmul_rn(__fdiv_rn(wts[s + u], sd.nodeWeight), numF);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int c = 0; c < kProfCols; ++c) {
                        float *a = &acc[lut[(w4[u] >> (8 * c)) & 0xFFu]][c][threadIdx.x];
                        *a = __fadd_rn(*a, ww[u]);
                    }
            }
            for (; s < sd.nRows; ++s) {
                const unsigned w4 = __ldg(reinterpret_cast<const unsigned *>(rows[s] + t0));
                const float ww = __fmul_rn(__fdiv_rn(wts[s], sd.nodeWeight), numF);
#pragma unroll
                for (int c = 0; c < kProfCols; ++c) {
                    float *a = &acc[lut[(w4 >> (8 * c)) & 0xFFu]][c][threadIdx.x];
                    *a = __fadd_rn(*a, ww);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < kProfCols; ++c)
#pragma unroll
            for (int v = 0; v < P; ++v) col[c][v] = acc[v][c][threadIdx.x];
    }
    if (nHere <= 0) continue;
    float *dst = raw + sd.rawOff + static_cast<long long>(t0) * P;
    float *fdst = (sd.freqOutOff >= 0) ? freqOut + sd.freqOutOff + static_cast<long long>(t0) * P : nullptr;
    const char *letters = "ACGTN";
    if (nHere == kProfCols) {
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        const float *flat = &col[0][0];
#pragma unroll
        for (int x = 0; x < 6; ++x) d4[x] = make_float4(flat[4 * x], flat[4 * x + 1], flat[4 * x + 2], flat[4 * x + 3]);
        if (fdst) {                                                        // helper.cpp:35-40
#pragma unroll
            for (int x = 0; x < 24; ++x) fdst[x] = __fmul_rn(__fdiv_rn(flat[x], static_cast<float>(sd.alnNum)), sd.nodeWeight);
        }
    } else {
#pragma unroll
        for (int c = 0; c < kProfCols; ++c)
#pragma unroll
            for (int v = 0; v < P; ++v)
                if (c < nHere) {
                    dst[c * P + v] = col[c][v];
                    if (fdst) fdst[c * P + v] = __fmul_rn(__fdiv_rn(col[c][v], static_cast<float>(sd.alnNum)), sd.nodeWeight);
                }
    }
    // getConsensus, helper.cpp:221-241, and the gap counts for the compaction pass
    unsigned packed = 0;
    float g4[kProfCols];
#pragma unroll
    for (int c = 0; c < kProfCols; ++c) {
        int best = P - 2;
        float top = 0.0f;
#pragma unroll
        for (int v = 0; v < P - 2; ++v)
            if (col[c][v] > top) { top = col[c][v]; best = v; }
        packed |= static_cast<unsigned>(static_cast<unsigned char>(letters[best])) << (8 * c);
        g4[c] = col[c][P - 1];
    }
    if (nHere == kProfCols) {
        *reinterpret_cast<unsigned *>(cons + sd.consOff + t0) = packed;
        *reinterpret_cast<float4 *>(gapCount + sd.consOff + t0) = make_float4(g4[0], g4[1], g4[2], g4[3]);
    } else {
#pragma unroll
        for (int c = 0; c < kProfCols; ++c)
            if (c < nHere) { cons[sd.consOff + t0 + c] = static_cast<char>((packed >> (8 * c)) & 0xFFu); gapCount[sd.consOff + t0 + c] = g4[c]; }
    }
    }
}

// Batched row transfers: rows sit at arbitrary places of the row pools, the host side of a transfer is one tightly packed
// staging buffer. One warp copies one row (16-byte vectors when both ends allow it).
struct RowCopy {
    char *dev;          // the row inside a pool
    long long stageOff; // its place in the staging buffer
    int len, pad;
};
__global__ void rowTransferKernel(const RowCopy *list, int n, char *stage, int toDevice) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const RowCopy rc = list[w];
    char *a = rc.dev, *b = stage + rc.stageOff;
    char *dst = toDevice ? a : b;
    const char *src = toDevice ? b : a;
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        const int n16 = rc.len >> 4;
        for (int k = lane; k < n16; k += 32) reinterpret_cast<uint4 *>(dst)[k] = reinterpret_cast<const uint4 *>(src)[k];
        for (int k = (n16 << 4) + lane; k < rc.len; k += 32) dst[k] = src[k];
    } else {
        for (int k = lane; k < rc.len; k += 32) dst[k] = src[k];
    }
}

// block-wide exclusive scan of one int per thread (kLvlThreads threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ int blockExclusiveScan(int v, int *warpSums, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += n;
    }
    if (lane == 31) warpSums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < kLvlThreads / 32) ? warpSums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += n;
        }
        if (lane < kLvlThreads / 32) warpSums[lane] = w;
    }
    __syncthreads();
    const int base = warp ? warpSums[warp - 1] : 0;
    *total = warpSums[kLvlThreads / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

// ---------------------------------------------------------------------------------------------------------------
// removeGappyColumns + calculatePSGP + packing. One CTA per side. Pass A counts the kept columns (the packed nucleotide
// layout needs the final length up front), pass B scans, writes the kept columns with their position-specific gap
// penalties in the DP layout and emits the (start,len) list of removed runs for addGappyColumnsBack.
// ---------------------------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(kLvlThreads) gappyCompactKernel(DevSide *sides, const float *raw, float *prof, int *runs, DevPair *pairs,
                                                                  float threshold, float gapOpen, float gapExtend) {
    __shared__ int warpSums[kLvlThreads / 32];
    __shared__ int sTotal;
    DevSide &sdRef = sides[blockIdx.x];
    const DevSide sd = sdRef;
    const float *col = raw + sd.rawOff;
    const bool enabled = (threshold != 1.0f);                               // helper.cpp:77
    const float numF = static_cast<float>(sd.alnNum);

    int kept = 0;
    for (int t = threadIdx.x; t < sd.alnLen; t += kLvlThreads) {
        const bool gappy = enabled && (__fdiv_rn(col[static_cast<long long>(t) * P + P - 1], numF) > threshold);   // helper.cpp:84
        kept += gappy ? 0 : 1;
    }
    int total;
    blockExclusiveScan(kept, warpSums, &total);
    const int newLen = total;
    const int n4 = (newLen + 3) / 4;
    float *out = prof + sd.profOff;

    const float scale = (P == 6) ? 0.5f : 1.0f;                             // helper.cpp:179
    const float minExtend = static_cast<float>(static_cast<double>(gapExtend) * 0.2);   // helper.cpp:180-181
    const float minOpen = static_cast<float>(static_cast<double>(gapOpen) * 0.1);
    const float openScaled = __fmul_rn(gapOpen, scale);

    int keptBase = 0, runBase = 0;
    bool oneHot = true;   // every kept column of this thread is one-hot (count 1.0 on one of the first P-1... letters, no gap)
    for (int c0 = 0; c0 < sd.alnLen; c0 += kLvlThreads) {
        const int t = c0 + threadIdx.x;
        const bool live = t < sd.alnLen;
        bool gappy = false, prevGappy = false, nextGappy = false;
        float v[P];
        if (live) {
#pragma unroll
            for (int x = 0; x < P; ++x) v[x] = col[static_cast<long long>(t) * P + x];
            if (enabled) {
                gappy = __fdiv_rn(v[P - 1], numF) > threshold;
                prevGappy = (t > 0) && (__fdiv_rn(col[static_cast<long long>(t - 1) * P + P - 1], numF) > threshold);
                nextGappy = (t + 1 < sd.alnLen) && (__fdiv_rn(col[static_cast<long long>(t + 1) * P + P - 1], numF) > threshold);
            }
        }
        const int isKept = (live && !gappy) ? 1 : 0;
        const int isStart = (live && gappy && !prevGappy) ? 1 : 0;
        int chunkKept, chunkStarts;
        const int idx = keptBase + blockExclusiveScan(isKept, warpSums, &chunkKept);
        const int runIdxIncl = runBase + blockExclusiveScan(isStart, warpSums, &chunkStarts) + isStart;
        if (isKept) {
            if (P == 6) {
                int ones = 0, zeros = 0;
#pragma unroll
                for (int x = 0; x < 5; ++x) { ones += (v[x] == 1.0f); zeros += (v[x] == 0.0f); }
                oneHot = oneHot && (ones == 1) && (zeros == 4) && (v[5] == 0.0f);
            }
            // calculatePSGP, helper.cpp:185-196 (the ratio is evaluated in double, as upstream)
            const float g = v[P - 1];
            float gOp = gapOpen, gEx = gapExtend;
            if (g > 0) {
                const double keep = static_cast<double>(__fsub_rn(numF, g)) * 1.0 / static_cast<double>(sd.alnNum);
                gOp = fminf(minOpen, static_cast<float>(static_cast<double>(openScaled) * keep));
                gEx = fminf(minExtend, static_cast<float>(static_cast<double>(gapExtend) * keep));
            }
            if (P == 6) {
                float4 *x = reinterpret_cast<float4 *>(out) + ntColIndex(idx, n4);
                float4 *y = x + 4 * static_cast<long long>(n4);
                *x = make_float4(v[0], v[1], v[2], v[3]);
                *y = make_float4(v[4], v[5], gOp, gEx);
            } else {
                float *d = out + static_cast<long long>(idx) * (P + 2);
#pragma unroll
                for (int x = 0; x < P; ++x) d[x] = v[x];
                d[P] = gOp; d[P + 1] = gEx;
            }
        }
        if (live && gappy) {
            int *r = runs + sd.runsOff;
            if (isStart) r[2 * (runIdxIncl - 1)] = t;
            if (!nextGappy) r[2 * (runIdxIncl - 1) + 1] = t;   // run end; turned into a length below
        }
        keptBase += chunkKept;
        runBase += chunkStarts;
    }
    const int allOneHot = __syncthreads_and(oneHot ? 1 : 0);
    for (int g = threadIdx.x; g < runBase; g += kLvlThreads) {
        int *r = runs + sd.runsOff + 2 * g;
        r[1] = r[1] - r[0] + 1;
    }
    if (threadIdx.x == 0) {
        sdRef.newLen = newLen;
        sdRef.nRuns = runBase;
        DevPair &pr = pairs[sd.pairIdx];
        if (sd.isQry) { pr.qryLen = newLen; pr.qryN4 = n4; }
        else { pr.refLen = newLen; pr.refN4 = n4; }
        if (P == 6 && allOneHot && newLen > 0) atomicOr(&pr.pad, sd.isQry ? kQryOneHot : kRefOneHot);   // DP fast path, talco_wavefront.cu
    }
    (void)sTotal;
}

// Nucleotide version (P = 6) for the HBM roofline: the gappy test reads the compact gap-count array profileBuildNtKernel wrote
// (4 bytes per column instead of a strided pass over the raw profile), a thread owns four consecutive columns per 1024-column
// chunk (one pair of block scans per chunk instead of per 256 columns), the raw columns come in as six 16-byte loads.
// Same arithmetic, same outputs (DP layout, run list, newLen, one-hot flag) as gappyCompactKernel<6>.
__global__ void __launch_bounds__(kLvlThreads) gappyCompactNtKernel(DevSide *sides, const float *raw, const float *gapCount, float *prof, int *runs,
                                                                     DevPair *pairs, float threshold, float gapOpen, float gapExtend) {
    constexpr int P = 6, C = 4;
    __shared__ int warpSums[kLvlThreads / 32];
    DevSide &sdRef = sides[blockIdx.x];
    const DevSide sd = sdRef;
    const float *col = raw + sd.rawOff;
    const float *gc = gapCount + sd.consOff;
    const bool enabled = (threshold != 1.0f);                               // helper.cpp:77
    const float numF = static_cast<float>(sd.alnNum);
    // helper.cpp:84 tests g / num > threshold with a float division per column. The quotient is monotonic in g, so the test is
    // the same as g >= gMin with gMin the smallest float whose quotient exceeds the threshold: found once per side by stepping a
    // few ulps around threshold * num with the very division the reference does (gap counts are >= 0).
    float gMin;
    {
        float c = fmaxf(__fmul_rn(threshold, numF), 0.0f);
        for (int it = 0; it < 64 && c > 0.0f && __fdiv_rn(c, numF) > threshold; ++it) c = __uint_as_float(__float_as_uint(c) - 1u);   // down to a failing value (or 0)
        for (int it = 0; it < 128 && !(__fdiv_rn(c, numF) > threshold); ++it) c = __uint_as_float(__float_as_uint(c) + 1u);            // up to the first passing one
        gMin = c;
        if (!(__fdiv_rn(c, numF) > threshold) || (c > 0.0f && __fdiv_rn(__uint_as_float(__float_as_uint(c) - 1u), numF) > threshold)) gMin = -1.0f;   // not bracketed: divide per column
    }
    auto isGappy = [&](float g) { return enabled && ((gMin >= 0.0f) ? (g >= gMin) : (__fdiv_rn(g, numF) > threshold)); };

    int kept = 0;
    for (int t = threadIdx.x; t < sd.alnLen; t += kLvlThreads) kept += isGappy(gc[t]) ? 0 : 1;
    int total;
    blockExclusiveScan(kept, warpSums, &total);
    const int newLen = total;
    const int n4 = (newLen + 3) / 4;
    float4 *outX = reinterpret_cast<float4 *>(prof + sd.profOff);
    float4 *outY = outX + 4 * static_cast<long long>(n4);
    int *r = runs + sd.runsOff;

    const float minExtend = static_cast<float>(static_cast<double>(gapExtend) * 0.2);   // helper.cpp:180-181
    const float minOpen = static_cast<float>(static_cast<double>(gapOpen) * 0.1);
    const float openScaled = __fmul_rn(gapOpen, 0.5f);                      // helper.cpp:179 (nucleotide scale)

    int keptBase = 0, runBase = 0;
    bool oneHot = true;
    for (int c0 = 0; c0 < sd.alnLen; c0 += kLvlThreads * C) {
        const int t0 = c0 + threadIdx.x * C;
        const int nHere = max(0, min(C, sd.alnLen - t0));
        float g[C];
        bool gp[C];
        if (nHere == C) {
            const float4 g4 = *reinterpret_cast<const float4 *>(gc + t0);
            g[0] = g4.x; g[1] = g4.y; g[2] = g4.z; g[3] = g4.w;
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c) g[c] = (c < nHere) ? gc[t0 + c] : 0.0f;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) gp[c] = (c < nHere) && isGappy(g[c]);
        const bool prevG = (nHere > 0 && t0 > 0) && isGappy(gc[t0 - 1]);
        const bool nextG = (t0 + C < sd.alnLen) && isGappy(gc[t0 + C]);
        int nKept = 0, nStart = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            nKept += (c < nHere && !gp[c]) ? 1 : 0;
            nStart += (gp[c] && !(c ? gp[c - 1] : prevG)) ? 1 : 0;
        }
        int chunkKept, chunkStarts;
        int idx = keptBase + blockExclusiveScan(nKept, warpSums, &chunkKept);
        int runIdx = runBase + blockExclusiveScan(nStart, warpSums, &chunkStarts);   // run starts before this thread's columns
        float v[C][P];
        if (nHere == C) {
            const float4 *src = reinterpret_cast<const float4 *>(col + static_cast<long long>(t0) * P);
            float *flat = &v[0][0];
#pragma unroll
            for (int x = 0; x < 6; ++x) { const float4 q = src[x]; flat[4 * x] = q.x; flat[4 * x + 1] = q.y; flat[4 * x + 2] = q.z; flat[4 * x + 3] = q.w; }
        } else {
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int x = 0; x < P; ++x) v[c][x] = (c < nHere) ? col[static_cast<long long>(t0 + c) * P + x] : 0.0f;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            if (c >= nHere) continue;
            const int t = t0 + c;
            if (!gp[c]) {
                int ones = 0, zeros = 0;
#pragma unroll
                for (int x = 0; x < 5; ++x) { ones += (v[c][x] == 1.0f); zeros += (v[c][x] == 0.0f); }
                oneHot = oneHot && (ones == 1) && (zeros == 4) && (v[c][5] == 0.0f);
                // calculatePSGP, helper.cpp:185-196 (the ratio is evaluated in double, as upstream)
                const float gg = v[c][P - 1];
                float gOp = gapOpen, gEx = gapExtend;
                if (gg > 0) {
                    const double keep = static_cast<double>(__fsub_rn(numF, gg)) * 1.0 / static_cast<double>(sd.alnNum);
                    gOp = fminf(minOpen, static_cast<float>(static_cast<double>(openScaled) * keep));
                    gEx = fminf(minExtend, static_cast<float>(static_cast<double>(gapExtend) * keep));
                }
                const long long at = ntColIndex(idx, n4);
                outX[at] = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
                outY[at] = make_float4(v[c][4], v[c][5], gOp, gEx);
                ++idx;
            } else {
                const bool pg = c ? gp[c - 1] : prevG;
                const bool ng = (c + 1 < C) ? ((c + 1 < nHere) ? gp[c + 1] : false) : nextG;
                if (!pg) { r[2 * runIdx] = t; ++runIdx; }
                if (!ng) r[2 * (runIdx - 1) + 1] = t;                       // run end; turned into a length below
            }
        }
        keptBase += chunkKept;
        runBase += chunkStarts;
    }
    const int allOneHot = __syncthreads_and(oneHot ? 1 : 0);
    for (int q = threadIdx.x; q < runBase; q += kLvlThreads) {
        int *rr = r + 2 * q;
        rr[1] = rr[1] - rr[0] + 1;
    }
    if (threadIdx.x == 0) {
        sdRef.newLen = newLen;
        sdRef.nRuns = runBase;
        DevPair &pr = pairs[sd.pairIdx];
        if (sd.isQry) { pr.qryLen = newLen; pr.qryN4 = n4; }
        else { pr.refLen = newLen; pr.refN4 = n4; }
        if (allOneHot && newLen > 0) atomicOr(&pr.pad, sd.isQry ? kQryOneHot : kRefOneHot);   // DP fast path, talco_wavefront.cu
    }
}

// ---------------------------------------------------------------------------------------------------------------
// addGappyColumnsBack + pairwiseGlobal (src/alignment-helper.cpp:324-375, 243-322). The merge of the removed-column runs
// back into the DP path is a strictly sequential walk over the path (a run goes in front of the first op at which the
// original coordinate reaches the run's start; runs hit on both sides at the same op are aligned against each other by a
// small affine-gap global alignment of the two consensus substrings). One warp per pair, lane 0 walks; the walk costs
// ~0.1-0.2 ms per pair and thousands of pairs walk concurrently, so the phase is invisible next to the DP.
// The consensus alignment keeps its matrices in shared memory; a pair with a run pair too large for them is redone by the
// LARGE instantiation of the same kernel (matrices in global scratch), see below.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kRestoreTbCells = 8192;       // (m+1)*(n+1) limit of the in-kernel consensus alignment
constexpr int kRestoreRowCap = 1024;        // n+1 limit

// pairwiseGlobal (alignment-helper.cpp:243-322) of consensus substrings s1[0,m) x s2[0,n) by one warp, row by row: the
// match and vertical-gap states of a row depend on the previous row only (lanes split the columns), the horizontal-gap
// state is a running maximum along the row that must be accumulated in the reference's order (lane 0, ~2 dependent
// instructions per cell). Every cell evaluates exactly the reference's expressions. Writes the ops to out[0, len) and
// returns len (same value in every lane).
__device__ __forceinline__ int consensusAlignWarp(int lane, const char *s1, int m, const char *s2, int n, const float *sScore, int M, int isProtein,
                                                  const signed char *aaLut, float gapOpen, float gapExtend, int8_t *tb,
                                                  float *rM, float *rX, float *rY, int rowStride, unsigned char *idx2, int8_t *out) {
    const size_t W = static_cast<size_t>(n) + 1;
    const float big = -1e9f;
    for (int t = lane; t < n; t += 32) { const unsigned char c = static_cast<unsigned char>(s2[t]); idx2[t] = static_cast<unsigned char>(isProtein ? letterIndexAa(c, aaLut) : letterIndexNt(c)); }
    for (int j = lane; j <= n; j += 32) { rM[j] = 0.0f; rX[j] = (j > 0) ? big : 0.0f; rY[j] = 0.0f; }
    __syncwarp();
    for (int i = 1; i <= m; ++i) {
        float *cM = rM + (i & 1) * rowStride, *cX = rX + (i & 1) * rowStride, *cY = rY + (i & 1) * rowStride;
        const float *pM = rM + ((i & 1) ^ 1) * rowStride, *pX = rX + ((i & 1) ^ 1) * rowStride, *pY = rY + ((i & 1) ^ 1) * rowStride;
        const unsigned char c1 = static_cast<unsigned char>(s1[i - 1]);
        const float *srow = sScore + (isProtein ? letterIndexAa(c1, aaLut) : letterIndexNt(c1)) * M;
        for (int j = 1 + lane; j <= n; j += 32) {
            cM[j] = __fadd_rn(srow[idx2[j - 1]], fmaxf(fmaxf(pM[j - 1], pX[j - 1]), pY[j - 1]));
            cX[j] = fmaxf(__fadd_rn(pM[j], gapOpen), __fadd_rn(pX[j], gapExtend));
        }
        if (lane == 0) { cM[0] = 0.0f; cX[0] = 0.0f; cY[0] = big; }
        __syncwarp();
        if (lane == 0) {
            float y = big;
#pragma unroll 4
            for (int j = 1; j <= n; ++j) {
                y = fmaxf(__fadd_rn(cM[j - 1], gapOpen), __fadd_rn(y, gapExtend));
                cY[j] = y;
            }
        }
        __syncwarp();
        for (int j = 1 + lane; j <= n; j += 32) {
            const float vm = cM[j], vx = cX[j], vy = cY[j];
            const float best = fmaxf(fmaxf(vm, vx), vy);
            tb[i * W + j] = (best == vm) ? 0 : ((best == vy) ? 1 : 2);
        }
    }
    __syncwarp();
    int len = 0;
    if (lane == 0) {
        for (int i = m, j = n; i > 0 || j > 0; ++len) {
            const int d = (i == 0) ? 1 : ((j == 0) ? 2 : tb[i * W + j]);
            if (d == 0) { --i; --j; } else if (d == 1) { --j; } else { --i; }
        }
        int at = len;
        for (int i = m, j = n; i > 0 || j > 0;) {
            const int d = (i == 0) ? 1 : ((j == 0) ? 2 : tb[i * W + j]);
            out[--at] = static_cast<int8_t>(d);
            if (d == 0) { --i; --j; } else if (d == 1) { --j; } else { --i; }
        }
    }
    __syncwarp();
    return __shfl_sync(0xffffffffu, len, 0);
}

// LARGE = false: the consensus alignment lives in shared memory; a coinciding run pair that does not fit makes the kernel
// finish the walk without writing (it only records the largest matrix and row it would need in `need`) and the pair is
// redone by the LARGE = true instantiation, whose matrices live in a per-block slice of a global scratch buffer sized by
// the host from `need`. Same code, same bits; only where the scratch is differs.
// A consensus alignment the walk has put off: the walk reserves m + n bytes of the final path for it (an alignment of m against n
// columns is at most that long), fills them with the skip code 3 and moves on; consensusJobsKernel aligns all put-off run pairs of the
// level in parallel and writes the ops at the start of the reserved bytes; pathCompactKernel squeezes the skip codes out. The walk
// itself never needs the alignment's result (it advances by m reference and n query columns whatever the alignment looks like), so
// the only sequential part left per pair is a walk of ~100 cycles per run.
struct RestoreJob {
    long long outOff;       // into finalPaths
    long long refOff, qryOff;   // into cons
    int m, n;
};
constexpr int8_t kOpSkip = 3;

template <bool LARGE>
__global__ void __launch_bounds__(32) gappyRestoreKernel(DevUpdate *ups, const int *updPair, const int *which, int nu, const DevPair *pairs,
                                                         DevResult *results, const DevSide *sides, const int *runs, const char *cons, int8_t *pathsWo,
                                                         int8_t *finalPaths, const float *score, int M, int isProtein, const signed char *aaLut,
                                                         float gapOpen, float gapExtend, long long *need, char *largeScratch,
                                                         long long largeCells, int largeCols, RestoreJob *jobs, int *jobCount, int jobCap) {
    constexpr int kTb = LARGE ? 16 : kRestoreTbCells, kRow = LARGE ? 4 : kRestoreRowCap;
    __shared__ int8_t sTb[kTb];
    __shared__ float sM[2 * kRow], sX[2 * kRow], sY[2 * kRow];
    __shared__ unsigned char sIdx2[kRow];
    // scratch of the consensus alignment: shared memory, or this block's slice of the global buffer
    // (slice layout: tb[largeCells] | M,X,Y [2][largeCols] floats each | idx2[largeCols])
    const long long cellCap = LARGE ? largeCells : kRestoreTbCells;
    const int colCap = LARGE ? largeCols : kRestoreRowCap;
    char *slice = LARGE ? largeScratch + static_cast<size_t>(blockIdx.x) * (static_cast<size_t>(largeCells) + 25ull * largeCols + 64) : nullptr;
    int8_t *tb = LARGE ? reinterpret_cast<int8_t *>(slice) : sTb;
    float *dM = LARGE ? reinterpret_cast<float *>(slice + ((largeCells + 15) & ~15ll)) : sM;
    float *dX = LARGE ? dM + 2 * static_cast<size_t>(largeCols) : sX;
    float *dY = LARGE ? dX + 2 * static_cast<size_t>(largeCols) : sY;
    unsigned char *idx2 = LARGE ? reinterpret_cast<unsigned char *>(dY + 2 * static_cast<size_t>(largeCols)) : sIdx2;
    // staged windows of the two run lists and of the path: the walk is a chain of dependent reads, which must not each
    // pay a trip to L2
    constexpr int kRunWin = 256, kOpWin = 1024;
    __shared__ int sRunR[2 * kRunWin], sRunQ[2 * kRunWin];
    __shared__ __align__(16) int8_t sOps[kOpWin];
    __shared__ float sScore[21 * 21];
    const int lane = threadIdx.x;
    for (int t = lane; t < M * M; t += 32) sScore[t] = score[t];
    __syncwarp();
    // every lane keeps the (uniform) walk state; lane-parallel parts: op copies, run fills, the consensus alignment
    for (int kk = blockIdx.x; kk < nu; kk += gridDim.x) {
        const int k = which ? which[kk] : kk;               // entry of the update list (the LARGE pass lists the pairs to redo)
        const int p = updPair[k];
        DevResult res = results[p];
        const DevPair pr = pairs[p];
        const DevSide sr = sides[2 * p], sq = sides[2 * p + 1];
        int8_t *aln = pathsWo + pr.alnOff;
        __syncwarp();
        if (lane == 0) { need[2 * k] = 0; need[2 * k + 1] = 0; }
        long long wantCells = 0;
        int wantCols = 0;
        if (res.status == kStatusEmptySide) {               // alignment-cpu.cpp:89-90: the other side's columns against nothing
            const int n = (pr.refLen < 1) ? max(pr.qryLen, 0) : max(pr.refLen, 0);
            const int8_t op = (pr.refLen < 1) ? 1 : 2;
            for (int a = lane; a < n; a += 32) aln[a] = op;
            res.status = 0; res.pathLen = n;
            if (lane == 0) { results[p].status = 0; results[p].pathLen = n; }
            __syncwarp();
        }
        if (res.status != 0) { if (lane == 0) ups[k].pathLen = 0; continue; }
        const int alnLen = res.pathLen;
        const int *runsR = runs + sr.runsOff, *runsQ = runs + sq.runsOff;
        const int nR = sr.nRuns, nQ = sq.nRuns;
        const char *consR = cons + sr.consOff, *consQ = cons + sq.consOff;
        int8_t *out = finalPaths + ups[k].pathOff;
        int baseR = 0, baseQ = 0, baseA = 0;
        auto stageRuns = [&](const int *src, int *dst, int base, int total) {
            __syncwarp();
            const int cnt = 2 * min(kRunWin, total - base);
            for (int t = lane; t < cnt; t += 32) dst[t] = src[2 * base + t];
            __syncwarp();
        };
        auto stageOps = [&](int base) {          // base is a multiple of 16; the path buffer is padded to 16 bytes
            __syncwarp();
            const int cnt = min(kOpWin, ((alnLen - base) + 15) & ~15);
            for (int t = lane * 16; t < cnt; t += 32 * 16)
                *reinterpret_cast<uint4 *>(sOps + t) = *reinterpret_cast<const uint4 *>(aln + base + t);
            __syncwarp();
        };
        stageRuns(runsR, sRunR, 0, nR);
        stageRuns(runsQ, sRunQ, 0, nQ);
        stageOps(0);
        int w = 0, r = 0, q = 0, gr = 0, gq = 0, a = 0;
        int nextR = (nR > 0) ? sRunR[0] : -1, nextQ = (nQ > 0) ? sRunQ[0] : -1;
        bool giveUp = false, holed = false;
        for (;;) {
            // runs that start at the current original coordinates go in front of op a (helper.cpp:338-362)
            const bool hitR = (r == nextR), hitQ = (q == nextQ);
            if (hitR && hitQ) {
                const int m = sRunR[2 * (gr - baseR) + 1], n = sRunQ[2 * (gq - baseQ) + 1];
                const long long cells = static_cast<long long>(m + 1) * (n + 1);
                if (cells > cellCap || n + 1 > colCap) {     // does not fit: finish the walk dry to learn the largest need
                    giveUp = true;
                    wantCells = max(wantCells, cells);
                    wantCols = max(wantCols, n + 1);
                }
                if (!giveUp) {
                    int slot = jobCap;
                    if (!LARGE && jobs != nullptr) {
                        if (lane == 0) slot = atomicAdd(jobCount, 1);
                        slot = __shfl_sync(0xffffffffu, slot, 0);
                    }
                    if (slot < jobCap) {       // put off: reserve m + n bytes, aligned later by consensusJobsKernel
                        if (lane == 0) {
                            RestoreJob jb;
                            jb.outOff = ups[k].pathOff + w; jb.refOff = sr.consOff + r; jb.qryOff = sq.consOff + q; jb.m = m; jb.n = n;
                            jobs[slot] = jb;
                        }
                        for (int t = lane; t < m + n; t += 32) out[w + t] = kOpSkip;
                        w += m + n;
                        holed = true;
                    } else {
                        w += consensusAlignWarp(lane, consR + r, m, consQ + q, n, sScore, M, isProtein, aaLut, gapOpen, gapExtend, tb, dM, dX, dY,
                                                colCap, idx2, out + w);
                    }
                }
                r += m; q += n;
            } else {
                if (hitR) {
                    const int len = sRunR[2 * (gr - baseR) + 1];
                    if (!giveUp) for (int t = lane; t < len; t += 32) out[w + t] = 2;
                    w += len; r += len;
                }
                if (hitQ) {
                    const int len = sRunQ[2 * (gq - baseQ) + 1];
                    if (!giveUp) for (int t = lane; t < len; t += 32) out[w + t] = 1;
                    w += len; q += len;
                }
            }
            if (hitR) {
                ++gr;
                if (gr < nR && gr >= baseR + kRunWin) { baseR = gr; stageRuns(runsR, sRunR, baseR, nR); }
                nextR = (gr < nR) ? sRunR[2 * (gr - baseR)] : -1;
            }
            if (hitQ) {
                ++gq;
                if (gq < nQ && gq >= baseQ + kRunWin) { baseQ = gq; stageRuns(runsQ, sRunQ, baseQ, nQ); }
                nextQ = (gq < nQ) ? sRunQ[2 * (gq - baseQ)] : -1;
            }
            if (a >= alnLen) break;
            // copy ops a, a+1, ... up to (not including) the first later op in front of which a run starts
            const int left = min(32, alnLen - a);
            if (a + left > baseA + kOpWin) { baseA = a & ~15; stageOps(baseA); }
            const int op = (lane < left) ? sOps[a - baseA + lane] : 3;
            // a run goes in front of the first op at which the number of consumed columns reaches its start
            const unsigned maskR = __ballot_sync(0xffffffffu, op == 0 || op == 2);
            const unsigned maskQ = __ballot_sync(0xffffffffu, op == 0 || op == 1);
            const unsigned below = (1u << lane) - 1u;
            const bool hitHere = (lane > 0) && (lane < left) && ((r + __popc(maskR & below) == nextR) || (q + __popc(maskQ & below) == nextQ));
            const unsigned hits = __ballot_sync(0xffffffffu, hitHere);
            const int take = hits ? (__ffs(hits) - 1) : left;      // >= 1
            if (lane < take && !giveUp) out[w + lane] = static_cast<int8_t>(op);
            const unsigned upto = (take >= 32) ? 0xffffffffu : ((1u << take) - 1u);
            r += __popc(maskR & upto);
            q += __popc(maskQ & upto);
            w += take; a += take;
        }
        __syncwarp();
        if (lane == 0) {
            if (giveUp) { need[2 * k] = wantCells; need[2 * k + 1] = wantCols; ups[k].pathLen = 0; ups[k].pad = 0; }
            else { ups[k].pathLen = w; ups[k].pad = holed ? 1 : 0; }
        }
    }
}

// The consensus alignments the walks put off, one warp per job (same routine, same bits as the in-line path).
__global__ void __launch_bounds__(32) consensusJobsKernel(const RestoreJob *jobs, const int *jobCount, int jobCap, const char *cons, int8_t *finalPaths,
                                                          const float *score, int M, int isProtein, const signed char *aaLut, float gapOpen, float gapExtend) {
    __shared__ int8_t sTb[kRestoreTbCells];
    __shared__ float sM[2 * kRestoreRowCap], sX[2 * kRestoreRowCap], sY[2 * kRestoreRowCap];
    __shared__ unsigned char sIdx2[kRestoreRowCap];
    __shared__ float sScore[21 * 21];
    const int lane = threadIdx.x;
    for (int t = lane; t < M * M; t += 32) sScore[t] = score[t];
    __syncwarp();
    const int nJobs = min(*jobCount, jobCap);
    for (int j = blockIdx.x; j < nJobs; j += gridDim.x) {
        const RestoreJob jb = jobs[j];
        consensusAlignWarp(lane, cons + jb.refOff, jb.m, cons + jb.qryOff, jb.n, sScore, M, isProtein, aaLut, gapOpen, gapExtend, sTb, sM, sX, sY,
                           kRestoreRowCap, sIdx2, finalPaths + jb.outOff);
        __syncwarp();
    }
}

// Squeezes the skip codes out of the final paths that contain reserved bytes (DevUpdate::pad set by the walk), in place: one block per
// pair walks the path in 1024-byte pieces; a piece is read completely before anything is written, and what is written lies at or before it.
__global__ void __launch_bounds__(kLvlThreads) pathCompactKernel(DevUpdate *ups, int nu, int8_t *finalPaths) {
    __shared__ int warpSums[kLvlThreads / 32];
    const int k = blockIdx.x;
    if (k >= nu || ups[k].pad == 0) return;
    int8_t *path = finalPaths + ups[k].pathOff;
    const int len = ups[k].pathLen;
    int base = 0;
    for (int c0 = 0; c0 < len; c0 += 4 * kLvlThreads) {
        const int at = c0 + 4 * threadIdx.x;
        int8_t v[4];
        int keep = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            v[e] = (at + e < len) ? path[at + e] : kOpSkip;
            keep += (v[e] != kOpSkip);
        }
        int total;
        int dst = base + blockExclusiveScan(keep, warpSums, &total);      // barriers inside: every thread has read its bytes before any write below
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (v[e] != kOpSkip) path[dst++] = v[e];
        base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) { ups[k].pathLen = base; ups[k].pad = 0; }
}

// ---------------------------------------------------------------------------------------------------------------
// Per-chunk prefix counts of a final path: chunk c gets (#ref-consuming ops, #qry-consuming ops) in path[0, c*1024).
// One warp per pair.
// ---------------------------------------------------------------------------------------------------------------
__global__ void pathChunkKernel(const DevUpdate *ups, int nPairs, const int8_t *paths, int *chunkCounts) {
    const int pair = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pair >= nPairs) return;
    const DevUpdate u = ups[pair];
    const int8_t *path = paths + u.pathOff;
    int *out = chunkCounts + u.chunkOff;
    int accR = 0, accQ = 0;
    const int nChunks = (u.pathLen + kPathChunk - 1) / kPathChunk;
    for (int c = 0; c < nChunks; ++c) {
        if (lane == 0) { out[2 * c] = accR; out[2 * c + 1] = accQ; }
        int r = 0, q = 0;
        const int end = min(u.pathLen, (c + 1) * kPathChunk);
        for (int k = c * kPathChunk + lane; k < end; k += 32) {
            const int op = path[k];
            r += (op == 0 || op == 2);
            q += (op == 0 || op == 1);
        }
        accR += __reduce_add_sync(0xffffffffu, r);
        accQ += __reduce_add_sync(0xffffffffu, q);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// updateAlignment (helper.cpp:381-401, 428-448) and updateFrequency (helper.cpp:506-539). grid = (pair, chunk).
// The block scans its 1024 path ops once and reuses the source indices for every member row of the pair.
// ---------------------------------------------------------------------------------------------------------------
template <int P>
__global__ void __launch_bounds__(kLvlThreads) rowUpdateKernel(const DevUpdate *ups, const int8_t *paths, const int *chunkCounts,
                                                               const char *const *rowIn, char *const *rowOut, const float *freq, float *merged) {
    __shared__ int warpSums[kLvlThreads / 32];
    __shared__ int srcR[kPathChunk], srcQ[kPathChunk];
    __shared__ int8_t ops[kPathChunk];
    const DevUpdate u = ups[blockIdx.x];          // grid.x = pair, grid.y = chunk of kPathChunk path ops
    const int k0 = blockIdx.y * kPathChunk;
    if (k0 >= u.pathLen) return;
    const int8_t *path = paths + u.pathOff;
    const int baseR = chunkCounts[u.chunkOff + 2 * blockIdx.y], baseQ = chunkCounts[u.chunkOff + 2 * blockIdx.y + 1];
    constexpr int PER = kPathChunk / kLvlThreads;   // 4 consecutive ops per thread
    int r[PER], q[PER], sumR = 0, sumQ = 0;
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        const int k = k0 + threadIdx.x * PER + e;
        const int op = (k < u.pathLen) ? path[k] : 3;
        ops[threadIdx.x * PER + e] = static_cast<int8_t>(op);
        r[e] = (op == 0 || op == 2); q[e] = (op == 0 || op == 1);
        sumR += r[e]; sumQ += q[e];
    }
    int tot;
    int exR = baseR + blockExclusiveScan(sumR, warpSums, &tot);
    int exQ = baseQ + blockExclusiveScan(sumQ, warpSums, &tot);
#pragma unroll
    for (int e = 0; e < PER; ++e) {
        srcR[threadIdx.x * PER + e] = exR; exR += r[e];
        srcQ[threadIdx.x * PER + e] = exQ; exQ += q[e];
    }
    __syncthreads();
    const int nHere = min(kPathChunk, u.pathLen - k0);
    // member rows (grid.z strides over them): ref members copy on 0/2, qry members on 0/1, '-' otherwise. A thread owns the four
    // consecutive ops it scanned: the bytes it takes from a row are consecutive in the source (starting at its exclusive source
    // index), so per row it reads the two aligned words that hold them (rows are 16-byte aligned and over-allocated), picks the
    // bytes with one byte permutation whose selector depends on the ops only, and writes one aligned 32-bit word.
    {
        constexpr int WORDS = PER / 4;
        const int e0 = threadIdx.x * PER;
        unsigned selR[WORDS], selQ[WORDS];     // per output byte: index 0..3 into the word's source window, or 4 = '-'
        int firstR[WORDS], firstQ[WORDS];
#pragma unroll
        for (int w = 0; w < WORDS; ++w) {
            unsigned sr = 0, sq = 0;
            int pr = 0, pq = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int op = ops[e0 + 4 * w + e];
                const bool takeR = (op == 0 || op == 2), takeQ = (op == 0 || op == 1);
                sr |= static_cast<unsigned>(takeR ? pr : 4) << (4 * e);
                sq |= static_cast<unsigned>(takeQ ? pq : 4) << (4 * e);
                pr += takeR; pq += takeQ;
            }
            selR[w] = sr; selQ[w] = sq;
            firstR[w] = srcR[e0 + 4 * w]; firstQ[w] = srcQ[e0 + 4 * w];
        }
        for (int m = blockIdx.z; m < u.nRef + u.nQry; m += gridDim.z) {
            const bool isRef = m < u.nRef;
            const char *in = rowIn[u.memberOff + m];
            char *out = rowOut[u.memberOff + m] + k0;
#pragma unroll
            for (int w = 0; w < WORDS; ++w) {
                const int eW = e0 + 4 * w;
                if (eW + 4 <= nHere) {
                    const int first = isRef ? firstR[w] : firstQ[w];
                    const unsigned *src = reinterpret_cast<const unsigned *>(in) + (first >> 2);
                    const unsigned window = __funnelshift_r(src[0], src[1], 8 * (first & 3));
                    *reinterpret_cast<unsigned *>(out + eW) = __byte_perm(window, 0x2D2D2D2Du, isRef ? selR[w] : selQ[w]);
                } else {
                    for (int e = eW; e < min(eW + 4, nHere); ++e) {
                        const int op = ops[e];
                        const bool take = (op == 0) || (op == (isRef ? 2 : 1));
                        out[e] = take ? in[isRef ? srcR[e] : srcQ[e]] : '-';
                    }
                }
            }
        }
    }
    if (u.mergedOff >= 0 && blockIdx.z == 0) {                              // updateFrequency, helper.cpp:513-531
        const float *fr = freq + u.freqRefOff, *fq = freq + u.freqQryOff;
        float *mg = merged + u.mergedOff + static_cast<long long>(k0) * P;
        for (int x = threadIdx.x; x < nHere * P; x += kLvlThreads) {
            const int e = x / P, v = x - e * P;
            const int op = ops[e];
            float val;
            if (op == 0) val = __fadd_rn(fr[static_cast<long long>(srcR[e]) * P + v], fq[static_cast<long long>(srcQ[e]) * P + v]);
            else if (op == 1) { val = fq[static_cast<long long>(srcQ[e]) * P + v]; if (v == P - 1) val = __fadd_rn(val, u.refWeight); }
            else if (op == 2) { val = fr[static_cast<long long>(srcR[e]) * P + v]; if (v == P - 1) val = __fadd_rn(val, u.qryWeight); }
            else val = 0.0f;
            mg[x] = val;
        }
    }
}

} // namespace twl

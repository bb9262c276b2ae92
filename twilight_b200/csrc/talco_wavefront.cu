// talco_wavefront.cu — the fast TALCO-XDrop kernel for nucleotide profiles (P = 6): one pair per CTA, the anti-diagonal
// wavefront lives in REGISTERS.
//
// Mapping. A CTA of NT threads owns W = KS*NT consecutive query rows at a time (KS rows per thread; 128 x 4 is the
// throughput instantiation, 512 x 2 the wide / low-latency one). Row i is handled by "slot" rho = i mod W, i.e. thread
// rho/KS, register slot rho%KS (a thread owns KS consecutive rows); as the live band [L,U] of the X-drop recurrence
// slides along the query, a thread whose rows fell below L is re-assigned to rows + W (W need not be a power of two). Per
// slot the thread keeps, in registers, the query column (6 counts + 2 position-specific gap penalties) and the
// H / I / D scores of the previous two anti-diagonals, so per cell and per diagonal the only memory traffic is one
// 32-byte reference column (two LDG.128 that hit L1; the lines needed a few diagonals ahead are prefetched) and one
// traceback byte (the four slots of a thread store one coalesced 32-bit word). Row-neighbour values come from the
// thread's own registers for slots 1..3 and from one warp shuffle for slot 0; warps exchange their edge values and the
// per-diagonal reduction (running maximum for the X-drop rule, first / last live row: folded into one word per diagonal with
// shared-memory atomics by the warps that hold a surviving cell) through shared memory with ONE barrier per diagonal. The four cells of a thread are evaluated branch-free so that their dependent FP chains
// interleave; warps that hold no live cell skip the evaluation.
//
// What is kept bit-identical to the reference CPU path (src/TALCO-XDrop.cpp:233-689): the float operation order of
// the score (see talco_score.cuh; the "DNA3Z" fast path below only drops terms that are exact zeros), tie rules, the
// prune rule against the previous diagonal's maximum, dead-end trimming, the convergence pointers INCLUDING the
// reference's rotating buffers indexed by (row - L[k]) — those live in shared memory with the reference's indexing so
// that the stale slots the reference reads are reproduced — the tile stop rule, the traceback start cell and the
// per-tile path concatenation of Align_freq (:62-108).
//
// A band wider than W-(KS-1) cannot be held; the pair keeps its finished tiles and is handed to a wider instantiation
// (running at the same time, see TalcoArgs::coMode, or launched afterwards) or to the generic kernel (talco_generic.cu),
// which resume it at the failing tile.
#include "talco_nt.cuh"

namespace twl {


struct WaveShared {
    alignas(16) int red[4][4];       // per diagonal (ring of four): {max score as ordered int, first live row, last live row, -}, folded by the
                                     // warps that hold a surviving cell with shared-memory atomics; after the barrier every thread reads one word
    float2 edge[2][32];         // per warp: H and I of the warp's last slot (row-neighbour of the next warp's first slot)
    unsigned convMask[3];
    int8_t ops[2 * kMaxMarker + 16];
    int refOff, qryOff, lastTile, error, nOps, opsBegin, tailLen, tailOp;
    int work, fed;
};

// KS = rows per thread: 4 for throughput (W = 4*NT); 2 with NT = 256 for levels with few pairs, where a CTA runs alone on
// its SM and the per-thread instruction stream, not the issue rate, sets the time of a diagonal.
// SC = score source: 0 = nucleotide profiles, similarity computed on the fly (scoreDiagonal); 1 = any alphabet, similarity read from the
// pair's anti-diagonal-major matrix written by simMatrixKernel (talco_sim.cu) and gap penalties from its compact array — the
// protein path: the 21 x 21 contraction is done once per cell by a dependency-free kernel at full occupancy, and the recurrence
// runs register-resident here instead of in the shared-memory generic kernel.
template <int NT, int MC, int KS, int SC>
__global__ void __launch_bounds__(NT, (NT <= 128 ? (SC ? 1024 : 640) / NT : 1)) talcoWavefrontKernel(const TalcoArgs a) {
    constexpr int kSlots = KS;
    constexpr int W = NT * kSlots;
    constexpr int NW = NT / 32;
    constexpr int CW = W + 4;                       // convergence arrays, reference indexing (row - L[k]) plus padding
    __shared__ WaveShared sh;
    __shared__ int sCS[3][CW], sCI[2][CW], sCD[2][CW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *tb = a.tbScratch + static_cast<size_t>(blockIdx.x) * a.tbStride;   // tb[k][rho], row stride W
    const int marker = a.marker;
    const int rho0 = tid * kSlots;
    if (a.coMode == 1 && tid == 0) atomicAdd(a.heartbeat, 1);
    if (a.coMode == 2 && tid == 0 && a.arrived) atomicAdd(a.arrived, 1);

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            if (a.coMode == 2) {
                // wide worker: fed pairs first, then the main queue; leave when every main pair is finished
                const int nMain = *a.nWorkPtr;
                int got = -1, fed = 0;
                int lastBeat = *reinterpret_cast<volatile int *>(a.heartbeat);
                unsigned long long lastChange = globalTimerNs();
                for (;;) {
                    // Every main pair finished: nothing more will be handed over. What is still queued is better served by
                    // the clean-up launch that follows (one CTA per SM) than by the few wide workers one after the other.
                    if (*reinterpret_cast<volatile int *>(a.mainDone) >= nMain) break;
                    const int cur = *reinterpret_cast<volatile int *>(a.feedCursor);
                    if (cur < *reinterpret_cast<volatile int *>(a.feedCount)) {
                        if (atomicCAS(a.feedCursor, cur, cur + 1) == cur) {
                            int e;
                            while ((e = *reinterpret_cast<volatile int *>(a.feedList + cur)) < 0) __nanosleep(64);
                            __threadfence();
                            got = e; fed = 1;
                            if (a.coTrace) a.coTrace[4 * e + 1] = globalTimerNs();
                            break;
                        }
                        continue;
                    }
                    // main-queue work only while the narrow kernel is demonstrably running (its CTAs beat once when they start)
                    if (*reinterpret_cast<volatile int *>(a.heartbeat) != 0 && *reinterpret_cast<volatile int *>(a.queue) < a.coTakeBelow) {
                        const int w = atomicAdd(a.queue, 1);
                        if (w < nMain) { got = a.order[w]; break; }
                    }
                    // Watchdog: the producers bump the heartbeat every tile (a few ms at most). No beat for 20 ms means the narrow
                    // kernel is not running next to us (profiler replay, CUDA_LAUNCH_BLOCKING): leave; whatever is handed over later
                    // is picked up by the clean-up launch that follows both kernels.
                    const int beat = *reinterpret_cast<volatile int *>(a.heartbeat) + *reinterpret_cast<volatile int *>(a.mainDone);
                    const unsigned long long now = globalTimerNs();
                    if (beat != lastBeat) { lastBeat = beat; lastChange = now; }
                    else if (now - lastChange > 20000000ull) { if (a.watchdog) atomicExch(a.watchdog, 1); break; }
                    __nanosleep(256);
                }
                sh.work = got; sh.fed = fed;
            } else {
                const int w = atomicAdd(a.queue, 1);
                sh.work = (w < *a.nWorkPtr) ? a.order[w] : -1;
                sh.fed = 0;
            }
        }
        __syncthreads();
        const int pairIdx = sh.work;
        if (pairIdx < 0) break;
        const bool fedPair = sh.fed != 0;
        const DevPair pr = a.pairs[pairIdx];
        if (pr.refLen < 1 || pr.qryLen < 1) {   // an empty side (after gappy-column removal): nothing to align, the host emits the trivial path
            if (tid == 0) {
                DevResult res;
                res.status = kStatusEmptySide; res.pathLen = 0; res.tiles = 0; res.pad = 0; res.cells = 0; res.diagonals = 0; res.resRefOff = 0; res.resQryOff = 0;
                a.results[pairIdx] = res;
                if (a.coMode && !fedPair) { __threadfence(); atomicAdd(a.mainDone, 1); }
            }
            continue;
        }
        const float *refCols = a.prof + pr.refOff;
        const float *qryCols = a.prof + pr.qryOff;
        int8_t *path = a.paths + pr.alnOff;

        const float negInf = -static_cast<float>(2.0 * pr.xdrop + 1.0);
        const float xdropF = static_cast<float>(pr.xdrop);
        const float denom = __fmul_rn(pr.refNum, pr.qryNum);
        const float rcp = __fdiv_rn(1.0f, denom);
        // 0: denominator is 1 (no division), 1: reciprocal-based exact division, 2: IEEE divide (mantissa of all ones)
        const int divMode = (denom == 1.0f) ? 0 : (((__float_as_int(denom) & 0x7fffff) == 0x7fffff) ? 2 : 1);
        const float gapChar = pr.gapChar;
        const int kind = pr.pad;   // kRefOneHot | kQryOneHot
        const float *simBase = nullptr;
        const float2 *gapRef = nullptr;
        int simStride = 0;
        if (SC == 1) {
            const DevSim si = a.simInfo[pairIdx];
            simBase = a.sim + si.simOff; gapRef = reinterpret_cast<const float2 *>(a.sim + si.gapOff); simStride = si.stride;
        }
        int refOff = 0, qryOff = 0, tile = 0, outPos = 0, status = 0;
        unsigned long long cells = 0, diagonals = 0;
        bool lastTile = false;
        if (a.resume || fedPair) {   // another kernel finished some tiles of this pair before its band capacity ran out
            DevResult prev;   // read around L1: the entry may have been written by a CTA of the co-running kernel a moment ago
            {
                static_assert(sizeof(DevResult) == 40, "DevResult layout");
                const unsigned long long *src = reinterpret_cast<const unsigned long long *>(a.results + pairIdx);
                unsigned long long *dst = reinterpret_cast<unsigned long long *>(&prev);
#pragma unroll
                for (int t = 0; t < 5; ++t) dst[t] = __ldcg(src + t);
            }
            if (prev.status == kStatusRetryWide) {
                refOff = prev.resRefOff; qryOff = prev.resQryOff; tile = prev.tiles; outPos = prev.pathLen;
                cells = prev.cells; diagonals = prev.diagonals;
            }
        }
        __syncthreads();   // every thread has read the previous result before thread 0 overwrites it at the end
        const int4 redInit = make_int4(orderedInt(negInf), 0x7fffffff, -0x7fffffff, 0);

        while (!lastTile) {
            const int refLen = pr.refLen - refOff, qryLen = pr.qryLen - qryOff;
            const int cap = min(pr.fLen, min(refLen, qryLen));
            for (int t = tid; t < CW; t += NT) {
                sCS[0][t] = sCS[1][t] = sCS[2][t] = -1;
                sCI[0][t] = sCI[1][t] = kInsBoundary;
                sCD[0][t] = sCD[1][t] = kDelBoundary;
            }
            if (tid < 3) sh.convMask[tid] = 3u;
            if (tid >= 32 && tid < 36) *reinterpret_cast<int4 *>(sh.red[tid - 32]) = redInit;
            __syncthreads();

            // per-slot register state
            float h1[kSlots], i1[kSlots], d1[kSlots], h2[kSlots];      // H,I,D of diagonal k-1 and H of k-2 for the slot's row
            float q[kSlots][6], gOpQ[kSlots], gExQ[kSlots];
#pragma unroll
            for (int c = 0; c < kSlots; ++c) {
                h1[c] = i1[c] = d1[c] = h2[c] = -1.0f; gOpQ[c] = gExQ[c] = 0.f;
#pragma unroll
                for (int t = 0; t < 6; ++t) q[c][t] = 0.f;
            }
            int rowBase = -1;                                          // first of the rows whose query columns this thread holds
            int iBase = rho0;                                          // first of the rows this thread computes (= rho0 mod W, >= window base)
            float leftHPrev = -1.0f;                                   // H[k-2] of the row below slot 0

            int L0 = 0, U0 = 0, L1 = 2, U1 = -2, L2 = 1, U2 = -1;
            float maxScore = 0.0f, maxScorePrime = negInf, convScore = 0.0f;
            bool converged = false, stopped = false;
            int convValue = 0, prevConvS = -1, lastK = 0, error = 0;
            unsigned tileCells = 0;
            const int nDiag = refLen + qryLen - 1;
            const float4 *refX = reinterpret_cast<const float4 *>(refCols);
            const float4 *qryX = reinterpret_cast<const float4 *>(qryCols), *qryY = qryX + 4 * static_cast<long long>(pr.qryN4);
            const int prevWarp = (warp + NW - 1) % NW;
            int g0 = 1;
            float2 *const edgeOut = &sh.edge[0][warp];

            int c0 = 2;                                                // k % 3, kept incrementally
            for (int k = 0; k < nDiag; ++k) {
                g0 ^= 1;
                c0 = (c0 == 2) ? 0 : c0 + 1;
                const int g1 = g0 ^ 1;
                const int width = U0 - L0 + 1;
                const int Lb = L0 & ~(kSlots - 1);                     // window base: multiple of 4 so a thread's rows never wrap
                if (static_cast<unsigned>(width - 1) >= static_cast<unsigned>(cap) || U0 - Lb >= W) {   // :323-338 (width <= 0 or > cap), plus this kernel's own capacity
                    error = (width <= 0) ? 1 : ((width > cap) ? 2 : kStatusRetryWide);
                    break;
                }
                tileCells += static_cast<unsigned>(width);
                const float pruneBelow = __fsub_rn(maxScore, xdropF);

                if (SC == 0 && (k & 3) == 0 && warp == 0) {
                    // Warm L1 for the lines the band edges will touch a few diagonals from now: one 128 B line holds 8 consecutive
                    // float4 of one stream = a span of 32 columns, and the leading edge moves one column per diagonal, so every
                    // fourth diagonal the 8 streams of each side are touched once (lanes 0-7 reference columns, 8-15 query rows).
                    const int j = (lane < 8) ? refOff + min(refLen - 1, k - L0 + 43) : qryOff + min(qryLen - 1, Lb + W + 43);
                    const int n4 = (lane < 8) ? pr.refN4 : pr.qryN4;
                    const float4 *base = (lane < 8) ? refX : qryX;
                    if (lane < 16) prefetchL1(base + static_cast<long long>(lane & 7) * n4 + (j >> 2));
                }

                // row-neighbour of slot 0: last slot of the previous thread (previous warp through shared memory)
                float nbH = __shfl_up_sync(0xffffffffu, h1[kSlots - 1], 1);
                float nbI = __shfl_up_sync(0xffffffffu, i1[kSlots - 1], 1);
                {
                    const float2 e = sh.edge[g1][prevWarp];
                    nbH = (lane == 0) ? e.x : nbH;
                    nbI = (lane == 0) ? e.y : nbI;
                }

                // rows iBase .. iBase+KS-1: the smallest row >= Lb that is congruent to rho0 mod W. L0 never moves back and
                // advances by less than W per diagonal (newL <= U0 < Lb + W), so one conditional step keeps it exact for any W.
                if (iBase < Lb) iBase += W;
                if (iBase != rowBase) {                                // thread re-assigned: fetch its 4 query columns
                    rowBase = iBase;
#pragma unroll
                    for (int c = 0; c < kSlots; ++c) {
                        if (SC == 1) {   // only the row's gap penalties: columns are P + 2 = 24 floats, penalties last
                            const float2 g = __ldg(reinterpret_cast<const float2 *>(qryCols + static_cast<size_t>(qryOff + min(iBase + c, qryLen - 1)) * 24 + 22));
                            gOpQ[c] = g.x; gExQ[c] = g.y;
                            continue;
                        }
                        const long long at = ntColIndex(qryOff + min(iBase + c, qryLen - 1), pr.qryN4);
                        const float4 x = __ldg(qryX + at);
                        const float4 y = __ldg(qryY + at);
                        q[c][0] = x.x; q[c][1] = x.y; q[c][2] = x.z; q[c][3] = x.w; q[c][4] = y.x; q[c][5] = y.y;
                        gOpQ[c] = y.z; gExQ[c] = y.w;
                        if (kind) {
                            // w[a] = sum over m of q[m]*S[a][m] in the reference's order: the whole contraction of this row
                            // against a one-hot reference column of letter a (kRefOneHot), or simply S[a][b] when the row
                            // itself is one-hot at b (kQryOneHot).
                            float w[5];
#pragma unroll
                            for (int l = 0; l < 5; ++l)
                                w[l] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(q[c][0], a.scoreNt[l * 5 + 0]), __fmul_rn(q[c][1], a.scoreNt[l * 5 + 1])),
                                                 __fmul_rn(q[c][2], a.scoreNt[l * 5 + 2])), __fmul_rn(q[c][3], a.scoreNt[l * 5 + 3])), __fmul_rn(q[c][4], a.scoreNt[l * 5 + 4]));
                            if (kind & kRefOneHot) {
                                // + the one surviving gap-character term (TALCO-XDrop.cpp:393 with r[a] = 1), then the division
#pragma unroll
                                for (int l = 0; l < 5; ++l) {
                                    const float n = __fmaf_rn(q[c][5], gapChar, w[l]);
                                    w[l] = (divMode == 0) ? n : __fdiv_rn(n, denom);
                                }
                            }
#pragma unroll
                            for (int l = 0; l < 5; ++l) q[c][l] = w[l];
                            q[c][5] = 0.0f;
                        }
                    }
                }

                float myMax = negInf;
                int myLo = 0x7fffffff, myHi = -0x7fffffff;
                unsigned tbWord = 0, actBits = 0;
                const bool anyAct = (iBase <= U0) && (iBase + kSlots - 1 >= L0);

                if (__any_sync(0xffffffffu, anyAct)) {
                    float num[kSlots], gOpR[kSlots], gExR[kSlots];
                    if (SC == 1) {
                        // cell (row i, column j) of the pair lives at [i + j][i]: the rows of a thread are adjacent words, the rows of a
                        // warp one contiguous run. Slots that are not live may read beyond the matrix (padded, values discarded).
                        const float *src = simBase + static_cast<long long>(refOff + qryOff + k) * simStride + (qryOff + iBase);
                        const float2 *gsrc = gapRef + (refOff + k - iBase);
#pragma unroll
                        for (int c = 0; c < kSlots; ++c) {
                            num[c] = __ldg(src + c);
                            const float2 g = __ldg(gsrc - c);
                            gOpR[c] = g.x; gExR[c] = g.y;
                        }
                    } else {
                        scoreDiagonal<MC, KS>(a, refX, pr.refN4, refOff + k, iBase, q, gapChar, kind, divMode, denom, rcp, num, gOpR, gExR);
                    }

                    // match candidates: H[k-2][i-1] + sim when the diagonal neighbour is inside its band ...
                    float match[kSlots];
#pragma unroll
                    for (int c = 0; c < kSlots; ++c) {
                        const int i = iBase + c;
                        const float diagH = (c == 0) ? leftHPrev : h2[c - 1 < 0 ? 0 : c - 1];
                        const bool diagIn = (i - 1 >= L2) && (i - 1 <= U2);
                        match[c] = diagIn ? __fadd_rn(diagH, num[c]) : negInf;
                    }
                    // ... except on diagonal 0 and on the first row / column of the first tile (TALCO-XDrop.cpp:369-371, 445-449):
                    // only the warp that holds row 0 or the cell of column 0 (row k) takes this path
                    const bool edgeCell = (tile == 0) && (iBase == 0 || static_cast<unsigned>(k - iBase) < static_cast<unsigned>(kSlots));
                    if (k == 0 || __any_sync(0xffffffffu, edgeCell)) {
#pragma unroll
                        for (int c = 0; c < kSlots; ++c) {
                            const int i = iBase + c, j = k - i;
                            if (tile == 0 && (i == 0 || j == 0)) {
                                if (i == 0 && j == 0) match[c] = num[c];
                                else match[c] = __fmaf_rn(a.gapExtend, static_cast<float>(max(0, max(refOff + j, qryOff + i) - 1)), __fadd_rn(num[c], a.gapOpen));
                            } else if (k == 0) match[c] = num[c];
                        }
                    }

                    // recurrences, highest slot first so every slot still sees its neighbour's previous-diagonal values
#pragma unroll
                    for (int c = kSlots - 1; c >= 0; --c) {
                        const int i = iBase + c;
                        const bool act = (i >= L0) && (i <= U0);
                        const float leftH = (c == 0) ? nbH : h1[c - 1 < 0 ? 0 : c - 1];
                        const float leftI = (c == 0) ? nbI : i1[c - 1 < 0 ? 0 : c - 1];
                        const bool upIn = (i <= U1);                    // i >= L0 >= L1 for live cells
                        const bool leftIn = (i > L1);                   // i-1 <= U0-1 <= U1 for live cells
                        const float delOpen = upIn ? __fadd_rn(h1[c], gOpR[c]) : negInf;
                        const float delExt = upIn ? __fadd_rn(d1[c], gExR[c]) : negInf;
                        const float insOpen = leftIn ? __fadd_rn(leftH, gOpQ[c]) : negInf;
                        const float insExt = leftIn ? __fadd_rn(leftI, gExQ[c]) : negInf;
                        const bool insFromIns = insExt >= insOpen, delFromDel = delExt >= delOpen;
                        const float insBest = fmaxf(insExt, insOpen), delBest = fmaxf(delExt, delOpen);
                        const bool mGeI = match[c] >= insBest, mGeD = match[c] >= delBest, iGtD = insBest > delBest;
                        const unsigned ptr = (mGeI && mGeD) ? 0u : ((!mGeI && iGtD) ? 1u : 2u);
                        float s = fmaxf(match[c], fmaxf(insBest, delBest));
                        s = (s < pruneBelow || !act) ? negInf : s;
                        h2[c] = h1[c]; h1[c] = s; i1[c] = insBest; d1[c] = delBest;
                        myMax = fmaxf(myMax, s);
                        const bool alive = s > negInf;
                        myLo = alive ? i : myLo;                        // slots run downwards: the last hit is the lowest row
                        myHi = (alive && myHi < 0) ? i : myHi;          // the first hit is the highest row
                        actBits |= (act ? 1u : 0u) << c;
                        tbWord |= (ptr | (insFromIns ? 4u : 0u) | (delFromDel ? 8u : 0u)) << (8 * c);
                    }
                    if (k <= marker && actBits) {
                        uint8_t *dst = tb + static_cast<size_t>(k) * W + rho0;
                        if (KS == 4) *reinterpret_cast<unsigned *>(dst) = tbWord;
                        else if (KS == 2) *reinterpret_cast<unsigned short *>(dst) = static_cast<unsigned short>(tbWord);
                        else *dst = static_cast<uint8_t>(tbWord);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < kSlots; ++c) h2[c] = h1[c];
                }
                leftHPrev = nbH;

                int cs[kSlots], ci[kSlots], cd[kSlots];
#pragma unroll
                for (int c = 0; c < kSlots; ++c) cs[c] = ci[c] = cd[c] = 0;
                if (k >= marker - 1 && __any_sync(0xffffffffu, actBits != 0)) {   // convergence pointers, reference indexing (:520-547)
                    const int c1 = (c0 == 0) ? 2 : c0 - 1, c2 = (c0 == 2) ? 0 : c0 + 1;
                    if (k <= marker) {
#pragma unroll
                        for (int c = 0; c < kSlots; ++c) {
                            if (actBits & (1u << c)) {
                                const int i = iBase + c, off = i - L0;
                                if (k == marker - 1) { cs[c] = (3 << 16) | (i & 0xFFFF); sCS[c0][off] = cs[c]; }
                                else {
                                    cs[c] = (i & 0xFFFF); ci[c] = (1 << 16) | (i & 0xFFFF); cd[c] = (2 << 16) | (i & 0xFFFF);
                                    sCS[c0][off] = cs[c]; sCI[g0][off] = ci[c]; sCD[g0][off] = cd[c];
                                }
                            }
                        }
                    } else {
                        // branch-free: every slot reads its five source slots at clamped offsets; only live cells store
                        const int *pCIp = sCI[g1], *pCDp = sCD[g1], *pCS1 = sCS[c1], *pCS2 = sCS[c2];
                        int *pCI = sCI[g0], *pCD = sCD[g0], *pCS = sCS[c0];
#pragma unroll
                        for (int c = 0; c < kSlots; ++c) {
                            const int i = iBase + c;
                            const int off = i - L0, offDiag = i - 1 - L2, offUp = i - L1, offLeft = offUp - 1;
                            const int oL = min(max(offLeft, 0), CW - 1), oU = min(max(offUp, 0), CW - 1), oD = min(max(offDiag, 0), CW - 1);
                            const unsigned nib = (tbWord >> (8 * c)) & 15u;
                            const int aI = pCIp[oL], bI = pCS1[oL], aD = pCDp[oU], bD = pCS1[oU], dS = pCS2[oD];
                            int vi = (nib & 4u) ? aI : ((bI != -1) ? bI : kInsBoundary);
                            vi = (offLeft >= 0) ? vi : kInsBoundary;
                            int vd = (nib & 8u) ? aD : ((bD != -1) ? bD : kDelBoundary);
                            vd = (offUp >= 0) ? vd : kDelBoundary;
                            const unsigned ptr = nib & 3u;
                            const int vs = (ptr == 0u) ? ((offDiag >= 0) ? dS : -1) : ((ptr == 1u) ? vi : vd);
                            ci[c] = vi; cd[c] = vd; cs[c] = vs;
                            if (actBits & (1u << c)) { pCI[off] = vi; pCD[off] = vd; pCS[off] = vs; }
                        }
                    }
                }

                // one barrier per diagonal: publish the warp's edge values and its reduction
                const int wHi = __reduce_max_sync(0xffffffffu, myHi);
                if (lane == 31) edgeOut[g0 * 32] = make_float2(h1[kSlots - 1], i1[kSlots - 1]);
                if (wHi >= 0) {                                        // the warp holds a cell that survived (warp-uniform)
                    const int wMax = __reduce_max_sync(0xffffffffu, orderedInt(myMax));
                    const int wLo = __reduce_min_sync(0xffffffffu, myLo);
                    if (lane == 0) {
                        int *r = sh.red[k & 3];
                        atomicMax(r, wMax); atomicMin(r + 1, wLo); atomicMax(r + 2, wHi);
                    }
                }
                if (tid == NT - 1) *reinterpret_cast<int4 *>(sh.red[(k + 2) & 3]) = redInit;   // last read after barrier k-2, next folded after barrier k+1
                __syncthreads();
                int oMax, newL, newU;
                {
                    const int4 r = *reinterpret_cast<const int4 *>(sh.red[k & 3]);
                    oMax = r.x; newL = r.y; newU = r.z;
                }
                if (newL == 0x7fffffff) { newL = U0 + 1; newU = L0 - 1; }
                maxScorePrime = fmaxf(maxScorePrime, orderedFloat(oMax));

                if (!converged && k >= marker && k < nDiag - 1) {       // :585-595
                    const int c2 = (c0 == 2) ? 0 : c0 + 1;
                    const int start = newL - L0;
                    const int vI = sCI[g0][start], vD = sCD[g0][start], vS = sCS[c0][start];
                    unsigned bad = 0;
#pragma unroll
                    for (int c = 0; c < kSlots; ++c) {
                        const int i = iBase + c;
                        if ((actBits & (1u << c)) && i > newL && i <= newU) {
                            if (ci[c] != vI || cd[c] != vD) bad |= 1u;
                            if (cs[c] != vS) bad |= 2u;
                        }
                    }
                    if (bad) atomicAnd(&sh.convMask[c0], ~bad);
                    if (tid == 0) sh.convMask[c2] = 3u;
                    __syncthreads();
                    const unsigned ok = sh.convMask[c0];
                    const int cS = (ok & 2u) ? vS : -1;
                    if ((ok & 1u) && vI == vD && vI == cS && prevConvS == cS && vI != -1) {
                        converged = true;
                        convValue = prevConvS;
                        convScore = maxScorePrime;
                    }
                    prevConvS = cS;
                }

                L2 = L1; U2 = U1; L1 = L0; U1 = U0;
                L0 = max(newL, max(0, k + 2 - refLen));
                U0 = min(qryLen - 1, newU + 1);
                maxScore = fmaxf(maxScorePrime, 0.0f);                   // max(0, max_score_prime), :607 (a -0 maximum compares and subtracts like +0)
                lastK = k;
                if (converged && maxScore > convScore) { stopped = true; break; }
            }
            if (error != kStatusRetryWide) {   // a tile that has to be redone by a wider kernel is not counted here
                cells += tileCells;
                diagonals += static_cast<unsigned long long>(lastK + 1);
            }
            const int nStored = min(lastK, marker) + 1;

            if (error) { status = error; break; }

            __syncthreads();
            if (tid == 0) {
                int convQry, convRef, startDiag, tbState, isLast = 0;
                if (stopped || lastK >= marker) {
                    const int v = stopped ? convValue : sCS[lastK % 3][0];
                    convQry = v & 0xFFFF;
                    tbState = static_cast<int8_t>((v >> 16) & 0xFFFF);
                    convRef = marker - convQry - ((tbState == 3) ? 1 : 0);
                    startDiag = (tbState == 3) ? nStored - 2 : nStored - 1;
                } else {
                    convQry = qryLen - 1; convRef = refLen - 1; startDiag = lastK; tbState = 0; isLast = 1;
                }
                if (convQry == (kDelBoundary & 0xFFFF)) { convQry = 0; convRef = marker; }
                else if (convQry == (kInsBoundary & 0xFFFF)) { convQry = marker; convRef = 0; }
                const int newRefOff = refOff + convRef, newQryOff = qryOff + convQry;
                int err = 0, tailLen = 0, tailOp = 0;
                if (pr.refLen - newRefOff < 0 || pr.qryLen - newQryOff < 0) err = 3;
                if (newRefOff == pr.refLen - 1 && newQryOff < pr.qryLen - 1) { tailLen = pr.qryLen - newQryOff - 1; tailOp = 1; isLast = 1; }
                if (newQryOff == pr.qryLen - 1 && newRefOff < pr.refLen - 1) { tailLen = pr.refLen - newRefOff - 1; tailOp = 2; isLast = 1; }
                if (newRefOff == pr.refLen - 1 && newQryOff == pr.qryLen - 1) isLast = 1;

                constexpr int opsCap = 2 * kMaxMarker + 16;
                int w = opsCap;
                if (!err) {
                    int kk = startDiag;
                    int row = static_cast<int16_t>(convQry), qi = row, ri = static_cast<int16_t>(convRef);
                    int state = static_cast<int8_t>(tbState) % 3;
                    const bool first = (tile == 0);
                    while (kk >= 0 && w > 0) {
                        // the path drifts about half a row per step: pull the line 24 diagonals back into L1 now
                        if (kk >= 24) prefetchL1(tb + static_cast<size_t>(kk - 24) * W + ((row + W - 12) % W));
                        const int cell = tb[static_cast<size_t>(kk) * W + (row % W)];
                        int dir;
                        if (state == 0) {
                            const int p = cell & 3;
                            if (p == 0) { dir = 0; }
                            else if (p == 1) { dir = 1; state = (cell & 4) ? 1 : 0; }
                            else { dir = 2; state = (cell & 8) ? 2 : 0; }
                        } else if (state == 1) { dir = 1; state = (cell & 4) ? 1 : 0; }
                        else { dir = 2; state = (cell & 8) ? 2 : 0; }
                        if (dir == 0) { kk -= 2; row -= 1; qi--; ri--; }
                        else if (dir == 1) { kk -= 1; row -= 1; qi--; }
                        else { kk -= 1; ri--; }
                        sh.ops[--w] = static_cast<int8_t>(dir);
                        if (first && (ri < 0 || qi < 0)) break;
                    }
                    if (first) {
                        while (ri > -1 && w > 0) { sh.ops[--w] = 2; ri--; }
                        while (qi > -1 && w > 0) { sh.ops[--w] = 1; qi--; }
                    }
                }
                sh.refOff = newRefOff; sh.qryOff = newQryOff; sh.lastTile = isLast; sh.error = err;
                sh.opsBegin = w; sh.nOps = opsCap - w; sh.tailLen = tailLen; sh.tailOp = tailOp;
            }
            __syncthreads();
            if (sh.error) { status = sh.error; break; }
            if (sh.nOps + sh.tailLen == 0) { status = 3; break; }
            {
                const int skip = (tile > 0) ? 1 : 0;
                const int nCopy = sh.nOps - skip;
                const int8_t *src = sh.ops + sh.opsBegin + skip;
                for (int t = tid; t < nCopy; t += NT) path[outPos + t] = src[t];
                const int8_t tailOp = static_cast<int8_t>(sh.tailOp);
                for (int t = tid; t < sh.tailLen; t += NT) path[outPos + nCopy + t] = tailOp;
                outPos += nCopy + sh.tailLen;
            }
            refOff = sh.refOff; qryOff = sh.qryOff; lastTile = sh.lastTile != 0;
            ++tile;
            if (a.coMode == 1 && tid == 0) atomicAdd(a.heartbeat, 1);
            __syncthreads();
        }

        if (tid == 0) {
            DevResult res;
            res.status = status;
            res.pathLen = status ? 0 : outPos;
            res.tiles = tile;
            res.pad = 0;
            res.cells = cells;
            res.diagonals = diagonals;
            res.resRefOff = refOff; res.resQryOff = qryOff;
            if (status == kStatusRetryWide) { res.pathLen = outPos; }
            a.results[pairIdx] = res;
            if (status == kStatusRetryWide) {
                if (a.coMode == 1) {   // hand the pair to the co-running wide workers: result and partial path first, then the entry
                    __threadfence();
                    if (a.coTrace) a.coTrace[4 * pairIdx + 0] = globalTimerNs();
                    const int slot = atomicAdd(a.feedCount, 1);
                    atomicExch(a.feedList + slot, pairIdx);
                } else if (a.overflowList != nullptr) a.overflowList[atomicAdd(a.overflowCount, 1)] = pairIdx;
            }
            if (a.coMode && !fedPair) { __threadfence(); atomicAdd(a.mainDone, 1); }
            if (a.coTrace && a.coMode == 2) { a.coTrace[4 * pairIdx + 2] = globalTimerNs(); a.coTrace[4 * pairIdx + 3] = fedPair ? 2 : 1; }
        }
    }
    if (a.coTrace && tid == 0 && blockIdx.x == 0) a.coTrace[4 * (*a.nWorkPtr) + (a.coMode == 2 ? 1 : 0)] = globalTimerNs();
}

// Co-run gate: one thread, launched on the narrow kernel's stream right before it. The narrow grid is sized to leave `want` SMs to
// the wide workers, but the block scheduler spreads its CTAs over every SM unless those SMs are already taken: without the gate
// the two launches race, and when the narrow kernel wins no wide CTA (512 threads, a whole SM) fits until the narrow kernel
// ends — every hand-over then waits for the end of the stage (measured: +6 ms on a 62 ms level). The gate returns as soon as
// the wide CTAs have counted themselves in, or after 300 us (kernels serialised by a profiler: nothing will arrive).
__global__ void coRunGateKernel(const int *arrived, int want) {
    const unsigned long long t0 = globalTimerNs();
    while (*reinterpret_cast<volatile const int *>(arrived) < want && globalTimerNs() - t0 < 300000ull) __nanosleep(200);
}
cudaError_t launchCoRunGate(const int *arrived, int want, cudaStream_t stream) {
    coRunGateKernel<<<1, 1, 0, stream>>>(arrived, want);
    return cudaGetLastError();
}

// The wavefront kernels hold bands of up to W-(KS-1) cells (the window base is rounded down to a multiple of KS).
int wavefrontBandCapacity(int threads, int slots) { return threads * slots - (slots - 1); }
int wavefrontWindow(int threads, int slots) { return threads * slots; }

// 1 when the matrix has the built-in nucleotide shape with an all-zero N row/column (see numerators4).
int nucleotideMatrixClass(const float *s) {
    const float A = s[0], B = s[2], C = s[1];
    for (int l = 0; l < 5; ++l)
        for (int m = 0; m < 5; ++m) {
            const float want = (l == 4 || m == 4) ? 0.0f : ((l == m) ? A : ((l - m == 2 || m - l == 2) ? B : C));
            if (s[l * 5 + m] != want) return 0;
        }
    return 1;
}

// Self-test of exactDiv against the IEEE divide (twl_selftest_division): counts the operand pairs on which they differ.
__global__ void divSelfTestKernel(const float *num, const float *den, int n, int *mismatches) {
    int bad = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const float d = den[t];
        const float viaRcp = exactDiv(num[t], d, __fdiv_rn(1.0f, d));
        const float ieee = __fdiv_rn(num[t], d);
        if (__float_as_int(viaRcp) != __float_as_int(ieee) && !((__float_as_int(d) & 0x7fffff) == 0x7fffff)) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}
cudaError_t launchDivSelfTest(const float *num, const float *den, int n, int *mismatches, cudaStream_t stream) {
    divSelfTestKernel<<<296, 256, 0, stream>>>(num, den, n, mismatches);
    return cudaGetLastError();
}

// supported instantiations: (threads, slots) = (128,4) throughput, (512,2) wide band / low latency, (512,1) narrow-band low latency
template <int MC>
static cudaError_t launchMc(int threads, int slots, const TalcoArgs &args, int grid, cudaStream_t stream) {
    if (threads == 128 && slots == 4) talcoWavefrontKernel<128, MC, 4, 0><<<grid, 128, 0, stream>>>(args);
    else if (threads == 512 && slots == 1) talcoWavefrontKernel<512, MC, 1, 0><<<grid, 512, 0, stream>>>(args);
    else if (threads == 512 && slots == 2) talcoWavefrontKernel<512, MC, 2, 0><<<grid, 512, 0, stream>>>(args);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// matClass: 0 / 1 nucleotide (see nucleotideMatrixClass); -1: scores come from the similarity matrix (protein path, SC = 1)
cudaError_t launchTalcoWavefront(int threads, int slots, int matClass, const TalcoArgs &args, int grid, cudaStream_t stream) {
    if (matClass < 0) {
        if (threads == 128 && slots == 4) talcoWavefrontKernel<128, 0, 4, 1><<<grid, 128, 0, stream>>>(args);
        else if (threads == 512 && slots == 2) talcoWavefrontKernel<512, 0, 2, 1><<<grid, 512, 0, stream>>>(args);
        else return cudaErrorInvalidValue;
        return cudaGetLastError();
    }
    return matClass == 1 ? launchMc<1>(threads, slots, args, grid, stream) : launchMc<0>(threads, slots, args, grid, stream);
}

int wavefrontMaxCtasPerSm(int threads, int slots, int matClass) {
    int n = 0;
#define TWL_OCC(NT_, MC_, KS_, SC_) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, talcoWavefrontKernel<NT_, MC_, KS_, SC_>, NT_, 0)
    if (matClass < 0) {
        if (threads == 128 && slots == 4) TWL_OCC(128, 0, 4, 1); else if (threads == 512 && slots == 2) TWL_OCC(512, 0, 2, 1);
    } else if (matClass == 1) {
        if (threads == 128 && slots == 4) TWL_OCC(128, 1, 4, 0); else if (threads == 512 && slots == 1) TWL_OCC(512, 1, 1, 0); else if (threads == 512 && slots == 2) TWL_OCC(512, 1, 2, 0);
    } else {
        if (threads == 128 && slots == 4) TWL_OCC(128, 0, 4, 0); else if (threads == 512 && slots == 1) TWL_OCC(512, 0, 1, 0); else if (threads == 512 && slots == 2) TWL_OCC(512, 0, 2, 0);
    }
#undef TWL_OCC
    return n;
}

} // namespace twl

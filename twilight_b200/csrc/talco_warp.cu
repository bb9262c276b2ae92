// talco_warp.cu — throughput variant of the nucleotide TALCO-XDrop kernel: ONE PAIR PER WARP.
//
// Where talco_wavefront.cu spreads an anti-diagonal over the 4 warps of a CTA (best latency for a single pair, but one
// CTA barrier and ~230 instructions of per-thread bookkeeping per diagonal, and a warp that idles when the band is
// narrower than 3/4 of the window), this kernel gives every warp its own pair: the band is swept in passes of 32 cells,
// the only synchronisation is __syncwarp, the per-diagonal bookkeeping is paid once per ~11 cells of a lane, and the
// serial traceback of one pair no longer stalls three other warps. The wavefront state (H of two diagonals, I, D) lives
// in shared memory indexed by (row mod kCap); I and D are updated in place and H needs two buffers because the passes
// run from the highest rows down and every pass reads before it writes. The convergence pointers keep the reference's
// (row - L[k]) indexing and rotation depth (stale slots are observable).
//
// Semantics, float operation order and all tie rules are those of talco_wavefront.cu / the reference CPU path
// (src/TALCO-XDrop.cpp:233-689). Bands wider than kCap-2 go to the overflow list (resumed by the wider kernels).
#include "talco_score.cuh"
#include "twl_device.cuh"

namespace twl {

constexpr int kWarpCap = 512;                 // rows of wavefront state per warp (band <= kWarpCap - 2)
constexpr int kWarpCW = kWarpCap + 4;

struct WarpShared {
    float H[2][kWarpCap];
    float I[kWarpCap];
    float D[kWarpCap];
    short CS[3][kWarpCW], CI[2][kWarpCW], CD[2][kWarpCW];   // packed convergence pointers (see packConv)
    int8_t ops[2 * kMaxMarker + 16];
};

// Convergence pointers are -1, -2 (I_BOUNDARY), -3 (D_BOUNDARY) or (state << 16) | row with state in 0..3 and
// row <= marker <= 1024: they fit 16 bits as state*2048 + row.
__device__ __forceinline__ short packConv(int v) { return (v < 0) ? static_cast<short>(v) : static_cast<short>(((v >> 16) << 11) | (v & 0x7FF)); }
__device__ __forceinline__ int unpackConv(short s) { return (s < 0) ? static_cast<int>(s) : (((static_cast<int>(s) >> 11) << 16) | (static_cast<int>(s) & 0x7FF)); }

template <int MC>
__device__ __forceinline__ float numeratorCell(const float (&r)[6], const float (&q)[6], const TalcoArgs &a) {
    if (MC == 0) {
        float n = 0.0f;
#pragma unroll
        for (int l = 0; l < 5; ++l) {
            const float t0 = __fmul_rn(__fmul_rn(q[0], a.scoreNt[l * 5 + 0]), r[l]);
            const float t1 = __fmul_rn(__fmul_rn(q[1], a.scoreNt[l * 5 + 1]), r[l]);
            const float t2 = __fmul_rn(__fmul_rn(q[2], a.scoreNt[l * 5 + 2]), r[l]);
            const float t3 = __fmul_rn(__fmul_rn(q[3], a.scoreNt[l * 5 + 3]), r[l]);
            const float t4 = __fmul_rn(__fmul_rn(q[4], a.scoreNt[l * 5 + 4]), r[l]);
            n = __fadd_rn(n, __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3), t4));
        }
        return n;
    }
    const float A = a.scoreNt[0], B = a.scoreNt[2], C = a.scoreNt[1];
    float qa[4], qb[4], qc[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) { qa[m] = __fmul_rn(q[m], A); qb[m] = __fmul_rn(q[m], B); qc[m] = __fmul_rn(q[m], C); }
    const float h0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qa[0], r[0]), __fmul_rn(qc[1], r[0])), __fmul_rn(qb[2], r[0])), __fmul_rn(qc[3], r[0]));
    const float h1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qc[0], r[1]), __fmul_rn(qa[1], r[1])), __fmul_rn(qc[2], r[1])), __fmul_rn(qb[3], r[1]));
    const float h2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qb[0], r[2]), __fmul_rn(qc[1], r[2])), __fmul_rn(qa[2], r[2])), __fmul_rn(qc[3], r[2]));
    const float h3 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qc[0], r[3]), __fmul_rn(qb[1], r[3])), __fmul_rn(qc[2], r[3])), __fmul_rn(qa[3], r[3]));
    return __fadd_rn(__fadd_rn(__fadd_rn(h0, h1), h2), h3);
}

template <int MC>
__global__ void __launch_bounds__(32) talcoWarpKernel(const TalcoArgs a) {
    __shared__ WarpShared sh;
    const int lane = threadIdx.x;
    constexpr int MASK = kWarpCap - 1;
    uint8_t *tb = a.tbScratch + static_cast<size_t>(blockIdx.x) * a.tbStride;   // tb[k][row mod kCap]
    const int marker = a.marker;

    for (;;) {
        int work = 0;
        if (lane == 0) work = atomicAdd(a.queue, 1);
        work = __shfl_sync(0xffffffffu, work, 0);
        if (work >= *a.nWorkPtr) break;
        const int pairIdx = a.order[work];
        const DevPair pr = a.pairs[pairIdx];
        if (pr.refLen < 1 || pr.qryLen < 1) {
            if (lane == 0) {
                DevResult res;
                res.status = kStatusEmptySide; res.pathLen = 0; res.tiles = 0; res.pad = 0; res.cells = 0; res.diagonals = 0; res.resRefOff = 0; res.resQryOff = 0;
                a.results[pairIdx] = res;
            }
            continue;
        }
        const float4 *refX = reinterpret_cast<const float4 *>(a.prof + pr.refOff), *refY = refX + 4 * static_cast<long long>(pr.refN4);
        const float4 *qryX = reinterpret_cast<const float4 *>(a.prof + pr.qryOff), *qryY = qryX + 4 * static_cast<long long>(pr.qryN4);
        int8_t *path = a.paths + pr.alnOff;
        const float negInf = -static_cast<float>(2.0 * pr.xdrop + 1.0);
        const float xdropF = static_cast<float>(pr.xdrop);
        const float denom = __fmul_rn(pr.refNum, pr.qryNum);
        const float rcp = __fdiv_rn(1.0f, denom);
        const int divMode = (denom == 1.0f) ? 0 : (((__float_as_int(denom) & 0x7fffff) == 0x7fffff) ? 2 : 1);
        const float gapChar = pr.gapChar;
        int refOff = 0, qryOff = 0, tile = 0, outPos = 0, status = 0;
        unsigned long long cells = 0, diagonals = 0;
        bool lastTile = false;

        while (!lastTile) {
            const int refLen = pr.refLen - refOff, qryLen = pr.qryLen - qryOff;
            const int cap = min(pr.fLen, min(refLen, qryLen));
            for (int t = lane; t < kWarpCW; t += 32) {
                sh.CS[0][t] = sh.CS[1][t] = sh.CS[2][t] = -1;
                sh.CI[0][t] = sh.CI[1][t] = static_cast<short>(kInsBoundary);
                sh.CD[0][t] = sh.CD[1][t] = static_cast<short>(kDelBoundary);
            }
            __syncwarp();
            int L0 = 0, U0 = 0, L1 = 2, U1 = -2, L2 = 1, U2 = -1;
            float maxScore = 0.0f, maxScorePrime = negInf, convScore = 0.0f;
            bool converged = false, stopped = false;
            int convValue = 0, prevConvS = -1, lastK = 0, error = 0;
            unsigned tileCells = 0;
            const int nDiag = refLen + qryLen - 1;
            int hb = 0;                                                 // H[hb] holds diagonal k-1, H[hb^1] holds k-2 and receives k
            int c0 = 0, c1 = 2, c2 = 1, g0 = 1;

            for (int k = 0; k < nDiag; ++k) {
                g0 ^= 1;
                const int g1 = g0 ^ 1;
                const int width = U0 - L0 + 1;
                if (width <= 0 || width > cap || width > kWarpCap - 2) {
                    error = (width <= 0) ? 1 : ((width > cap) ? 2 : kStatusRetryWide);
                    break;
                }
                tileCells += static_cast<unsigned>(width);
                const float pruneBelow = __fsub_rn(maxScore, xdropF);
                const bool special = (k == 0) || (tile == 0 && (L0 == 0 || U0 == k));
                const bool conv = (k >= marker - 1);
                float *Hprev = sh.H[hb], *Hout = sh.H[hb ^ 1];
                float myMax = negInf;
                int myLo = 0x7fffffff, myHi = -0x7fffffff;
                const int nPass = (width + 31) >> 5;

                for (int p = nPass - 1; p >= 0; --p) {                  // highest rows first: a pass reads before any lower pass writes
                    const int i = L0 + (p << 5) + lane;
                    const bool act = i <= U0;
                    const int ic = min(i, U0);                          // clamp so that idle lanes read valid columns
                    const int j = k - ic;
                    float r[6], q[6], gOpR, gExR, gOpQ, gExQ;
                    {
                        const long long ar = ntColIndex(refOff + j, pr.refN4), aq = ntColIndex(qryOff + ic, pr.qryN4);
                        const float4 x = __ldg(refX + ar), y = __ldg(refY + ar), u = __ldg(qryX + aq), v = __ldg(qryY + aq);
                        r[0] = x.x; r[1] = x.y; r[2] = x.z; r[3] = x.w; r[4] = y.x; r[5] = y.y; gOpR = y.z; gExR = y.w;
                        q[0] = u.x; q[1] = u.y; q[2] = u.z; q[3] = u.w; q[4] = v.x; q[5] = v.y; gOpQ = v.z; gExQ = v.w;
                    }
                    const int si = ic & MASK, sl = (ic - 1) & MASK;
                    const float hUp = Hprev[si], dUp = sh.D[si], hLeft = Hprev[sl], iLeft = sh.I[sl], hDiag = Hout[sl];
                    __syncwarp();                                       // every lane has read before anyone overwrites

                    float num = numeratorCell<MC>(r, q, a);
                    if (__any_sync(0xffffffffu, q[5] != 0.0f)) {
#pragma unroll
                        for (int l = 0; l < 5; ++l) num = __fmaf_rn(__fmul_rn(r[l], q[5]), gapChar, num);
                    }
                    if (__any_sync(0xffffffffu, r[5] != 0.0f)) {
#pragma unroll
                        for (int m = 0; m < 5; ++m) num = __fmaf_rn(__fmul_rn(r[5], q[m]), gapChar, num);
                    }
                    if (divMode == 1) num = exactDiv(num, denom, rcp);
                    else if (divMode == 2) num = __fdiv_rn(num, denom);

                    const bool upIn = (ic <= U1) && (ic >= L1);
                    const bool leftIn = (ic - 1 >= L1) && (ic - 1 <= U1);
                    const bool diagIn = (ic - 1 >= L2) && (ic - 1 <= U2);
                    float match = diagIn ? __fadd_rn(hDiag, num) : negInf;
                    if (special) {
                        const bool onEdge0 = (tile == 0) && (ic == 0 || j == 0);
                        if (onEdge0) {
                            if (ic == 0 && j == 0) match = num;
                            else match = __fmaf_rn(a.gapExtend, static_cast<float>(max(0, max(refOff + j, qryOff + ic) - 1)), __fadd_rn(num, a.gapOpen));
                        } else if (k == 0) match = num;
                    }
                    const float delOpen = upIn ? __fadd_rn(hUp, gOpR) : negInf;
                    const float delExt = upIn ? __fadd_rn(dUp, gExR) : negInf;
                    const float insOpen = leftIn ? __fadd_rn(hLeft, gOpQ) : negInf;
                    const float insExt = leftIn ? __fadd_rn(iLeft, gExQ) : negInf;
                    const bool insFromIns = insExt >= insOpen, delFromDel = delExt >= delOpen;
                    const float insBest = fmaxf(insExt, insOpen), delBest = fmaxf(delExt, delOpen);
                    const bool mGeI = match >= insBest, mGeD = match >= delBest, iGtD = insBest > delBest;
                    const int ptr = (mGeI && mGeD) ? 0 : ((!mGeI && iGtD) ? 1 : 2);
                    float s = fmaxf(match, fmaxf(insBest, delBest));
                    if (s < pruneBelow) s = negInf;
                    if (act) {
                        Hout[si] = s; sh.I[si] = insBest; sh.D[si] = delBest;
                        myMax = fmaxf(myMax, s);
                        if (s > negInf) { myLo = min(myLo, i); myHi = max(myHi, i); }
                        if (k <= marker) tb[static_cast<size_t>(k) * kWarpCap + si] = static_cast<uint8_t>(ptr | (insFromIns ? 4 : 0) | (delFromDel ? 8 : 0));
                        if (conv) {                                     // :520-547, reference indexing
                            const int off = i - L0, offDiag = i - 1 - L2, offUp = i - L1, offLeft = offUp - 1;
                            if (k == marker - 1) sh.CS[c0][off] = packConv((3 << 16) | (i & 0xFFFF));
                            else if (k == marker) {
                                sh.CS[c0][off] = packConv(i & 0xFFFF);
                                sh.CI[g0][off] = packConv((1 << 16) | (i & 0xFFFF));
                                sh.CD[g0][off] = packConv((2 << 16) | (i & 0xFFFF));
                            } else {
                                short vi, vd;
                                if (insFromIns) vi = (offLeft >= 0) ? sh.CI[g1][offLeft] : static_cast<short>(kInsBoundary);
                                else { const short t = (offLeft >= 0) ? sh.CS[c1][offLeft] : static_cast<short>(-1); vi = (t != -1) ? t : static_cast<short>(kInsBoundary); }
                                if (delFromDel) vd = (offUp >= 0) ? sh.CD[g1][offUp] : static_cast<short>(kDelBoundary);
                                else { const short t = (offUp >= 0) ? sh.CS[c1][offUp] : static_cast<short>(-1); vd = (t != -1) ? t : static_cast<short>(kDelBoundary); }
                                const short vs = (ptr == 0) ? ((offDiag >= 0) ? sh.CS[c2][offDiag] : static_cast<short>(-1)) : ((ptr == 1) ? vi : vd);
                                sh.CI[g0][off] = vi; sh.CD[g0][off] = vd; sh.CS[c0][off] = vs;
                            }
                        }
                    }
                }
                __syncwarp();
                hb ^= 1;

                int newL = __reduce_min_sync(0xffffffffu, myLo), newU = __reduce_max_sync(0xffffffffu, myHi);
                const int oMax = __reduce_max_sync(0xffffffffu, __float_as_int(myMax) ^ ((__float_as_int(myMax) >> 31) & 0x7fffffff));
                if (newL == 0x7fffffff) { newL = U0 + 1; newU = L0 - 1; }
                maxScorePrime = fmaxf(maxScorePrime, __int_as_float(oMax ^ ((oMax >> 31) & 0x7fffffff)));

                if (!converged && k >= marker && k < nDiag - 1) {       // :585-595
                    const int start = newL - L0, len = newU - newL;
                    const short vI = sh.CI[g0][start], vD = sh.CD[g0][start], vS = sh.CS[c0][start];
                    unsigned bad = 0;
                    for (int t = 1 + lane; t <= len; t += 32) {
                        if (sh.CI[g0][start + t] != vI || sh.CD[g0][start + t] != vD) bad |= 1u;
                        if (sh.CS[c0][start + t] != vS) bad |= 2u;
                    }
                    bad = __reduce_or_sync(0xffffffffu, bad);
                    const int cS = (bad & 2u) ? -1 : unpackConv(vS);
                    const int cI = unpackConv(vI), cD = unpackConv(vD);
                    if (!(bad & 1u) && cI == cD && cI == cS && prevConvS == cS && cI != -1) {
                        converged = true;
                        convValue = prevConvS;
                        convScore = maxScorePrime;
                    }
                    prevConvS = cS;
                }

                L2 = L1; U2 = U1; L1 = L0; U1 = U0;
                L0 = max(newL, max(0, k + 2 - refLen));
                U0 = min(qryLen - 1, newU + 1);
                { const int t = c1; c1 = c0; c0 = c2; c2 = t; }
                maxScore = (maxScorePrime < 0.0f) ? 0.0f : maxScorePrime;
                lastK = k;
                if (converged && maxScore > convScore) { stopped = true; break; }
            }
            if (error != kStatusRetryWide) {
                cells += tileCells;
                diagonals += static_cast<unsigned long long>(lastK + 1);
            }
            if (error) { status = error; break; }
            const int nStored = min(lastK, marker) + 1;

            // ---- tile epilogue (:614-689): lane 0 walks the traceback, the warp copies the ops
            int newRefOff = 0, newQryOff = 0, isLast = 0, err = 0, tailLen = 0, tailOp = 0, w = 0;
            constexpr int opsCap = 2 * kMaxMarker + 16;
            if (lane == 0) {
                int convQry, convRef, startDiag, tbState;
                if (stopped || lastK >= marker) {
                    const int v = stopped ? convValue : unpackConv(sh.CS[lastK % 3][0]);
                    convQry = v & 0xFFFF;
                    tbState = static_cast<int8_t>((v >> 16) & 0xFFFF);
                    convRef = marker - convQry - ((tbState == 3) ? 1 : 0);
                    startDiag = (tbState == 3) ? nStored - 2 : nStored - 1;
                } else {
                    convQry = qryLen - 1; convRef = refLen - 1; startDiag = lastK; tbState = 0; isLast = 1;
                }
                if (convQry == (kDelBoundary & 0xFFFF)) { convQry = 0; convRef = marker; }
                else if (convQry == (kInsBoundary & 0xFFFF)) { convQry = marker; convRef = 0; }
                newRefOff = refOff + convRef; newQryOff = qryOff + convQry;
                if (pr.refLen - newRefOff < 0 || pr.qryLen - newQryOff < 0) err = 3;
                if (newRefOff == pr.refLen - 1 && newQryOff < pr.qryLen - 1) { tailLen = pr.qryLen - newQryOff - 1; tailOp = 1; isLast = 1; }
                if (newQryOff == pr.qryLen - 1 && newRefOff < pr.refLen - 1) { tailLen = pr.refLen - newRefOff - 1; tailOp = 2; isLast = 1; }
                if (newRefOff == pr.refLen - 1 && newQryOff == pr.qryLen - 1) isLast = 1;
                w = opsCap;
                if (!err) {
                    int kk = startDiag;
                    int row = static_cast<int16_t>(convQry), qi = row, ri = static_cast<int16_t>(convRef);
                    int state = static_cast<int8_t>(tbState) % 3;
                    const bool first = (tile == 0);
                    while (kk >= 0 && w > 0) {
                        if (kk >= 24) asm volatile("prefetch.global.L1 [%0];" ::"l"(tb + static_cast<size_t>(kk - 24) * kWarpCap + ((row - 12) & MASK)));
                        const int cell = tb[static_cast<size_t>(kk) * kWarpCap + (row & MASK)];
                        int dir;
                        if (state == 0) {
                            const int pp = cell & 3;
                            if (pp == 0) { dir = 0; }
                            else if (pp == 1) { dir = 1; state = (cell & 4) ? 1 : 0; }
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
            }
            __syncwarp();
            newRefOff = __shfl_sync(0xffffffffu, newRefOff, 0); newQryOff = __shfl_sync(0xffffffffu, newQryOff, 0);
            isLast = __shfl_sync(0xffffffffu, isLast, 0); err = __shfl_sync(0xffffffffu, err, 0);
            tailLen = __shfl_sync(0xffffffffu, tailLen, 0); tailOp = __shfl_sync(0xffffffffu, tailOp, 0); w = __shfl_sync(0xffffffffu, w, 0);
            const int nOps = opsCap - w;
            if (err) { status = err; break; }
            if (nOps + tailLen == 0) { status = 3; break; }
            {
                const int skip = (tile > 0) ? 1 : 0;
                const int nCopy = nOps - skip;
                const int8_t *src = sh.ops + w + skip;
                for (int t = lane; t < nCopy; t += 32) path[outPos + t] = src[t];
                for (int t = lane; t < tailLen; t += 32) path[outPos + nCopy + t] = static_cast<int8_t>(tailOp);
                outPos += nCopy + tailLen;
            }
            refOff = newRefOff; qryOff = newQryOff; lastTile = isLast != 0;
            ++tile;
            __syncwarp();
        }

        if (lane == 0) {
            if (status == kStatusRetryWide && a.overflowList != nullptr) a.overflowList[atomicAdd(a.overflowCount, 1)] = pairIdx;
            DevResult res;
            res.status = status;
            res.pathLen = (status && status != kStatusRetryWide) ? 0 : outPos;
            res.tiles = tile;
            res.pad = 0;
            res.cells = cells;
            res.diagonals = diagonals;
            res.resRefOff = refOff; res.resQryOff = qryOff;
            a.results[pairIdx] = res;
        }
        __syncwarp();
    }
}

int warpKernelBandCapacity() { return kWarpCap - 2; }
int warpKernelWindow() { return kWarpCap; }

cudaError_t launchTalcoWarp(int matClass, const TalcoArgs &args, int grid, cudaStream_t stream) {
    if (matClass == 1) talcoWarpKernel<1><<<grid, 32, 0, stream>>>(args);
    else talcoWarpKernel<0><<<grid, 32, 0, stream>>>(args);
    return cudaGetLastError();
}

int warpKernelMaxCtasPerSm(int matClass) {
    int n = 0;
    if (matClass == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, talcoWarpKernel<1>, 32, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, talcoWarpKernel<0>, 32, 0);
    return n;
}

} // namespace twl

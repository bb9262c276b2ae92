// talco_generic.cu — wide-band TALCO-XDrop kernel: one pair per CTA, wavefront state in shared memory (band up to
// `stateCap` cells) or in a per-CTA global scratch (any band up to fLen). It is the kernel that handles every band
// width the reference accepts; the register-resident kernel in talco_wavefront.cu takes the common narrow bands and
// hands pairs that overflow its capacity to this one.
//
// Semantics follow the reference CPU path cell for cell (src/TALCO-XDrop.cpp:233-689 Tile, :134-231 Traceback,
// :62-108 Align_freq), including the rotating wavefront buffers indexed by (row - L[k]) whose stale slots the
// convergence pointers read (SURVEY.md §7 hard part 2). Nothing here is derived from src/cuda.
#include "talco_score.cuh"
#include "twl_device.cuh"

namespace twl {

constexpr int kGenThreads = 256;
constexpr int kGenWarps = kGenThreads / 32;

__device__ __forceinline__ float warpMax(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, d));
    return v;
}

template <int P>
struct ColLoad;
template <>
struct ColLoad<6> {
    // de-interleaved nucleotide layout (twl_device.cuh): X = (c0 c1 c2 c3), Y = (c4 c5 gapOpen gapExtend)
    static __device__ __forceinline__ void load(const float *side, int n4, int j, float (&c)[6], float &gOp, float &gEx) {
        const float4 *v = reinterpret_cast<const float4 *>(side) + ntColIndex(j, n4);
        const float4 a = __ldg(v);
        const float4 b = __ldg(v + 4 * static_cast<long long>(n4));
        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y;
        gOp = b.z; gEx = b.w;
    }
};
template <>
struct ColLoad<22> {
    static __device__ __forceinline__ void load(const float *side, int /*n4*/, int j, float (&c)[22], float &gOp, float &gEx) {
        const float4 *v = reinterpret_cast<const float4 *>(side + static_cast<size_t>(j) * 24);
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const float4 a = __ldg(v + t);
            c[4 * t] = a.x; c[4 * t + 1] = a.y; c[4 * t + 2] = a.z; c[4 * t + 3] = a.w;
        }
        const float4 b = __ldg(v + 5);
        c[20] = b.x; c[21] = b.y; gOp = b.z; gEx = b.w;
    }
};

template <int P>
__device__ __forceinline__ float numeratorOf(const float (&r)[P], const float (&q)[P], const float *S, float g);
template <>
__device__ __forceinline__ float numeratorOf<6>(const float (&r)[6], const float (&q)[6], const float *S, float g) {
    return numeratorNt(r, q, S, g);
}
template <>
__device__ __forceinline__ float numeratorOf<22>(const float (&r)[22], const float (&q)[22], const float *S, float g) {
    return numeratorAa(r, q, S, g);
}

// Shared per-CTA bookkeeping for one tile.
struct TileShared {
    int ftrLen[kMaxMarker + 1];   // band width of diagonal k <= marker (ftr_length)
    int ftrLo[kMaxMarker + 1];    // lower row of diagonal k <= marker (ftr_lower_limit)
    int8_t ops[2 * kMaxMarker + 16];
    float redMax[2][kGenWarps];
    int redLo[2][kGenWarps];
    int redHi[2][kGenWarps];
    unsigned convMask[3];
    // tile epilogue broadcast
    int refOff, qryOff, lastTile, error, nOps, opsBegin, tailLen, tailOp;
    int work;
};

template <int P, bool GLOBAL_STATE>
__global__ void __launch_bounds__(kGenThreads) talcoGenericKernel(const TalcoArgs a) {
    constexpr int MS = (P - 1) * (P - 1);
    extern __shared__ __align__(16) unsigned char dynSmem[];
    __shared__ TileShared sh;
    __shared__ float sScore[MS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < MS; t += kGenThreads) sScore[t] = a.score[t];

    const int capPad = a.stateCap + 4;
    float *stateBase = GLOBAL_STATE ? (a.stateScratch + static_cast<size_t>(blockIdx.x) * a.stateStride)
                                    : reinterpret_cast<float *>(dynSmem);
    float *S[3] = {stateBase, stateBase + capPad, stateBase + 2 * capPad};
    float *I[2] = {stateBase + 3 * capPad, stateBase + 4 * capPad};
    float *D[2] = {stateBase + 5 * capPad, stateBase + 6 * capPad};
    int *CS[3] = {reinterpret_cast<int *>(stateBase + 7 * capPad), reinterpret_cast<int *>(stateBase + 8 * capPad),
                  reinterpret_cast<int *>(stateBase + 9 * capPad)};
    int *CI[2] = {reinterpret_cast<int *>(stateBase + 10 * capPad), reinterpret_cast<int *>(stateBase + 11 * capPad)};
    int *CD[2] = {reinterpret_cast<int *>(stateBase + 12 * capPad), reinterpret_cast<int *>(stateBase + 13 * capPad)};
    uint8_t *tb = a.tbScratch + static_cast<size_t>(blockIdx.x) * a.tbStride;
    const int marker = a.marker;

    for (;;) {
        __syncthreads();
        if (tid == 0) sh.work = atomicAdd(a.queue, 1);
        __syncthreads();
        const int work = sh.work;
        if (work >= *a.nWorkPtr) break;
        const int pairIdx = a.order[work];
        const DevPair pr = a.pairs[pairIdx];
        if (pr.refLen < 1 || pr.qryLen < 1) {   // an empty side (after gappy-column removal): nothing to align, the host emits the trivial path
            if (tid == 0) {
                DevResult res;
                res.status = kStatusEmptySide; res.pathLen = 0; res.tiles = 0; res.pad = 0; res.cells = 0; res.diagonals = 0; res.resRefOff = 0; res.resQryOff = 0;
                a.results[pairIdx] = res;
            }
            continue;
        }
        const float *refCols = a.prof + pr.refOff;
        const float *qryCols = a.prof + pr.qryOff;
        int8_t *path = a.paths + pr.alnOff;

        const float negInf = -static_cast<float>(2.0 * pr.xdrop + 1.0);   // TALCO-XDrop.cpp:252
        const float xdropF = static_cast<float>(pr.xdrop);
        const float denom = __fmul_rn(pr.refNum, pr.qryNum);              // :269
        int refOff = 0, qryOff = 0, tile = 0, outPos = 0, status = 0;
        unsigned long long cells = 0, diagonals = 0;
        bool lastTile = false;
        if (a.resume) {   // the previous kernel of the chain finished some tiles of this pair before its band capacity ran out
            const DevResult prev = a.results[pairIdx];
            if (prev.status == kStatusRetryWide) {
                refOff = prev.resRefOff; qryOff = prev.resQryOff; tile = prev.tiles; outPos = prev.pathLen;
                cells = prev.cells; diagonals = prev.diagonals;
            }
        }
        __syncthreads();   // every thread has read the previous result before thread 0 overwrites it at the end

        while (!lastTile) {                                                // Align_freq, :77-106
            const int refLen = pr.refLen - refOff, qryLen = pr.qryLen - qryOff;
            const int cap = min(pr.fLen, min(refLen, qryLen));             // :258
            const unsigned long long cellsAtTile = cells, diagAtTile = diagonals;
            // wavefront state init, :301-308
            for (int t = tid; t < capPad; t += kGenThreads) {
                S[0][t] = S[1][t] = S[2][t] = -1.0f;
                I[0][t] = I[1][t] = D[0][t] = D[1][t] = -1.0f;
                CS[0][t] = CS[1][t] = CS[2][t] = -1;
                CI[0][t] = CI[1][t] = kInsBoundary;
                CD[0][t] = CD[1][t] = kDelBoundary;
            }
            if (tid < 3) sh.convMask[tid] = 3u;
            __syncthreads();

            int L0 = 0, U0 = 0, L1 = 2, U1 = -2, L2 = 1, U2 = -1;         // L={0,1,2} U={0,-1,-2} seen from k = 0
            float maxScore = 0.0f, maxScorePrime = negInf, convScore = 0.0f;
            bool converged = false, stopped = false;
            int convValue = 0, prevConvS = -1, lastK = 0, tbTotal = 0, nStored = 0;
            int error = 0;
            const int nDiag = refLen + qryLen - 1;

            for (int k = 0; k < nDiag; ++k) {
                const int c0 = k % 3, c1 = (k + 2) % 3, c2 = (k + 1) % 3, g0 = k & 1, g1 = g0 ^ 1;
                if (L0 >= U0 + 1) { error = 1; break; }                    // :323-329
                const int width = U0 - L0 + 1;
                if (width > cap) { error = 2; break; }                     // :331-338
                if (width > a.stateCap) { error = kStatusRetryWide; break; }
                if (k <= marker) {
                    if (tid == 0) { sh.ftrLen[k] = width; sh.ftrLo[k] = L0; }
                    nStored = k + 1;
                }
                cells += static_cast<unsigned long long>(width);
                diagonals += 1;
                const float pruneBelow = __fsub_rn(maxScore, xdropF);      // :495
                float myMax = negInf;
                int myLo = 0x7fffffff, myHi = -0x7fffffff;

                for (int i = L0 + tid; i <= U0; i += kGenThreads) {
                    const int j = k - i;
                    const int off = i - L0, offDiag = i - 1 - L2, offUp = i - L1, offLeft = offUp - 1;
                    float match = negInf, insOpen = negInf, insExt = negInf, delOpen = negInf, delExt = negInf;
                    float r[P], q[P], gOpR, gExR, gOpQ, gExQ;
                    ColLoad<P>::load(refCols, pr.refN4, refOff + j, r, gOpR, gExR);
                    ColLoad<P>::load(qryCols, pr.qryN4, qryOff + i, q, gOpQ, gExQ);
                    const bool diagIn = offDiag >= 0 && offDiag <= U2 - L2;
                    const bool onEdge0 = (tile == 0) && (i == 0 || j == 0);
                    if (k == 0 || diagIn || onEdge0) {
                        const float sim = __fdiv_rn(numeratorOf<P>(r, q, sScore, pr.gapChar), denom);
                        if (onEdge0) {
                            if (i == 0 && j == 0) match = sim;
                            else match = __fmaf_rn(a.gapExtend, static_cast<float>(max(0, max(refOff + j, qryOff + i) - 1)),
                                                   __fadd_rn(sim, a.gapOpen));                        // :448
                        } else if (offDiag < 0) match = sim;
                        else match = __fadd_rn(S[c2][offDiag], sim);
                    }
                    if (offUp >= 0 && offUp <= U1 - L1) {
                        delOpen = __fadd_rn(S[c1][offUp], gOpR);
                        delExt = __fadd_rn(D[g1][offUp], gExR);
                    }
                    if (offLeft >= 0 && offLeft <= U1 - L1) {
                        insOpen = __fadd_rn(S[c1][offLeft], gOpQ);
                        insExt = __fadd_rn(I[g1][offLeft], gExQ);
                    }
                    const bool insFromIns = insExt >= insOpen, delFromDel = delExt >= delOpen;
                    const float insBest = insFromIns ? insExt : insOpen, delBest = delFromDel ? delExt : delOpen;
                    int ptr;
                    float s;
                    if (match >= insBest) {
                        if (match >= delBest) { s = match; ptr = 0; }
                        else { s = delBest; ptr = 2; }
                    } else if (insBest > delBest) { s = insBest; ptr = 1; }
                    else { s = delBest; ptr = 2; }
                    if (s < pruneBelow) s = negInf;
                    I[g0][off] = insBest;
                    D[g0][off] = delBest;
                    S[c0][off] = s;
                    myMax = fmaxf(myMax, s);
                    if (s > negInf) { myLo = min(myLo, i); myHi = max(myHi, i); }

                    if (k == marker - 1) {
                        CS[c0][off] = (3 << 16) | (i & 0xFFFF);
                    } else if (k == marker) {
                        CS[c0][off] = (i & 0xFFFF);
                        CI[g0][off] = (1 << 16) | (i & 0xFFFF);
                        CD[g0][off] = (2 << 16) | (i & 0xFFFF);
                    } else if (k > marker) {                               // :527-547
                        int ci, cd;
                        if (insFromIns) ci = (offLeft >= 0) ? CI[g1][offLeft] : kInsBoundary;
                        else { const int v = (offLeft >= 0) ? CS[c1][offLeft] : -1; ci = (v != -1) ? v : kInsBoundary; }
                        if (delFromDel) cd = (offUp >= 0) ? CD[g1][offUp] : kDelBoundary;
                        else { const int v = (offUp >= 0) ? CS[c1][offUp] : -1; cd = (v != -1) ? v : kDelBoundary; }
                        CI[g0][off] = ci;
                        CD[g0][off] = cd;
                        CS[c0][off] = (ptr == 0) ? ((offDiag >= 0) ? CS[c2][offDiag] : -1) : ((ptr == 1) ? ci : cd);
                    }
                    if (k <= marker) tb[tbTotal + off] = static_cast<uint8_t>(ptr | (insFromIns ? 4 : 0) | (delFromDel ? 8 : 0));
                }
                if (k <= marker) tbTotal += width;

                // block reduction of (max score, first live row, last live row)
                myMax = warpMax(myMax);
                myLo = __reduce_min_sync(0xffffffffu, myLo);
                myHi = __reduce_max_sync(0xffffffffu, myHi);
                if (lane == 0) { sh.redMax[g0][warp] = myMax; sh.redLo[g0][warp] = myLo; sh.redHi[g0][warp] = myHi; }
                __syncthreads();
                float diagMax = sh.redMax[g0][0];
                int newL = sh.redLo[g0][0], newU = sh.redHi[g0][0];
#pragma unroll
                for (int w = 1; w < kGenWarps; ++w) {
                    diagMax = fmaxf(diagMax, sh.redMax[g0][w]);
                    newL = min(newL, sh.redLo[g0][w]);
                    newU = max(newU, sh.redHi[g0][w]);
                }
                if (newL == 0x7fffffff) { newL = U0 + 1; newU = L0 - 1; }  // every cell pruned, :563-583
                maxScorePrime = fmaxf(maxScorePrime, diagMax);

                if (!converged && k >= marker && k < nDiag - 1) {          // :585-595 (cannot fire below the marker)
                    const int start = newL - L0, len = newU - newL;
                    const int vI = CI[g0][start], vD = CD[g0][start], vS = CS[c0][start];
                    unsigned bad = 0;
                    for (int t = 1 + tid; t <= len; t += kGenThreads) {
                        if (CI[g0][start + t] != vI || CD[g0][start + t] != vD) bad |= 1u;
                        if (CS[c0][start + t] != vS) bad |= 2u;
                    }
                    if (bad) atomicAnd(&sh.convMask[c0], ~bad);
                    if (tid == 0) sh.convMask[c2] = 3u;                    // slot of diagonal k+1; last read in k-2
                    __syncthreads();
                    const unsigned ok = sh.convMask[c0];
                    const int cS = (ok & 2u) ? vS : -1;
                    const bool idOk = (ok & 1u) != 0;
                    if (idOk && vI == vD && vI == cS && prevConvS == cS && vI != -1) {
                        converged = true;
                        convValue = prevConvS;
                        convScore = maxScorePrime;
                    }
                    prevConvS = cS;
                }

                const int nextL = max(newL, max(0, k + 2 - refLen));        // :597-604
                const int nextU = min(qryLen - 1, newU + 1);
                L2 = L1; U2 = U1; L1 = L0; U1 = U0; L0 = nextL; U0 = nextU;
                maxScore = (maxScorePrime < 0.0f) ? 0.0f : maxScorePrime;   // :607
                lastK = k;
                if (converged && maxScore > convScore) { stopped = true; break; }
            }

            if (error) {
                if (error == kStatusRetryWide) { cells = cellsAtTile; diagonals = diagAtTile; }   // the wide variant redoes this tile
                status = error;
                break;
            }

            // ---- tile epilogue by one thread: traceback start (:614-652), tails (:671-679), Traceback (:134-231)
            __syncthreads();
            if (tid == 0) {
                int convQry, convRef, startDiag, tbState, isLast = 0;
                if (stopped || lastK >= marker) {
                    const int v = stopped ? convValue : CS[lastK % 3][0];
                    convQry = v & 0xFFFF;
                    tbState = static_cast<int8_t>((v >> 16) & 0xFFFF);
                    convRef = marker - convQry - ((tbState == 3) ? 1 : 0);
                    startDiag = (tbState == 3) ? nStored - 2 : nStored - 1;
                } else {
                    convQry = qryLen - 1;
                    convRef = refLen - 1;
                    startDiag = lastK;
                    tbState = 0;
                    isLast = 1;
                }
                if (convQry == (kDelBoundary & 0xFFFF)) { convQry = 0; convRef = marker; }
                else if (convQry == (kInsBoundary & 0xFFFF)) { convQry = marker; convRef = 0; }
                const int newRefOff = refOff + convRef, newQryOff = qryOff + convQry;
                int err = 0, tailLen = 0, tailOp = 0;
                if (pr.refLen - newRefOff < 0 || pr.qryLen - newQryOff < 0) err = 3;
                if (newRefOff == pr.refLen - 1 && newQryOff < pr.qryLen - 1) { tailLen = pr.qryLen - newQryOff - 1; tailOp = 1; isLast = 1; }
                if (newQryOff == pr.qryLen - 1 && newRefOff < pr.refLen - 1) { tailLen = pr.refLen - newRefOff - 1; tailOp = 2; isLast = 1; }
                if (newRefOff == pr.refLen - 1 && newQryOff == pr.qryLen - 1) isLast = 1;

                // traceback, writing ops from the back of sh.ops so that they read forward
                constexpr int opsCap = 2 * kMaxMarker + 16;
                int w = opsCap;
                if (!err) {
                    int kk = startDiag;
                    int row = static_cast<int16_t>(convQry), qi = row, ri = static_cast<int16_t>(convRef);
                    int state = static_cast<int8_t>(tbState) % 3;
                    // base address of diagonal startDiag inside the flat tb
                    int base = tbTotal;
                    for (int d = nStored - 1; d >= startDiag && d >= 0; --d) base -= sh.ftrLen[d];
                    const bool first = (tile == 0);
                    while (kk >= 0 && w > 0) {
                        int addr = base + (row - sh.ftrLo[kk]);
                        addr = max(0, min(addr, tbTotal - 1));
                        const int cell = tb[addr];
                        int dir;
                        if (state == 0) {
                            const int p = cell & 3;
                            if (p == 0) { dir = 0; }
                            else if (p == 1) { dir = 1; state = (cell & 4) ? 1 : 0; }
                            else { dir = 2; state = (cell & 8) ? 2 : 0; }
                        } else if (state == 1) { dir = 1; state = (cell & 4) ? 1 : 0; }
                        else { dir = 2; state = (cell & 8) ? 2 : 0; }
                        if (dir == 0) {
                            if (kk >= 1) base -= sh.ftrLen[kk - 1];
                            if (kk >= 2) base -= sh.ftrLen[kk - 2];
                            kk -= 2; row -= 1; qi--; ri--;
                        } else if (dir == 1) {
                            if (kk >= 1) base -= sh.ftrLen[kk - 1];
                            kk -= 1; row -= 1; qi--;
                        } else {
                            if (kk >= 1) base -= sh.ftrLen[kk - 1];
                            kk -= 1; ri--;
                        }
                        sh.ops[--w] = static_cast<int8_t>(dir);
                        if (first && (ri < 0 || qi < 0)) break;
                    }
                    if (first) {                                           // :221-230
                        while (ri > -1 && w > 0) { sh.ops[--w] = 2; ri--; }
                        while (qi > -1 && w > 0) { sh.ops[--w] = 1; qi--; }
                    }
                }
                sh.refOff = newRefOff; sh.qryOff = newQryOff; sh.lastTile = isLast; sh.error = err;
                sh.opsBegin = w; sh.nOps = opsCap - w; sh.tailLen = tailLen; sh.tailOp = tailOp;
            }
            __syncthreads();
            if (sh.error) { status = sh.error; break; }
            if (sh.nOps + sh.tailLen == 0) { status = 3; break; }          // empty tile path: Align_freq clears and returns
            {
                const int skip = (tile > 0) ? 1 : 0;                       // :99 drop the re-aligned tile origin
                const int nCopy = sh.nOps - skip;
                const int8_t *src = sh.ops + sh.opsBegin + skip;
                for (int t = tid; t < nCopy; t += kGenThreads) path[outPos + t] = src[t];
                const int8_t tailOp = static_cast<int8_t>(sh.tailOp);
                for (int t = tid; t < sh.tailLen; t += kGenThreads) path[outPos + nCopy + t] = tailOp;
                outPos += nCopy + sh.tailLen;
            }
            refOff = sh.refOff; qryOff = sh.qryOff; lastTile = sh.lastTile != 0;
            ++tile;
            __syncthreads();
        }

        if (tid == 0 && status == kStatusRetryWide && a.overflowList != nullptr) {
            a.overflowList[atomicAdd(a.overflowCount, 1)] = pairIdx;
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
            if (status == kStatusRetryWide) res.pathLen = outPos;
            a.results[pairIdx] = res;
        }
    }
}

// ---- launch helpers used by twl_api.cu ----------------------------------------------------------------------
size_t genericStateWords(int stateCap) { return static_cast<size_t>(14) * (stateCap + 4); }

cudaError_t launchTalcoGeneric(int P, bool globalState, const TalcoArgs &args, int grid, size_t dynSmemBytes, cudaStream_t stream) {
    if (P == 6) {
        if (globalState) talcoGenericKernel<6, true><<<grid, kGenThreads, 0, stream>>>(args);
        else {
            cudaFuncSetAttribute(talcoGenericKernel<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dynSmemBytes));
            talcoGenericKernel<6, false><<<grid, kGenThreads, dynSmemBytes, stream>>>(args);
        }
    } else {
        if (globalState) talcoGenericKernel<22, true><<<grid, kGenThreads, 0, stream>>>(args);
        else {
            cudaFuncSetAttribute(talcoGenericKernel<22, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(dynSmemBytes));
            talcoGenericKernel<22, false><<<grid, kGenThreads, dynSmemBytes, stream>>>(args);
        }
    }
    return cudaGetLastError();
}

int genericThreads() { return kGenThreads; }

} // namespace twl

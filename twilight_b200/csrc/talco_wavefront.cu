// talco_wavefront.cu — the fast TALCO-XDrop kernel for nucleotide profiles (P = 6): one pair per CTA, the anti-diagonal
// wavefront lives in REGISTERS.
//
// Mapping. A CTA of NT threads owns W = 4*NT consecutive query rows at a time. Row i is handled by "slot"
// rho = i mod W, i.e. thread rho/4, register slot rho%4; as the live band [L,U] of the X-drop recurrence slides
// along the query, a slot whose row fell below L is re-assigned to row + W. Per slot the thread keeps, in registers,
// the query column (6 counts + 2 position-specific gap penalties) and the H / I / D scores of the previous two
// anti-diagonals, so per cell and per diagonal the only memory traffic is one 32-byte reference column (two LDG.128
// that hit L1; the columns needed a few diagonals ahead are prefetched) and one traceback byte (the four slots of a
// thread store one coalesced 32-bit word). Row-neighbour values come from the thread's own registers for slots 1..3
// and from one warp shuffle for slot 0; warps exchange their edge values and the per-diagonal reduction
// (running maximum for the X-drop rule, first / last live row) through shared memory with ONE barrier per diagonal.
//
// What is kept bit-identical to the reference CPU path (src/TALCO-XDrop.cpp:233-689): the float operation order of
// the score (talco_score.cuh), tie rules, the prune rule against the previous diagonal's maximum, the dead-end
// trimming, the convergence pointers INCLUDING the reference's rotating buffers indexed by (row - L[k]) — those
// live in shared memory with the reference's indexing so that the stale slots the reference reads are reproduced —
// the tile stop rule, the traceback start cell and the per-tile path concatenation of Align_freq (:62-108).
//
// A band wider than W cannot be held; the pair is then appended to an overflow list and re-run by a wider
// instantiation or by the generic kernel (talco_generic.cu).
#include "talco_score.cuh"
#include "twl_device.cuh"

namespace twl {

constexpr int kSlots = 4;

struct WaveShared {
    int4 red[2][8];            // per warp: (max score as ordered int, first live row, last live row, -)
    float2 edge[2][8];         // per warp: H and I of the warp's last slot (row-neighbour of the next warp's first slot)
    unsigned convMask[3];
    int8_t ops[2 * kMaxMarker + 16];
    int refOff, qryOff, lastTile, error, nOps, opsBegin, tailLen, tailOp;
    int work;
};

__device__ __forceinline__ int orderedInt(float f) {
    const int b = __float_as_int(f);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float orderedFloat(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7fffffff)); }

__device__ __forceinline__ void prefetchL1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int NT>
__global__ void __launch_bounds__(NT) talcoWavefrontKernel(const TalcoArgs a) {
    constexpr int W = NT * kSlots;
    constexpr int NW = NT / 32;
    constexpr int PW = 8;
    constexpr int CW = W + 4;                       // convergence arrays, reference indexing (row - L[k]) plus padding
    __shared__ WaveShared sh;
    __shared__ int sCS[3][CW], sCI[2][CW], sCD[2][CW];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t *tb = a.tbScratch + static_cast<size_t>(blockIdx.x) * a.tbStride;   // tb[k][rho], row stride W
    const int marker = a.marker;
    const int rho0 = tid * kSlots;

    for (;;) {
        __syncthreads();
        if (tid == 0) sh.work = atomicAdd(a.queue, 1);
        __syncthreads();
        const int work = sh.work;
        if (work >= *a.nWorkPtr) break;
        const int pairIdx = a.order[work];
        const DevPair pr = a.pairs[pairIdx];
        const float *refCols = a.prof + pr.refOff;
        const float *qryCols = a.prof + pr.qryOff;
        int8_t *path = a.paths + pr.alnOff;

        const float negInf = -static_cast<float>(2.0 * pr.xdrop + 1.0);
        const float xdropF = static_cast<float>(pr.xdrop);
        const float denom = __fmul_rn(pr.refNum, pr.qryNum);
        const bool unitDenom = (denom == 1.0f);
        const float gapChar = pr.gapChar;
        int refOff = 0, qryOff = 0, tile = 0, outPos = 0, status = 0;
        unsigned long long cells = 0, diagonals = 0;
        bool lastTile = false;

        while (!lastTile) {
            const int refLen = pr.refLen - refOff, qryLen = pr.qryLen - qryOff;
            const int cap = min(pr.fLen, min(refLen, qryLen));
            for (int t = tid; t < CW; t += NT) {
                sCS[0][t] = sCS[1][t] = sCS[2][t] = -1;
                sCI[0][t] = sCI[1][t] = kInsBoundary;
                sCD[0][t] = sCD[1][t] = kDelBoundary;
            }
            if (tid < 3) sh.convMask[tid] = 3u;
            __syncthreads();

            // per-slot register state
            float h1[kSlots], i1[kSlots], d1[kSlots], h2[kSlots];      // H,I,D of diagonal k-1 and H of k-2 for the slot's row
            float q[kSlots][6], gOpQ[kSlots], gExQ[kSlots];
            int rowCur[kSlots];
#pragma unroll
            for (int c = 0; c < kSlots; ++c) { h1[c] = i1[c] = d1[c] = h2[c] = -1.0f; rowCur[c] = -1; gOpQ[c] = gExQ[c] = 0.f;
#pragma unroll
                for (int t = 0; t < 6; ++t) q[c][t] = 0.f; }
            float leftHPrev = -1.0f;                                   // H[k-2] of the row below slot 0

            int L0 = 0, U0 = 0, L1 = 2, U1 = -2, L2 = 1, U2 = -1;
            float maxScore = 0.0f, maxScorePrime = negInf, convScore = 0.0f;
            bool converged = false, stopped = false;
            int convValue = 0, prevConvS = -1, lastK = 0, nStored = 0, error = 0;
            const int nDiag = refLen + qryLen - 1;
            const float *refTile = refCols + static_cast<size_t>(refOff) * PW;
            const float *qryTile = qryCols + static_cast<size_t>(qryOff) * PW;

            for (int k = 0; k < nDiag; ++k) {
                const int c0 = k % 3, c1 = (k + 2) % 3, c2 = (k + 1) % 3, g0 = k & 1, g1 = g0 ^ 1;
                if (L0 >= U0 + 1) { error = 1; break; }
                const int width = U0 - L0 + 1;
                if (width > cap) { error = 2; break; }
                if (width > W) { error = kStatusRetryWide; break; }
                if (k <= marker) nStored = k + 1;
                cells += static_cast<unsigned long long>(width);
                diagonals += 1;
                const float pruneBelow = __fsub_rn(maxScore, xdropF);

                if (tid == 0) {   // warm L1 for the columns the band edges will touch a few diagonals from now
                    const int jAhead = min(refLen - 1, k - L0 + 12);
                    const int iAhead = min(qryLen - 1, L0 + W + 8);
                    prefetchL1(refTile + static_cast<size_t>(jAhead) * PW);
                    prefetchL1(qryTile + static_cast<size_t>(iAhead) * PW);
                }

                // row-neighbour of slot 0: last slot of the previous thread (previous warp through shared memory)
                float nbH = __shfl_up_sync(0xffffffffu, h1[kSlots - 1], 1);
                float nbI = __shfl_up_sync(0xffffffffu, i1[kSlots - 1], 1);
                if (lane == 0) {
                    const float2 e = sh.edge[g1][(warp + NW - 1) % NW];
                    nbH = e.x; nbI = e.y;
                }

                float myMax = negInf;
                int myLo = 0x7fffffff, myHi = -0x7fffffff;
                float nh[kSlots], ni[kSlots], nd[kSlots];
                unsigned tbWord = 0;
                unsigned liveBits = 0;
                int cs[kSlots], ci[kSlots], cd[kSlots];

#pragma unroll
                for (int c = 0; c < kSlots; ++c) {
                    const int rho = rho0 + c;
                    const int i = L0 + ((rho - L0) & (W - 1));
                    nh[c] = h1[c]; ni[c] = i1[c]; nd[c] = d1[c];
                    cs[c] = ci[c] = cd[c] = 0;
                    if (i != rowCur[c]) {                               // slot re-assigned: fetch its query column
                        rowCur[c] = i;
                        if (i < qryLen) {
                            const float4 x = __ldg(reinterpret_cast<const float4 *>(qryTile + static_cast<size_t>(i) * PW));
                            const float4 y = __ldg(reinterpret_cast<const float4 *>(qryTile + static_cast<size_t>(i) * PW) + 1);
                            q[c][0] = x.x; q[c][1] = x.y; q[c][2] = x.z; q[c][3] = x.w; q[c][4] = y.x; q[c][5] = y.y;
                            gOpQ[c] = y.z; gExQ[c] = y.w;
                        }
                    }
                    if (i <= U0) {
                        const int j = k - i;
                        const float leftH = (c == 0) ? nbH : h1[(c + kSlots - 1) % kSlots];
                        const float leftI = (c == 0) ? nbI : i1[(c + kSlots - 1) % kSlots];
                        const float diagH = (c == 0) ? leftHPrev : h2[(c + kSlots - 1) % kSlots];
                        const float4 x = __ldg(reinterpret_cast<const float4 *>(refTile + static_cast<size_t>(j) * PW));
                        const float4 y = __ldg(reinterpret_cast<const float4 *>(refTile + static_cast<size_t>(j) * PW) + 1);
                        const float r[6] = {x.x, x.y, x.z, x.w, y.x, y.y};
                        const float gOpR = y.z, gExR = y.w;
                        const bool upIn = (i <= U1);                    // i >= L0 >= L1 always
                        const bool leftIn = (i > L1);                   // i-1 <= U0-1 <= U1 always
                        const bool diagIn = (i - 1 >= L2) && (i - 1 <= U2);
                        const bool onEdge0 = (tile == 0) && (i == 0 || j == 0);
                        float match = negInf;
                        if (k == 0 || diagIn || onEdge0) {
                            const float num = numeratorNt(r, q[c], a.scoreNt, gapChar);
                            const float sim = unitDenom ? num : __fdiv_rn(num, denom);
                            if (onEdge0) {
                                if (i == 0 && j == 0) match = sim;
                                else match = __fmaf_rn(a.gapExtend, static_cast<float>(max(0, max(refOff + j, qryOff + i) - 1)), __fadd_rn(sim, a.gapOpen));
                            } else if (!diagIn) match = sim;           // k == 0
                            else match = __fadd_rn(diagH, sim);
                        }
                        const float delOpen = upIn ? __fadd_rn(h1[c], gOpR) : negInf;
                        const float delExt = upIn ? __fadd_rn(d1[c], gExR) : negInf;
                        const float insOpen = leftIn ? __fadd_rn(leftH, gOpQ[c]) : negInf;
                        const float insExt = leftIn ? __fadd_rn(leftI, gExQ[c]) : negInf;
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
                        nh[c] = s; ni[c] = insBest; nd[c] = delBest;
                        myMax = fmaxf(myMax, s);
                        if (s > negInf) { myLo = min(myLo, i); myHi = max(myHi, i); }
                        liveBits |= 1u << c;
                        tbWord |= static_cast<unsigned>(ptr | (insFromIns ? 4 : 0) | (delFromDel ? 8 : 0)) << (8 * c);

                        if (k >= marker - 1) {                          // convergence pointers, reference indexing
                            const int off = i - L0, offDiag = i - 1 - L2, offUp = i - L1, offLeft = offUp - 1;
                            if (k == marker - 1) {
                                cs[c] = (3 << 16) | (i & 0xFFFF);
                                sCS[c0][off] = cs[c];
                            } else if (k == marker) {
                                cs[c] = (i & 0xFFFF); ci[c] = (1 << 16) | (i & 0xFFFF); cd[c] = (2 << 16) | (i & 0xFFFF);
                                sCS[c0][off] = cs[c]; sCI[g0][off] = ci[c]; sCD[g0][off] = cd[c];
                            } else {
                                int vi, vd;
                                if (insFromIns) vi = (offLeft >= 0) ? sCI[g1][offLeft] : kInsBoundary;
                                else { const int v = (offLeft >= 0) ? sCS[c1][offLeft] : -1; vi = (v != -1) ? v : kInsBoundary; }
                                if (delFromDel) vd = (offUp >= 0) ? sCD[g1][offUp] : kDelBoundary;
                                else { const int v = (offUp >= 0) ? sCS[c1][offUp] : -1; vd = (v != -1) ? v : kDelBoundary; }
                                const int vs = (ptr == 0) ? ((offDiag >= 0) ? sCS[c2][offDiag] : -1) : ((ptr == 1) ? vi : vd);
                                ci[c] = vi; cd[c] = vd; cs[c] = vs;
                                sCI[g0][off] = vi; sCD[g0][off] = vd; sCS[c0][off] = vs;
                            }
                        }
                    }
                }
                if (k <= marker && liveBits) *reinterpret_cast<unsigned *>(tb + static_cast<size_t>(k) * W + rho0) = tbWord;

                // rotate the register wavefront
                leftHPrev = nbH;
#pragma unroll
                for (int c = 0; c < kSlots; ++c) { h2[c] = h1[c]; h1[c] = nh[c]; i1[c] = ni[c]; d1[c] = nd[c]; }

                // one barrier per diagonal: publish the warp's edge values and its reduction
                const int wMax = __reduce_max_sync(0xffffffffu, orderedInt(myMax));
                const int wLo = __reduce_min_sync(0xffffffffu, myLo);
                const int wHi = __reduce_max_sync(0xffffffffu, myHi);
                if (lane == 31) sh.edge[g0][warp] = make_float2(h1[kSlots - 1], i1[kSlots - 1]);
                if (lane == 0) sh.red[g0][warp] = make_int4(wMax, wLo, wHi, 0);
                __syncthreads();
                int oMax = sh.red[g0][0].x, newL = sh.red[g0][0].y, newU = sh.red[g0][0].z;
#pragma unroll
                for (int w = 1; w < NW; ++w) {
                    const int4 t = sh.red[g0][w];
                    oMax = max(oMax, t.x); newL = min(newL, t.y); newU = max(newU, t.z);
                }
                if (newL == 0x7fffffff) { newL = U0 + 1; newU = L0 - 1; }
                maxScorePrime = fmaxf(maxScorePrime, orderedFloat(oMax));

                if (!converged && k >= marker && k < nDiag - 1) {       // :585-595
                    const int start = newL - L0;
                    const int vI = sCI[g0][start], vD = sCD[g0][start], vS = sCS[c0][start];
                    unsigned bad = 0;
#pragma unroll
                    for (int c = 0; c < kSlots; ++c) {
                        if (liveBits & (1u << c)) {
                            const int i = rowCur[c];
                            if (i > newL && i <= newU) {
                                if (ci[c] != vI || cd[c] != vD) bad |= 1u;
                                if (cs[c] != vS) bad |= 2u;
                            }
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

                const int nextL = max(newL, max(0, k + 2 - refLen));
                const int nextU = min(qryLen - 1, newU + 1);
                L2 = L1; U2 = U1; L1 = L0; U1 = U0; L0 = nextL; U0 = nextU;
                maxScore = (maxScorePrime < 0.0f) ? 0.0f : maxScorePrime;
                lastK = k;
                if (converged && maxScore > convScore) { stopped = true; break; }
            }

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
                        const int cell = tb[static_cast<size_t>(kk) * W + (row & (W - 1))];
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
            __syncthreads();
        }

        if (tid == 0) {
            if (status == kStatusRetryWide && a.overflowList != nullptr) a.overflowList[atomicAdd(a.overflowCount, 1)] = pairIdx;
            DevResult res;
            res.status = status;
            res.pathLen = status ? 0 : outPos;
            res.tiles = tile;
            res.pad = 0;
            res.cells = cells;
            res.diagonals = diagonals;
            a.results[pairIdx] = res;
        }
    }
}

int wavefrontBandCapacity(int threads) { return threads * kSlots; }

cudaError_t launchTalcoWavefront(int threads, const TalcoArgs &args, int grid, cudaStream_t stream) {
    switch (threads) {
    case 64: talcoWavefrontKernel<64><<<grid, 64, 0, stream>>>(args); break;
    case 128: talcoWavefrontKernel<128><<<grid, 128, 0, stream>>>(args); break;
    case 256: talcoWavefrontKernel<256><<<grid, 256, 0, stream>>>(args); break;
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int wavefrontMaxCtasPerSm(int threads) {
    int n = 0;
    switch (threads) {
    case 64: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, talcoWavefrontKernel<64>, 64, 0); break;
    case 128: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, talcoWavefrontKernel<128>, 128, 0); break;
    case 256: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, talcoWavefrontKernel<256>, 256, 0); break;
    default: break;
    }
    return n;
}

} // namespace twl

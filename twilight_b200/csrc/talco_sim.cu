// talco_sim.cu — protein path, step 1: the column-pair similarity of EVERY cell of every pair of the batch, computed by a
// dependency-free kernel before the recurrence runs.
//
// Why. For proteins the similarity numerator is a 21 x 21 contraction written out term by term in the reference
// (src/TALCO-XDrop.cpp:405-440): ~50 float operations per non-zero reference letter and cell, against ~25 for the whole
// three-state recurrence. Inside the anti-diagonal kernel that work sits behind one CTA barrier per diagonal and runs at a few
// per cent of the FP32 pipe. It does not depend on the DP state at all — only on (query column i, reference column j) — and with
// 5 x BLOSUM62 scores the x-drop band is nearly the whole anti-diagonal, so the full |qry| x |ref| matrix is what
// the recurrence reads anyway. Here it is computed at full occupancy; the recurrence then runs in the register-resident wavefront
// kernel (talco_wavefront.cu, SC = 1) and reads one float per cell.
//
// Mapping. A warp owns 32 query rows (one per lane, the row's 22 counts in registers) and walks 32 reference columns; the column's
// counts, its non-zero-letter mask and the score-matrix row are warp-uniform (shared-memory broadcasts), so the loop over the
// column's non-zero letters does not diverge and every lane does useful work. Results of a 32 x 32 tile are transposed through
// shared memory and written by anti-diagonals, because the matrix is stored anti-diagonal-major ([i + j][i], see DevSim): on one
// anti-diagonal the rows of a wavefront thread are adjacent words and the rows of a warp one contiguous run.
//
// Bit-exactness: the numerator follows numeratorAa (talco_score.cuh) operation for operation — zero reference letters skipped (they
// contribute exact zeros), the rest in the reference's order — and the quotient is the IEEE quotient (exactDiv).
#include "talco_score.cuh"
#include "twl_device.cuh"

namespace twl {

constexpr int kSimWarps = 2;                 // warps (32-row strips) per block
constexpr int kSimCols = 32;                 // reference columns per tile
constexpr int kSimColTiles = 4;              // consecutive column tiles a block walks with the same query rows

__global__ void __launch_bounds__(32 * kSimWarps) simMatrixAaKernel(const float *prof, const DevPair *pairs, const int *order, const DevSim *simInfo,
                                                                    float *sim, const float *score) {
    __shared__ __align__(16) float sS[21 * 24];       // score-matrix rows padded to 24 floats: a row is five 128-bit broadcasts + one word
    __shared__ float sR[kSimCols][24];
    __shared__ unsigned sLive[kSimCols];
    __shared__ float sOut[kSimWarps][kSimCols][34];   // [column][row], row stride 34: conflict-free both by column and by anti-diagonal

    const int pairIdx = order[blockIdx.y];
    const DevPair pr = pairs[pairIdx];
    if (pr.refLen < 1 || pr.qryLen < 1) return;
    const int nColTiles = (pr.refLen + kSimCols - 1) / kSimCols;
    const int nColGroups = (nColTiles + kSimColTiles - 1) / kSimColTiles;
    const int nRowTiles = (pr.qryLen + 32 * kSimWarps - 1) / (32 * kSimWarps);
    if (static_cast<int>(blockIdx.x) >= nColGroups * nRowTiles) return;
    const int colGroup = blockIdx.x % nColGroups, rowTile = blockIdx.x / nColGroups;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = (rowTile * kSimWarps + warp) * 32;
    const DevSim si = simInfo[pairIdx];
    const float *refCols = prof + pr.refOff, *qryCols = prof + pr.qryOff;

    for (int t = tid; t < 21 * 24; t += 32 * kSimWarps) {
        const int l = t / 24, m = t - l * 24;
        sS[t] = (m < 21) ? score[l * 21 + m] : 0.0f;
    }
    const int i = i0 + lane;
    float q[22];
    {
        const float4 *v = reinterpret_cast<const float4 *>(qryCols + static_cast<size_t>(min(i, pr.qryLen - 1)) * 24);
#pragma unroll
        for (int t = 0; t < 5; ++t) {
            const float4 x = __ldg(v + t);
            q[4 * t] = x.x; q[4 * t + 1] = x.y; q[4 * t + 2] = x.z; q[4 * t + 3] = x.w;
        }
        const float4 y = __ldg(v + 5);
        q[20] = y.x; q[21] = y.y;
    }
    const float denom = __fmul_rn(pr.refNum, pr.qryNum);              // TALCO-XDrop.cpp:269
    const float rcp = __fdiv_rn(1.0f, denom);
    const bool ieeeDiv = (__float_as_int(denom) & 0x7fffff) == 0x7fffff;   // the reciprocal method is exact except for these denominators
    const float g = pr.gapChar;

    for (int ct = 0; ct < kSimColTiles; ++ct) {
        const int j0 = (colGroup * kSimColTiles + ct) * kSimCols;
        if (j0 >= pr.refLen) break;
        __syncthreads();                                               // the previous tile's columns are no longer read
        for (int t = tid; t < kSimCols * 24; t += 32 * kSimWarps) {
            const int jj = t / 24, w = t - jj * 24;
            sR[jj][w] = (j0 + jj < pr.refLen) ? __ldg(refCols + static_cast<size_t>(j0 + jj) * 24 + w) : 0.0f;
        }
        __syncthreads();
        if (tid < kSimCols) {
            unsigned live = 0;
#pragma unroll
            for (int l = 0; l < 21; ++l) live |= (sR[tid][l] != 0.0f ? 1u : 0u) << l;
            sLive[tid] = live;
            // compact (gapOpen, gapExtend) of the reference columns, once per column tile
            if (rowTile == 0 && j0 + tid < pr.refLen)
                reinterpret_cast<float2 *>(sim + si.gapOff)[j0 + tid] = make_float2(sR[tid][22], sR[tid][23]);
        }
        __syncthreads();
        const int nCols = min(kSimCols, pr.refLen - j0);

        for (int jj = 0; jj < nCols; ++jj) {
            const float *r = sR[jj];
            float num = 0.0f;
            unsigned live = sLive[jj];
            while (live) {                                             // warp-uniform: the column is the same for every lane
                const int l = __ffs(live) - 1;
                live &= live - 1;
                const float rl = r[l];
                float Sl[24];
                {
                    const float4 *row = reinterpret_cast<const float4 *>(sS + l * 24);
#pragma unroll
                    for (int t = 0; t < 6; ++t) { const float4 x = row[t]; Sl[4 * t] = x.x; Sl[4 * t + 1] = x.y; Sl[4 * t + 2] = x.z; Sl[4 * t + 3] = x.w; }
                }
                float v[8];
#pragma unroll
                for (int m = 0; m < 8; ++m)
                    v[m] = __fmaf_rn(rl, __fmul_rn(q[8 + m], Sl[8 + m]), __fmul_rn(__fmul_rn(q[m], Sl[m]), rl));
#pragma unroll
                for (int m = 16; m < 21; ++m) num = __fmaf_rn(__fmul_rn(rl, q[m]), Sl[m], num);
                float h = __fadd_rn(v[0], v[1]);
#pragma unroll
                for (int m = 2; m < 8; ++m) h = __fadd_rn(h, v[m]);
                num = __fadd_rn(h, num);
            }
            // Gap-character terms (numeratorAa's tail). A term with a zero factor adds an exact zero, so the query-gap loop runs only
            // in lanes whose row holds gaps and only over the column's non-zero letters, the reference-gap loop only when the
            // column holds gaps (warp-uniform); the surviving terms keep the reference's order and form.
            if (q[21] != 0.0f) {
                unsigned lv = sLive[jj] & 0xFFFFFu;
                while (lv) {
                    const int l = __ffs(lv) - 1;
                    lv &= lv - 1;
                    num = __fadd_rn(num, __fmul_rn(__fmul_rn(r[l], q[21]), g));
                }
                if (sLive[jj] & (1u << 20)) num = __fmaf_rn(__fmul_rn(r[20], q[21]), g, num);
            }
            if (r[21] != 0.0f) {
#pragma unroll
                for (int m = 0; m < 20; ++m) num = __fadd_rn(num, __fmul_rn(__fmul_rn(q[m], r[21]), g));
                num = __fmaf_rn(g, __fmul_rn(r[21], q[20]), num);
            }
            sOut[warp][jj][lane] = ieeeDiv ? __fdiv_rn(num, denom) : exactDiv(num, denom, rcp);
        }
        __syncwarp();
        // anti-diagonal d of the tile holds the cells (row i0 + ii, column j0 + d - ii): one run of adjacent words of [i + j][i]
        float *dst = sim + si.simOff + static_cast<long long>(i0 + j0) * si.stride + i0 + lane;
        for (int d = 0; d < 32 + nCols - 1; ++d) {
            const int jj = d - lane;
            if (jj >= 0 && jj < nCols && i < pr.qryLen) dst[static_cast<long long>(d) * si.stride] = sOut[warp][jj][lane];
        }
        __syncwarp();
    }
}

// grid.x must cover the pair with the most tiles: tiles(refLen, qryLen) for upper bounds of the lengths
int simMatrixTiles(int refLen, int qryLen) {
    const int colTiles = (refLen + kSimCols - 1) / kSimCols;
    return ((colTiles + kSimColTiles - 1) / kSimColTiles) * ((qryLen + 32 * kSimWarps - 1) / (32 * kSimWarps));
}

cudaError_t launchSimMatrixAa(const float *prof, const DevPair *pairs, const int *order, int nOrder, const DevSim *simInfo, float *sim, const float *score,
                              int maxTiles, cudaStream_t stream) {
    if (nOrder <= 0 || maxTiles <= 0) return cudaSuccess;
    for (int begin = 0; begin < nOrder; begin += 65535) {           // grid.y limit
        const int cnt = (nOrder - begin < 65535) ? nOrder - begin : 65535;
        simMatrixAaKernel<<<dim3(maxTiles, cnt), 32 * kSimWarps, 0, stream>>>(prof, pairs, order + begin, simInfo, sim, score);
    }
    return cudaGetLastError();
}

} // namespace twl

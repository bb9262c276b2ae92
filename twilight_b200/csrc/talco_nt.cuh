// talco_nt.cuh — pieces shared by the register-resident nucleotide TALCO-XDrop kernels (talco_wavefront.cu: one role per thread;
// talco_duo.cu: recurrence warps + score warps): the similarity score of the cells a thread holds on one anti-diagonal, in the
// exact operation order of the reference (src/TALCO-XDrop.cpp:372-444), and small helpers.
#pragma once
#include "talco_score.cuh"
#include "twl_device.cuh"

namespace twl {

__device__ __forceinline__ unsigned long long globalTimerNs() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
// order-preserving int image of a float (for integer max reductions)
__device__ __forceinline__ int orderedInt(float f) {
    const int b = __float_as_int(f);
    return b ^ ((b >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float orderedFloat(int o) { return __int_as_float(o ^ ((o >> 31) & 0x7fffffff)); }

__device__ __forceinline__ void prefetchL1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// Column-pair numerators of the four slots of a thread. MC = 0: any 5x5 matrix, full reference order.
// MC = 1 ("DNA3Z"): the built-in nucleotide matrix shape (scoring-matrix.cpp:103-112 without --wildcard):
// S[l][l] = A, S[l][m] = B for |l-m| = 2, C otherwise, and an all-zero N row/column. The N terms of the reference sum
// are exact zeros and are dropped; every remaining product and sum is evaluated in the reference order.
template <int MC, int KS>
__device__ __forceinline__ void numerators4(const float (&r)[KS][6], const float (&q)[KS][6], const TalcoArgs &a,
                                            float (&num)[KS]) {
    if (MC == 0) {
#pragma unroll
        for (int c = 0; c < KS; ++c) {
            float n = 0.0f;
#pragma unroll
            for (int l = 0; l < 5; ++l) {
                const float t0 = __fmul_rn(__fmul_rn(q[c][0], a.scoreNt[l * 5 + 0]), r[c][l]);
                const float t1 = __fmul_rn(__fmul_rn(q[c][1], a.scoreNt[l * 5 + 1]), r[c][l]);
                const float t2 = __fmul_rn(__fmul_rn(q[c][2], a.scoreNt[l * 5 + 2]), r[c][l]);
                const float t3 = __fmul_rn(__fmul_rn(q[c][3], a.scoreNt[l * 5 + 3]), r[c][l]);
                const float t4 = __fmul_rn(__fmul_rn(q[c][4], a.scoreNt[l * 5 + 4]), r[c][l]);
                n = __fadd_rn(n, __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3), t4));
            }
            num[c] = n;
        }
    } else {
        const float A = a.scoreNt[0], B = a.scoreNt[2], C = a.scoreNt[1];
#pragma unroll
        for (int c = 0; c < KS; ++c) {
            float qa[4], qb[4], qc[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) { qa[m] = __fmul_rn(q[c][m], A); qb[m] = __fmul_rn(q[c][m], B); qc[m] = __fmul_rn(q[c][m], C); }
            const float r0 = r[c][0], r1 = r[c][1], r2 = r[c][2], r3 = r[c][3];
            const float h0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qa[0], r0), __fmul_rn(qc[1], r0)), __fmul_rn(qb[2], r0)), __fmul_rn(qc[3], r0));
            const float h1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qc[0], r1), __fmul_rn(qa[1], r1)), __fmul_rn(qc[2], r1)), __fmul_rn(qb[3], r1));
            const float h2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qb[0], r2), __fmul_rn(qc[1], r2)), __fmul_rn(qa[2], r2)), __fmul_rn(qc[3], r2));
            const float h3 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(qc[0], r3), __fmul_rn(qb[1], r3)), __fmul_rn(qc[2], r3)), __fmul_rn(qa[3], r3));
            num[c] = __fadd_rn(__fadd_rn(__fadd_rn(h0, h1), h2), h3);
        }
    }
}

// Similarity scores (TALCO-XDrop.cpp:372-444: numerator / denominator) of the KS cells a thread holds on the anti-diagonal whose
// slot-0 cell pairs row iBase with global reference column colBase - iBase (colBase = refOff + k), plus the position-specific
// gap penalties of those reference columns. Contains warp votes: call it with the whole warp.
template <int MC, int KS>
__device__ __forceinline__ void scoreDiagonal(const TalcoArgs &a, const float4 *refX, int refN4, int colBase, int iBase, const float (&q)[KS][6],
                                              float gapChar, int kind, int divMode, float denom, float rcp,
                                              float (&num)[KS], float (&gOpR)[KS], float (&gExR)[KS]) {
    float r[KS][6];
    bool gapQ = false, gapR = false;
    {
        // Slot c reads global reference column m-c with m = colBase - iBase. iBase is a multiple of 4, so m&3 is the same for
        // every thread: the four slots hit the four streams of the de-interleaved layout at float4 index m>>2 (or one less once
        // m-c crosses a multiple of 4). Columns outside [0, refLen) are only touched by slots that are not live; the profile
        // buffer is padded so the reads stay inside the allocation and their values are discarded.
        // 32-bit index arithmetic (a side has far fewer than 2^31 float4): the stream part of the index is warp-uniform, the thread
        // adds its own column group; one 64-bit multiply-add per load instead of a carry chain
        const int m = colBase - iBase;
        const int u = colBase & 3;
        const int group = m >> 2;
        const int yOff = 4 * refN4;
#pragma unroll
        for (int c = 0; c < KS; ++c) {
            int at;
            if (KS == 4) {
                const int stream = (u - c) & 3;
                at = stream * refN4 - ((c > u) ? 1 : 0) + group;
            } else {
                at = static_cast<int>(ntColIndex(m - c, refN4));   // rows per thread < 4: the stream differs between threads
            }
            const float4 x = __ldg(refX + at);
            const float4 y = __ldg(refX + (at + yOff));
            r[c][0] = x.x; r[c][1] = x.y; r[c][2] = x.z; r[c][3] = x.w; r[c][4] = y.x; r[c][5] = y.y;
            gOpR[c] = y.z; gExR[c] = y.w;
            gapR = gapR || (y.y != 0.0f);
            gapQ = gapQ || (q[c][5] != 0.0f);
        }
    }
    if (kind & kRefOneHot) {
        // The reference side is a single gap-free sequence: its columns are exactly one-hot (count 1.0), so every term of the
        // reference's sum except those of the one present letter is an exact zero, and the similarity depends only on (query row,
        // reference letter). q[c][a] holds that value for letter a, already divided (computed when the row was fetched); the FMA
        // chain just picks it.
#pragma unroll
        for (int c = 0; c < KS; ++c)
            num[c] = __fmaf_rn(r[c][4], q[c][4], __fmaf_rn(r[c][3], q[c][3], __fmaf_rn(r[c][2], q[c][2],
                     __fmaf_rn(r[c][1], q[c][1], __fmul_rn(r[c][0], q[c][0])))));
        return;
    }
    if (kind & kQryOneHot) {
        // The query side is one-hot: q[c][l] holds S[l][b] for the row's letter b; the surviving terms are S[l][b]*r[l], summed
        // in the reference's order, then the one non-zero gap term.
#pragma unroll
        for (int c = 0; c < KS; ++c) {
            float n = __fmul_rn(q[c][0], r[c][0]);
            n = __fadd_rn(n, __fmul_rn(q[c][1], r[c][1]));
            n = __fadd_rn(n, __fmul_rn(q[c][2], r[c][2]));
            n = __fadd_rn(n, __fmul_rn(q[c][3], r[c][3]));
            n = __fadd_rn(n, __fmul_rn(q[c][4], r[c][4]));
            num[c] = __fmaf_rn(r[c][5], gapChar, n);
        }
    } else {
        numerators4<MC, KS>(r, q, a, num);
        // gap-character terms (TALCO-XDrop.cpp:393-394): each loop adds exact zeros unless the query (resp. reference) column
        // holds gaps, so it is skipped when no lane of the warp needs it
        if (__any_sync(0xffffffffu, gapQ)) {
#pragma unroll
            for (int c = 0; c < KS; ++c)
#pragma unroll
                for (int l = 0; l < 5; ++l) num[c] = __fmaf_rn(__fmul_rn(r[c][l], q[c][5]), gapChar, num[c]);
        }
        if (__any_sync(0xffffffffu, gapR)) {
#pragma unroll
            for (int c = 0; c < KS; ++c)
#pragma unroll
                for (int m = 0; m < 5; ++m) num[c] = __fmaf_rn(__fmul_rn(r[c][5], q[c][m]), gapChar, num[c]);
        }
    }
    if (divMode == 1) {
        // reciprocal-based exact division; numerators that are non-zero but tiny (|n| < 2^-60, where the quotient or the FMA
        // residual could leave the normal range) take the IEEE divide instead
        unsigned tiny = 0xffffffffu;
#pragma unroll
        for (int c = 0; c < KS; ++c) tiny = min(tiny, (__float_as_uint(num[c]) & 0x7fffffffu) - 1u);
        if (tiny < 0x21800000u - 1u) {
#pragma unroll
            for (int c = 0; c < KS; ++c) num[c] = __fdiv_rn(num[c], denom);
        } else {
#pragma unroll
            for (int c = 0; c < KS; ++c) num[c] = exactDivNormal(num[c], denom, rcp);
        }
    } else if (divMode == 2) {
#pragma unroll
        for (int c = 0; c < KS; ++c) num[c] = __fdiv_rn(num[c], denom);
    }
}

} // namespace twl

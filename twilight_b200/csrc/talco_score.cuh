// talco_score.cuh — the column-pair similarity numerator of the TALCO-XDrop recurrence, in the exact operation
// order of the reference x86 build (src/TALCO-XDrop.cpp:372-444 with TALCO_SIMD, GCC contraction on; see
// oracle/twl_oracle.cpp for the CPU statement of the same order). Every product and sum is an explicit
// round-to-nearest intrinsic so nvcc cannot fuse or reassociate anything; fused operations of the reference build are
// explicit __fmaf_rn.
#pragma once
#include <cuda_runtime.h>

namespace twl {

// num/den rounded to nearest, bit-identical to IEEE division (Markstein: with rcp = RN(1/den), q = RN(num*rcp),
// r = num - den*q exactly by FMA, RN(q + r*rcp) is the correctly rounded quotient). Tiny numerators (where q could be
// subnormal) take the IEEE divide.
__device__ __forceinline__ float exactDiv(float num, float den, float rcp) {
    const float q = __fmul_rn(num, rcp);
    const float r = __fmaf_rn(-q, den, num);
    float res = __fmaf_rn(r, rcp, q);
    const float an = fabsf(num);
    if (an < 1e-18f && an > 0.0f) res = __fdiv_rn(num, den);
    return res;
}
// The same without the tiny-numerator guard: the caller guarantees num == 0 or |num| >= 2^-60.
__device__ __forceinline__ float exactDivNormal(float num, float den, float rcp) {
    const float q = __fmul_rn(num, rcp);
    const float r = __fmaf_rn(-q, den, num);
    return __fmaf_rn(r, rcp, q);
}

// Nucleotide, P = 6. r[0..5], q[0..5] are profile columns, S the 5x5 matrix (row-major, any address space), g the
// gap-character score.
template <typename MatPtr>
__device__ __forceinline__ float numeratorNt(const float (&r)[6], const float (&q)[6], MatPtr S, float g) {
    float num = 0.0f;
#pragma unroll
    for (int l = 0; l < 5; ++l) {
        const float t0 = __fmul_rn(__fmul_rn(q[0], S[l * 5 + 0]), r[l]);
        const float t1 = __fmul_rn(__fmul_rn(q[1], S[l * 5 + 1]), r[l]);
        const float t2 = __fmul_rn(__fmul_rn(q[2], S[l * 5 + 2]), r[l]);
        const float t3 = __fmul_rn(__fmul_rn(q[3], S[l * 5 + 3]), r[l]);
        const float t4 = __fmul_rn(__fmul_rn(q[4], S[l * 5 + 4]), r[l]);
        const float h = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3), t4);
        num = __fadd_rn(num, h);
    }
    // Gap-character terms: all ten are exact zeros (and leave num unchanged) unless a gap count is non-zero.
    if (q[5] != 0.0f || r[5] != 0.0f) {
#pragma unroll
        for (int l = 0; l < 5; ++l) num = __fmaf_rn(__fmul_rn(r[l], q[5]), g, num);
#pragma unroll
        for (int m = 0; m < 5; ++m) num = __fmaf_rn(__fmul_rn(r[5], q[m]), g, num);
    }
    return num;
}

// Same, with the query-side products qs[l*5+m] = q[m]*S[l][m] precomputed (they depend on the query column only).
__device__ __forceinline__ float numeratorNtPre(const float (&r)[6], const float (&qs)[25], const float (&q)[6], float g) {
    float num = 0.0f;
#pragma unroll
    for (int l = 0; l < 5; ++l) {
        const float t0 = __fmul_rn(qs[l * 5 + 0], r[l]);
        const float t1 = __fmul_rn(qs[l * 5 + 1], r[l]);
        const float t2 = __fmul_rn(qs[l * 5 + 2], r[l]);
        const float t3 = __fmul_rn(qs[l * 5 + 3], r[l]);
        const float t4 = __fmul_rn(qs[l * 5 + 4], r[l]);
        const float h = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3), t4);
        num = __fadd_rn(num, h);
    }
    if (q[5] != 0.0f || r[5] != 0.0f) {
#pragma unroll
        for (int l = 0; l < 5; ++l) num = __fmaf_rn(__fmul_rn(r[l], q[5]), g, num);
#pragma unroll
        for (int m = 0; m < 5; ++m) num = __fmaf_rn(__fmul_rn(r[5], q[m]), g, num);
    }
    return num;
}

// Protein, P = 22; S is the 21x21 matrix.
// Reference letters with a zero count contribute only exact zeros (every product has the factor r[l] = 0 and is added to
// a finite sum), so they are skipped: each lane walks the non-zero letters of ITS column in increasing order, which keeps
// the order of the floating-point sum; a leaf column (one-hot) costs one of 21 iterations, a typical profile column 3-6.
// The sign of a zero sum may differ from the reference's (-0 vs +0); it cannot reach any comparison or stored value.
template <typename MatPtr>
__device__ __forceinline__ float numeratorAa(const float (&r)[22], const float (&q)[22], MatPtr S, float g) {
    float num = 0.0f;
    unsigned live = 0;
#pragma unroll
    for (int l = 0; l < 21; ++l) live |= (r[l] != 0.0f ? 1u : 0u) << l;
    while (live) {
        const int l = __ffs(live) - 1;
        live &= live - 1;
        const float rl = r[l];
        float v[8];
#pragma unroll
        for (int m = 0; m < 8; ++m)
            v[m] = __fmaf_rn(rl, __fmul_rn(q[8 + m], S[l * 21 + 8 + m]), __fmul_rn(__fmul_rn(q[m], S[l * 21 + m]), rl));
#pragma unroll
        for (int m = 16; m < 21; ++m) num = __fmaf_rn(__fmul_rn(rl, q[m]), S[l * 21 + m], num);
        float h = __fadd_rn(v[0], v[1]);
#pragma unroll
        for (int m = 2; m < 8; ++m) h = __fadd_rn(h, v[m]);
        num = __fadd_rn(h, num);
    }
    if (q[21] != 0.0f || r[21] != 0.0f) {
#pragma unroll
        for (int l = 0; l < 20; ++l) num = __fadd_rn(num, __fmul_rn(__fmul_rn(r[l], q[21]), g));
        num = __fmaf_rn(__fmul_rn(r[20], q[21]), g, num);
#pragma unroll
        for (int m = 0; m < 20; ++m) num = __fadd_rn(num, __fmul_rn(__fmul_rn(q[m], r[21]), g));
        num = __fmaf_rn(g, __fmul_rn(r[21], q[20]), num);
    }
    return num;
}

} // namespace twl

// tools/tc_probe/tc_score_probe.cu — EXPERIMENT (not part of the product library): should the protein substitution-score
// contraction P1 * S * P2^T of the TALCO-XDrop cell (src/TALCO-XDrop.cpp:405-440; 21 x 21 matrix, profile width 22) run on
// tensor cores? north_star: "only where ncu shows it beats CUDA-core math". Four ways to score the same rows x columns tile:
//   A exact     the product path: numeratorAa (talco_score.cuh), the reference's operation order, zero letters skipped
//   D factored  T'[i][l] = sum_m q_i[m] S[l][m] (+ gap-character terms folded in) once per row, then a 22-term FP32 dot per cell
//               on the CUDA cores: same mathematics, different summation order -> NOT bit-identical
//   B tf32      the same factored product as a GEMM tile on the tensor cores (mma.sync m16n8k8 TF32, K = 24): operands
//               rounded to 10 mantissa bits
//   C 3xtf32    error-compensated split (hi/lo TF32 parts, 3 MMAs per k-step): FP32-like accuracy on tensor cores
// (mma.sync is the legacy tensor path; it bounds from above what a tcgen05 version of a K = 24 micro-GEMM could gain, because the
// per-cell work that remains — recurrences, pruning, traceback — is untouched by either.)
// Prints one JSON line: G cell-scores/s of each variant and their deviation from A.
#include "../../twilight_b200/csrc/talco_score.cuh"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

constexpr int P = 22, M = 21, KP = 24;

// A: one thread per cell (i = row of the query profile, j = column of the reference profile), lanes along i like the DP kernels
__global__ void exactKernel(const float *Q, const float *R, const float *S, float g, int nQ, int nR, float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j0 = blockIdx.y * 64;
    if (i >= nQ) return;
    float q[P];
#pragma unroll
    for (int t = 0; t < P; ++t) q[t] = Q[i * P + t];
    for (int j = j0; j < min(nR, j0 + 64); ++j) {
        float r[P];
#pragma unroll
        for (int t = 0; t < P; ++t) r[t] = __ldg(R + j * P + t);
        out[static_cast<size_t>(j) * nQ + i] = twl::numeratorAa(r, q, S, g);
    }
}

// T'[i][l], l < 21: sum_m q[m] S[l][m] + g q[21];  T'[i][21] = g * sum_m q[m];  so that score = sum_{l<22} T'[i][l] r[l]
__global__ void factorKernel(const float *Q, const float *S, float g, int nQ, float *T) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nQ) return;
    float sum = 0.f;
    for (int m = 0; m < M; ++m) sum += Q[i * P + m];
    for (int l = 0; l < M; ++l) {
        float t = 0.f;
        for (int m = 0; m < M; ++m) t = fmaf(Q[i * P + m], S[l * M + m], t);
        T[i * KP + l] = fmaf(g, Q[i * P + M], t);
    }
    T[i * KP + 21] = g * sum; T[i * KP + 22] = 0.f; T[i * KP + 23] = 0.f;
}

__global__ void denseKernel(const float *T, const float *R, int nQ, int nR, float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j0 = blockIdx.y * 64;
    if (i >= nQ) return;
    float t[P];
#pragma unroll
    for (int l = 0; l < P; ++l) t[l] = T[i * KP + l];
    for (int j = j0; j < min(nR, j0 + 64); ++j) {
        float acc = 0.f;
#pragma unroll
        for (int l = 0; l < P; ++l) acc = fmaf(t[l], __ldg(R + j * P + l), acc);
        out[static_cast<size_t>(j) * nQ + i] = acc;
    }
}

__device__ __forceinline__ unsigned toTf32(float x) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return r; }
__device__ __forceinline__ void mmaTf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// One warp: 16 rows x 64 columns (8 tiles of 16x8), K = 24 in three k-steps. R24 is the reference profile padded to 24 floats.
template <bool SPLIT>
__global__ void tensorKernel(const float *T, const float *R24, int nQ, int nR, float *out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int i0 = warp * 16, j0 = blockIdx.y * 64;
    if (i0 >= nQ) return;
    const int gr = lane >> 2, gc = lane & 3;
    unsigned aHi[3][4], aLo[3][4];
#pragma unroll
    for (int ks = 0; ks < 3; ++ks) {
        const float v[4] = {T[(i0 + gr) * KP + ks * 8 + gc], T[(i0 + gr + 8) * KP + ks * 8 + gc], T[(i0 + gr) * KP + ks * 8 + gc + 4], T[(i0 + gr + 8) * KP + ks * 8 + gc + 4]};
#pragma unroll
        for (int x = 0; x < 4; ++x) { aHi[ks][x] = toTf32(v[x]); aLo[ks][x] = SPLIT ? toTf32(v[x] - __uint_as_float(aHi[ks][x])) : 0u; }
    }
    for (int jt = 0; jt < 8; ++jt) {
        const int j = j0 + jt * 8;
        if (j >= nR) break;
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 3; ++ks) {
            const float b0 = __ldg(R24 + (j + gr) * KP + ks * 8 + gc), b1 = __ldg(R24 + (j + gr) * KP + ks * 8 + gc + 4);
            unsigned bHi[2] = {toTf32(b0), toTf32(b1)};
            if (SPLIT) {
                unsigned bLo[2] = {toTf32(b0 - __uint_as_float(bHi[0])), toTf32(b1 - __uint_as_float(bHi[1]))};
                mmaTf32(c, aLo[ks], bHi);
                mmaTf32(c, aHi[ks], bLo);
            }
            mmaTf32(c, aHi[ks], bHi);
        }
        const int col = j + 2 * gc;
        out[static_cast<size_t>(col) * nQ + i0 + gr] = c[0];
        out[static_cast<size_t>(col + 1) * nQ + i0 + gr] = c[1];
        out[static_cast<size_t>(col) * nQ + i0 + gr + 8] = c[2];
        out[static_cast<size_t>(col + 1) * nQ + i0 + gr + 8] = c[3];
    }
}

int main(int argc, char **argv) {
    const int nQ = 512, nR = (argc > 1) ? std::atoi(argv[1]) : 8192, reps = 20;
    std::mt19937 rng(7);
    std::vector<float> S(M * M), Q(static_cast<size_t>(nQ) * P, 0.f), R(static_cast<size_t>(nR) * P, 0.f), R24(static_cast<size_t>(nR) * KP, 0.f);
    for (int l = 0; l < M; ++l)
        for (int m = l; m < M; ++m) { const float v = 5.0f * (static_cast<int>(rng() % 16) - (l == m ? 0 : 8)); S[l * M + m] = S[m * M + l] = (l == 20 || m == 20) ? 0.f : v; }
    auto fill = [&](std::vector<float> &prof, int n) {      // profile columns of a family of 8: 1-5 distinct letters, sometimes gaps, weights ~1
        for (int c = 0; c < n; ++c) {
            int members = 8;
            const int gaps = (rng() % 4 == 0) ? static_cast<int>(rng() % 4) : 0;
            const int kinds = 1 + static_cast<int>(rng() % 5);
            int letters[5];
            for (int k = 0; k < kinds; ++k) letters[k] = static_cast<int>(rng() % 20);
            for (int s = 0; s < members; ++s) {
                const float w = 0.5f + static_cast<float>(rng() % 1000) / 1000.0f;
                if (s < gaps) prof[static_cast<size_t>(c) * P + 21] += w; else prof[static_cast<size_t>(c) * P + letters[rng() % kinds]] += w;
            }
        }
    };
    fill(Q, nQ); fill(R, nR);
    for (int j = 0; j < nR; ++j) for (int l = 0; l < P; ++l) R24[static_cast<size_t>(j) * KP + l] = R[static_cast<size_t>(j) * P + l];
    const float g = -5.0f;
    float *dS, *dQ, *dR, *dR24, *dT, *dOut[4];
    CK(cudaMalloc(&dS, S.size() * 4)); CK(cudaMalloc(&dQ, Q.size() * 4)); CK(cudaMalloc(&dR, R.size() * 4)); CK(cudaMalloc(&dR24, R24.size() * 4));
    CK(cudaMalloc(&dT, static_cast<size_t>(nQ) * KP * 4));
    for (auto &o : dOut) CK(cudaMalloc(&o, static_cast<size_t>(nQ) * nR * 4));
    CK(cudaMemcpy(dS, S.data(), S.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dQ, Q.data(), Q.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dR, R.data(), R.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dR24, R24.data(), R24.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const dim3 gridCell((nQ + 127) / 128, (nR + 63) / 64), gridTc((nQ / 16 * 32 + 127) / 128, (nR + 63) / 64);
    float ms[4] = {0, 0, 0, 0};
    for (int v = 0; v < 4; ++v) {
        for (int r = 0; r < reps + 2; ++r) {
            if (r == 2) CK(cudaEventRecord(e0));
            if (v == 0) exactKernel<<<gridCell, 128>>>(dQ, dR, dS, g, nQ, nR, dOut[0]);
            else {
                factorKernel<<<(nQ + 127) / 128, 128>>>(dQ, dS, g, nQ, dT);      // counted: it belongs to the factored variants
                if (v == 1) denseKernel<<<gridCell, 128>>>(dT, dR, nQ, nR, dOut[1]);
                else if (v == 2) tensorKernel<false><<<gridTc, 128>>>(dT, dR24, nQ, nR, dOut[2]);
                else tensorKernel<true><<<gridTc, 128>>>(dT, dR24, nQ, nR, dOut[3]);
            }
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        CK(cudaEventElapsedTime(&ms[v], e0, e1));
        ms[v] /= reps;
    }
    const size_t n = static_cast<size_t>(nQ) * nR;
    std::vector<float> h[4];
    for (int v = 0; v < 4; ++v) { h[v].resize(n); CK(cudaMemcpy(h[v].data(), dOut[v], n * 4, cudaMemcpyDeviceToHost)); }
    double maxAbs[4] = {0, 0, 0, 0}, maxRel[4] = {0, 0, 0, 0};
    size_t differ[4] = {0, 0, 0, 0};
    double scale = 0;
    for (size_t k = 0; k < n; ++k) scale = std::max(scale, static_cast<double>(std::fabs(h[0][k])));
    for (int v = 1; v < 4; ++v)
        for (size_t k = 0; k < n; ++k) {
            const double d = std::fabs(static_cast<double>(h[v][k]) - h[0][k]);
            maxAbs[v] = std::max(maxAbs[v], d);
            if (h[0][k] != 0.f) maxRel[v] = std::max(maxRel[v], d / std::fabs(h[0][k]));
            differ[v] += (h[v][k] != h[0][k]);
        }
    const char *name[4] = {"exact_cuda_core", "factored_fp32_cuda_core", "tf32_tensor_core", "3xtf32_tensor_core"};
    std::printf("{\"rows\": %d, \"cols\": %d, \"cells\": %zu, \"max_abs_score\": %.1f, \"variants\": {", nQ, nR, n, scale);
    for (int v = 0; v < 4; ++v)
        std::printf("%s\"%s\": {\"ms\": %.4f, \"gscores_per_s\": %.2f, \"max_abs_dev\": %.6g, \"max_rel_dev\": %.3g, \"cells_not_bit_identical\": %zu}", v ? ", " : "",
                    name[v], ms[v], n / (ms[v] * 1e-3) / 1e9, maxAbs[v], maxRel[v], differ[v]);
    std::printf("}}\n");
    return 0;
}

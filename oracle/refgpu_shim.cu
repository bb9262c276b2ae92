// oracle/refgpu_shim.cu — TEST / MEASUREMENT INFRASTRUCTURE ONLY (never on the product path).
//
// Times the reference's OWN CUDA kernel, device_function::parallelProfileAlignment_Fast (src/cuda/device-function.cu:753,
// compiled unmodified where it lies, for sm_100 as the reference's CMake would on a box with a B200), on host-fed profiles
// of one guide-tree level, launched the way the reference host does it (src/cuda/alignment-gpu.cu:206-333: rounds of at
// most _BLOCKSIZE = 2048 pairs, one pair per 256-thread block, memBlock == 1, alnLen preset to 0 = "global, include head").
// It answers SURVEY.md §2's bar "beat the existing GPU kernel compiled for sm_100" with a number; the kernel is NOT
// result-equivalent to the reference CPU path (int16 scores, marker 200, wavefront cap 1350, x-drop 600*|gapExtend|, no
// gappy-column removal), so only time per pair and failure counts are comparable, not paths.
#include "device-function.cuh"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

#define RG_CUDA(call)                                                                                  \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) { std::fprintf(stderr, "refgpu: %s: %s\n", #call, cudaGetErrorString(e_)); return -1; } \
    } while (0)

extern "C" {

// freq [nPairs][2][seqLen][P], gapOpen / gapExtend [nPairs][2][seqLen], len / num [2*nPairs], param [(P-1)^2 + 4]
// (matrix, gapOpen, gapExtend, gapBoundary, xdrop) — the layouts of GPU_pointers (src/msa.hpp:220-268).
// Outputs: aln [nPairs][2*seqLen] (host), alnLen [nPairs] (-1 = the kernel gave up on the pair), *kernelMs = sum of the
// CUDA-event times of the kernel launches only (inputs already resident, like the B200 arm's `value`).
int refgpu_level(int device, int P, int nPairs, int seqLen, const float *freq, const float *gapOpen, const float *gapExtend,
                 const int32_t *len, const int32_t *num, const float *param, int8_t *aln, int32_t *alnLen, float *kernelMs, int repeats) {
    RG_CUDA(cudaSetDevice(device));
    const int paramSize = (P - 1) * (P - 1) + 4;
    const int block = device_function::_BLOCKSIZE;
    float *dFreq = nullptr, *dOp = nullptr, *dEx = nullptr, *dParam = nullptr;
    int8_t *dAln = nullptr;
    int32_t *dLen = nullptr, *dNum = nullptr, *dAlnLen = nullptr, *dInfo = nullptr;
    const size_t perPairF = static_cast<size_t>(P) * 2 * seqLen, perPairG = static_cast<size_t>(2) * seqLen;
    RG_CUDA(cudaMalloc(&dFreq, perPairF * nPairs * sizeof(float)));
    RG_CUDA(cudaMalloc(&dOp, perPairG * nPairs * sizeof(float)));
    RG_CUDA(cudaMalloc(&dEx, perPairG * nPairs * sizeof(float)));
    RG_CUDA(cudaMalloc(&dAln, perPairG * nPairs));
    RG_CUDA(cudaMalloc(&dLen, sizeof(int32_t) * 2 * nPairs));
    RG_CUDA(cudaMalloc(&dNum, sizeof(int32_t) * 2 * nPairs));
    RG_CUDA(cudaMalloc(&dAlnLen, sizeof(int32_t) * nPairs));
    RG_CUDA(cudaMalloc(&dInfo, sizeof(int32_t) * 4 * ((nPairs + block - 1) / block)));
    RG_CUDA(cudaMalloc(&dParam, sizeof(float) * paramSize));
    RG_CUDA(cudaMemcpy(dFreq, freq, perPairF * nPairs * sizeof(float), cudaMemcpyHostToDevice));
    RG_CUDA(cudaMemcpy(dOp, gapOpen, perPairG * nPairs * sizeof(float), cudaMemcpyHostToDevice));
    RG_CUDA(cudaMemcpy(dEx, gapExtend, perPairG * nPairs * sizeof(float), cudaMemcpyHostToDevice));
    RG_CUDA(cudaMemcpy(dLen, len, sizeof(int32_t) * 2 * nPairs, cudaMemcpyHostToDevice));
    RG_CUDA(cudaMemcpy(dNum, num, sizeof(int32_t) * 2 * nPairs, cudaMemcpyHostToDevice));
    RG_CUDA(cudaMemcpy(dParam, param, sizeof(float) * paramSize, cudaMemcpyHostToDevice));
    const int rounds = (nPairs + block - 1) / block;
    std::vector<int32_t> info(4 * rounds);
    for (int r = 0; r < rounds; ++r) {
        info[4 * r] = std::min(block, nPairs - r * block);
        info[4 * r + 1] = seqLen; info[4 * r + 2] = P; info[4 * r + 3] = block;
    }
    RG_CUDA(cudaMemcpy(dInfo, info.data(), sizeof(int32_t) * info.size(), cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    RG_CUDA(cudaEventCreate(&e0));
    RG_CUDA(cudaEventCreate(&e1));
    float best = -1.f;
    for (int rep = 0; rep < std::max(1, repeats); ++rep) {
        RG_CUDA(cudaMemset(dAlnLen, 0, sizeof(int32_t) * nPairs));   // hostAlnLen[i] = i / pairPerMemBlock = 0 (alignment-gpu.cu:96)
        RG_CUDA(cudaMemset(dAln, 0, perPairG * nPairs));
        RG_CUDA(cudaEventRecord(e0));
        for (int r = 0; r < rounds; ++r) {
            const size_t at = static_cast<size_t>(r) * block;
            device_function::parallelProfileAlignment_Fast<<<block, device_function::_THREAD_NUM>>>(
                dFreq + perPairF * at, dAln + perPairG * at, dLen + 2 * at, dNum + 2 * at, dAlnLen + at, dInfo + 4 * r, dOp + perPairG * at,
                dEx + perPairG * at, dParam);
        }
        RG_CUDA(cudaEventRecord(e1));
        RG_CUDA(cudaEventSynchronize(e1));
        RG_CUDA(cudaGetLastError());
        float ms = 0.f;
        RG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (best < 0 || ms < best) best = ms;
    }
    if (kernelMs) *kernelMs = best;
    RG_CUDA(cudaMemcpy(aln, dAln, perPairG * nPairs, cudaMemcpyDeviceToHost));
    RG_CUDA(cudaMemcpy(alnLen, dAlnLen, sizeof(int32_t) * nPairs, cudaMemcpyDeviceToHost));
    cudaFree(dFreq); cudaFree(dOp); cudaFree(dEx); cudaFree(dAln); cudaFree(dLen); cudaFree(dNum); cudaFree(dAlnLen); cudaFree(dInfo); cudaFree(dParam);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

} // extern "C"

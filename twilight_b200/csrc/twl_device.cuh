// twl_device.cuh — device-side types shared by the TALCO-XDrop kernels and the C-ABI host code.
//
// HBM layout of a staged level batch (one allocation each, sized by the host per batch):
//   prof    float  packed column profiles, one column = PW = P+2 floats, 32-byte aligned:
//                  [0..P-1] weighted letter counts (calculateProfile, alignment-helper.cpp:8),
//                  [P] position-specific gap-open, [P+1] gap-extend (calculatePSGP, alignment-helper.cpp:168).
//                  PW = 8 for nucleotides (one 32 B sector per column), 24 for proteins.
//   pairs   DevPair[nPairs]   where each pair's ref/qry columns start inside `prof`, lengths, counts, per-pair
//                  Talco parameters (gapCharScore, xdrop, fLen: alignment-cpu.cpp:86-88, 116-129)
//   paths   int8   alignment paths, pair p at alnOff, capacity refLen+qryLen
//   results DevResult[nPairs]
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace twl {

constexpr int kMaxMarker = 1024;          // Talco_xdrop::Params::marker, TALCO-XDrop.cpp:51
constexpr int kInsBoundary = -2;          // I_BOUNDARY, TALCO-XDrop.cpp:33
constexpr int kDelBoundary = -3;          // D_BOUNDARY, TALCO-XDrop.cpp:34
constexpr int kStatusRetryWide = 100;     // internal: band exceeded this kernel's state capacity, rerun on the wide variant

struct DevPair {
    long long refOff;   // offset (in floats) of the first ref column inside prof
    long long qryOff;
    long long alnOff;   // offset (in bytes) of this pair's path inside paths
    int refLen, qryLen;
    float refNum, qryNum;
    float gapChar;
    int xdrop, fLen;
    int pad;
};

struct DevResult {
    int status;
    int pathLen;
    int tiles;
    int pad;
    unsigned long long cells;
    unsigned long long diagonals;
};

struct TalcoArgs {
    const float *prof;
    const DevPair *pairs;
    DevResult *results;
    int8_t *paths;
    const int *order;        // pair indices to process, heaviest first
    int *queue;              // global work cursor (device)
    const int *nWorkPtr;     // number of entries of `order` (device; may be produced by an earlier kernel)
    int *overflowList;       // pairs whose band exceeded stateCap are appended here for the wide variant (nullable)
    int *overflowCount;
    int marker;
    float gapOpen, gapExtend;
    const float *score;      // (P-1)^2 row-major (device)
    float scoreNt[25];       // the same matrix by value for the nucleotide kernels (lands in the constant bank)
    uint8_t *tbScratch;      // per-CTA traceback bytes
    size_t tbStride;
    float *stateScratch;     // per-CTA wavefront state when it does not fit in shared memory
    size_t stateStride;      // in 4-byte words
    int stateCap;            // cells per wavefront array (excluding padding)
};

} // namespace twl

// twl_device.cuh — device-side types shared by the TALCO-XDrop kernels and the C-ABI host code.
//
// HBM layout of a staged level batch (one allocation each, sized by the host per batch):
//   prof    float  packed column profiles, one column = PW = P+2 floats:
//                  [0..P-1] weighted letter counts (calculateProfile, alignment-helper.cpp:8),
//                  [P] position-specific gap-open, [P+1] gap-extend (calculatePSGP, alignment-helper.cpp:168).
//                  Proteins (PW = 24): array of columns, 96 B each.
//                  Nucleotides (PW = 8): a column is two float4 halves X = (A C G T) and Y = (N gap gapOpen gapExtend),
//                  and a side of n columns is stored as 8 streams of n4 = ceil(n/4) float4 each:
//                      X of column j at float4 index (j&3)*n4 + (j>>2),  Y at 4*n4 + (j&3)*n4 + (j>>2).
//                  The wavefront kernel gives each thread 4 consecutive rows, so at a fixed register slot the 32
//                  lanes of a warp read columns j, j-4, j-8, ...: with this de-interleaved layout that is ONE
//                  contiguous 512 B run per LDG.128 (4 L1 wavefronts) instead of 32 different 128 B lines.
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
constexpr int kRefOneHot = 1;             // every ref column is exactly one-hot: one of A,C,G,T,N has count 1.0, all else 0
constexpr int kQryOneHot = 2;
constexpr int kStatusEmptySide = 200;     // internal: one side has no columns
constexpr int kStatusRetryWide = 100;     // internal: band exceeded this kernel's state capacity, rerun on the wide variant

struct DevPair {
    long long refOff;   // offset (in floats) of the first ref column inside prof
    long long qryOff;
    long long alnOff;   // offset (in bytes) of this pair's path inside paths
    int refLen, qryLen;
    float refNum, qryNum;
    float gapChar;
    int xdrop, fLen;
    int pad;            // nucleotide fast-path flags: kRefOneHot / kQryOneHot (set by the packers when a side is a single gap-free sequence)
    int refN4, qryN4;   // nucleotide layout: stream length ceil(len/4) of each side
};

// float4 index of the X half of nucleotide column j inside a side with stream length n4 (Y half: + 4*n4)
__host__ __device__ inline long long ntColIndex(int j, int n4) { return static_cast<long long>(j & 3) * n4 + (j >> 2); }

struct DevResult {
    int status;
    int pathLen;        // with status kStatusRetryWide: ops already written for the completed tiles
    int tiles;          //                               number of completed tiles
    int pad;
    unsigned long long cells;       // completed tiles only
    unsigned long long diagonals;
    int resRefOff, resQryOff;       // with status kStatusRetryWide: where the failing tile starts; the next kernel of the chain resumes there
};

// Protein path: where a pair's similarity matrix and compact reference-side gap penalties live inside TalcoArgs::sim (offsets in
// floats). Matrix cell (query row i, reference column j) is at simOff + (i + j) * stride + i; the (gapOpen, gapExtend) pair of
// reference column j at gapOff + 2 * j. Both regions are padded so that the wavefront kernel's dead slots read inside the buffer.
struct DevSim {
    long long simOff;
    long long gapOff;
    int stride;
    int pad;
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
    int resume;              // this stage works on an overflow list: pairs continue at the tile recorded in `results`
    // Co-running wide worker (talco_wavefront.cu): the narrow kernel and a few wide CTAs run at the same time. The narrow
    // kernel hands pairs whose band outgrew its window to `feed*`; a wide CTA takes fed pairs first and otherwise works on
    // the main queue like everybody else, so it never idles while there is work and overflowed pairs do not wait for the
    // end of the stage.
    int coMode;              // 0 off, 1 narrow producer (counts finished pairs in mainDone), 2 wide worker
    int coTakeBelow;         // mode 2: also take main-queue entries with index below this (0 = never). Large batches only, and not
                             // the last wave, so that pairs handed over near the end find an idle wide worker
    int *mainDone;           // pairs of the main queue that are completely finished (either kernel)
    int *arrived;            // mode 2: wide CTAs count themselves in when they start; the host launches the narrow kernel behind a gate
                             // that waits for this count, so the wide workers hold their SMs before the narrow CTAs fill the machine
    int *watchdog;           // set by a wide worker that gave up waiting (kernels were not co-scheduled): the host then stops co-running
    int *heartbeat;          // bumped by the narrow kernel at every tile: lets a waiting wide worker tell "still running" from
                             // "not running at all" (kernels serialised by a profiler or CUDA_LAUNCH_BLOCKING)
    int *feedList;           // mode 2: entries appended by the producers (-1 until written)
    int *feedCount;          // mode 2: number of appended entries
    int *feedCursor;         // mode 2: next entry to take
    const float *sim;        // protein path (talco_sim.cu): similarity matrices + gap arrays of the batch, see DevSim
    const DevSim *simInfo;   // per pair
    unsigned long long *coTrace;   // diagnostics (nullable): per pair [4] = handed over, taken, finished (globaltimer ns), taker's role
};

} // namespace twl

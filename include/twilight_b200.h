/* twilight_b200.h — C ABI of the B200-native TWILIGHT alignment path (libtwilight_b200.so).
 *
 * This is the drop-in boundary for the reference's per-level alignment kernel:
 *   msa::alnFunction                                  (src/msa.hpp:175, invoked at src/progressive.cpp:180)
 *   cpu::alignmentKernel_CPU / parallelAlignmentCPU   (src/alignment-cpu.cpp:32, 36-183)
 *   gpu::alignmentKernel_GPU / parallelAlignmentGPU   (src/cuda/alignment-gpu.cu:12, 182-450)
 * The C++ adapter with the alnFunction signature that a maintainer links into the unchanged host lives in
 * twilight_b200/host/alignment_b200.cpp; INTEGRATION.md shows the two-line change in twilight-main.cpp.
 *
 * Conventions: every function returns 0 on success or a negative TWL_E_* code, never throws, never exits.
 * A context is bound to one CUDA device and is used from one host thread at a time (as the reference's
 * per-GPU GPU_pointers object, src/msa.hpp:220). There is NO CPU fallback: without a usable CUDA device
 * twl_init fails with TWL_E_NO_DEVICE.
 *
 * Alignment-path codes (int8): 0 = both advance, 1 = query-only column, 2 = reference-only column (src/msa.hpp:50,
 * SURVEY.md conventions). "ref" is the first node of a pair, "qry" the second.
 */
#ifndef TWILIGHT_B200_H
#define TWILIGHT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TWL_OK 0
#define TWL_E_NO_DEVICE (-1)   /* no CUDA device / wrong architecture */
#define TWL_E_CUDA (-2)        /* a CUDA call failed; see twl_last_error */
#define TWL_E_ARG (-3)         /* invalid argument */
#define TWL_E_NOMEM (-4)       /* device or pinned-host allocation failed */
#define TWL_E_STATE (-5)       /* call order violated (e.g. params not set) */

/* Per-pair status, mirrors Talco_xdrop errorType (src/TALCO-XDrop.cpp:248) */
#define TWL_ST_OK 0
#define TWL_ST_XDROP 1         /* band died under the x-drop rule (TALCO-XDrop.cpp:323-329) */
#define TWL_ST_BAND 2          /* anti-diagonal wider than fLen (TALCO-XDrop.cpp:331-338) */
#define TWL_ST_FATAL 3         /* index overrun (TALCO-XDrop.cpp:311-318, 658-667) */

typedef struct twl_ctx twl_ctx;

/* ---- lifecycle: replaces Option::getGpuInfo (src/cuda/gpu-info.cu:6-94) and the GPU_pointers
 *      allocate/free pair (src/cuda/alignment-gpu.cu:18-138) ------------------------------------------------ */
int twl_device_count(void);
int twl_init(int device, twl_ctx **out);
void twl_destroy(twl_ctx *ctx);
const char *twl_last_error(const twl_ctx *ctx);   /* ctx may be NULL: returns the last init error */

/* ---- scoring parameters: msa::Params (src/msa.hpp:98-109, src/scoring-matrix.cpp:81-137) and the
 *      Talco_xdrop::Params defaults derived from it (src/TALCO-XDrop.cpp:36-53). M = 5 (nucleotide, profile width
 *      6) or 21 (protein, profile width 22). score is M*M row-major and is copied. -------------------------- */
int twl_set_params(twl_ctx *ctx, const float *score, int M, float gap_open, float gap_extend, float gap_boundary);
/* marker = tile marker (default 1024, TALCO-XDrop.cpp:51); 1 <= marker <= 1024. */
int twl_set_marker(twl_ctx *ctx, int marker);

/* ---- level batch of profile pairs: the DP + traceback of every pair of one guide-tree level.
 *      Replaces the Talco_xdrop::Align_freq call of alignment-cpu.cpp:98-107 (and the kernel launch of
 *      alignment-gpu.cu:294-333). Semantics are those of the reference CPU path (float scores, marker 1024,
 *      fLen 4096, x-drop 1000*|gapExtend|), not of src/cuda. ------------------------------------------------ */
typedef struct {
    const float *freq_ref;      /* [ref_len][P] row-major column profile after gappy-column removal */
    const float *freq_qry;      /* [qry_len][P] */
    const float *gap_open_ref;  /* [ref_len] position-specific penalties (calculatePSGP, alignment-helper.cpp:168) */
    const float *gap_ext_ref;
    const float *gap_open_qry;  /* [qry_len] */
    const float *gap_ext_qry;
    int32_t ref_len, qry_len;   /* both >= 1 */
    float ref_num, qry_num;     /* sequences per side (denominator, TALCO-XDrop.cpp:269) */
    float gap_char_score;       /* gapExtend, or 0 for tasks 1/2 and >10000 sequences (alignment-cpu.cpp:88) */
    int32_t xdrop;              /* <=0: default 1000*|gapExtend| (TALCO-XDrop.cpp:49) */
    int32_t flen;               /* <=0: default 4096 (TALCO-XDrop.cpp:50) */
} twl_profile_pair;

typedef struct {
    int32_t status;             /* TWL_ST_* */
    int32_t path_len;           /* 0 when status != 0 */
    int32_t tiles;              /* TALCO tiles executed */
    int32_t reserved;
    uint64_t cells;             /* DP cell updates: sum over diagonals of (U-L+1), SURVEY.md §8(d) */
    uint64_t diagonals;
} twl_pair_result;

/* One call = host->device copy of the profiles, the DP/traceback kernels, device->host copy of the paths.
 * paths[p] must hold ref_len+qry_len bytes. */
int twl_align_profiles(twl_ctx *ctx, const twl_profile_pair *pairs, int n_pairs, int8_t *const *paths,
                       twl_pair_result *results);

/* The same work split in three so that a caller (bench.py, or a pipeline that overlaps levels) can keep the batch
 * resident in HBM: stage() packs + uploads, run() launches the kernels on the resident batch (may be called
 * repeatedly), fetch() downloads paths and results of the last run(). */
int twl_batch_stage(twl_ctx *ctx, const twl_profile_pair *pairs, int n_pairs);
int twl_batch_run(twl_ctx *ctx);
int twl_batch_fetch(twl_ctx *ctx, int8_t *const *paths, twl_pair_result *results);

/* ---- device-resident row store: mirror of SequenceDB::SequenceInfo (src/msa.hpp:113-134, src/sequencedb.cpp:8-76).
 *      A row is the current (aligned) text of one sequence; the level kernels read and rewrite rows in HBM, ping-pong
 *      between two buffers per row exactly like alnStorage[2] / changeStorage(). ids are small non-negative integers
 *      (SequenceInfo::id). ------------------------------------------------------------------------------------- */
int twl_rows_upload(twl_ctx *ctx, int n, const int32_t *ids, const char *const *rows, const int32_t *lens, const float *weights);
int twl_rows_download(twl_ctx *ctx, int n, const int32_t *ids, char *const *dst, int32_t *lens);   /* dst[i] holds >= current length */
/* Multi-GPU node migration without a host bounce: pack the current text of n rows into one contiguous
 * DEVICE buffer (row i at offsets[i], 16-byte aligned, offsets returned to the host) so it can be sent to another rank
 * over NVLink (NCCL send/recv, cudaMemcpyPeer), and create / overwrite n rows from such a buffer on the receiving context.
 * dev_dst must hold sum(align16(len)) bytes; both calls return with the work complete on the context's stream. */
int twl_rows_export(twl_ctx *ctx, int n, const int32_t *ids, void *dev_dst, size_t cap_bytes, int32_t *lens, int64_t *offsets);
int twl_rows_import(twl_ctx *ctx, int n, const int32_t *ids, const int32_t *lens, const float *weights, const void *dev_src,
                    const int64_t *offsets);
/* Multi-GPU inside ONE process (one context per device, as the reference's per-GPU GPU_pointers objects,
 * src/cuda/alignment-gpu.cu:226-253): move n rows from the context that holds them to another context's device — packed on the
 * source, one cudaMemcpyPeer over NVLink / NVSwitch (peer access is enabled on first use), unpacked on the destination; the
 * rows are gone from `src` afterwards. twl_rows_drop forgets rows (their buffers are reused by later rows). */
int twl_rows_migrate(twl_ctx *src, twl_ctx *dst, int n, const int32_t *ids);
int twl_rows_drop(twl_ctx *ctx, int n, const int32_t *ids);
int twl_rows_length(twl_ctx *ctx, int32_t id);
/* The same for n rows at once: lens[i] = current length of row ids[i], or -1. */
int twl_rows_lengths(twl_ctx *ctx, int n, const int32_t *ids, int32_t *lens);   /* current length of a row, <0 if unknown */
int twl_rows_clear(twl_ctx *ctx);

/* ---- one guide-tree level on the resident rows: for every pair calculateProfile + getConsensus +
 *      removeGappyColumns + calculatePSGP (device), TALCO-XDrop DP + traceback (device), addGappyColumnsBack (host,
 *      O(path) merge of run lists), updateFrequency + updateAlignment (device). Replaces the body of
 *      parallelAlignmentCPU (src/alignment-cpu.cpp:46-176) for pairs whose members are all resident rows. ------- */
typedef struct {
    const int32_t *seq_ids;     /* member rows in Node::seqsIncluded order */
    int32_t n_ids;
    int32_t aln_len;            /* Node::alnLen  */
    int32_t aln_num;            /* Node::alnNum  */
    float aln_weight;           /* Node::alnWeight */
    const float *msa_freq;      /* Node::msaFreq flattened [aln_len][P], or NULL when not cached */
} twl_node_side;

typedef struct {
    twl_node_side ref, qry;
    int32_t flags;              /* TWL_PAIR_PROFILE_ONLY: run calculateProfile (and its msaFreq caching) but no alignment; TWL_PAIR_NO_ROW_UPDATE */
    int32_t reserved;
} twl_level_pair;
#define TWL_PAIR_PROFILE_ONLY 1
#define TWL_PAIR_NO_ROW_UPDATE 2   /* align and merge msaFreq but leave the rows untouched (PLACE_WO_TREE, merge of sub-alignments) */

typedef struct {
    int32_t status;             /* TWL_ST_*; rows are rewritten only when 0 */
    int32_t path_len;           /* length of the final path (with gappy columns) = new Node::alnLen */
    int32_t tiles;
    int32_t cached;             /* bit 0: ref msaFreq was cached by this call (helper.cpp:35-40), bit 1: qry, bit 2: merged msaFreq available */
    uint64_t cells;
    uint64_t diagonals;
    int32_t ref_len_dp, qry_len_dp;   /* profile lengths after gappy-column removal */
} twl_level_result;

/* current_task: SequenceDB::currentTask (0 normal, 1 deferred re-alignment, 2 merge): selects gapCharScore and the
 * error protocol of alignment-cpu.cpp:88,108-129 (tasks 1/2 retry with wider x-drop / fLen inside this call).
 * gappy_threshold: Option::gappyVertical (0.95; 1.0 disables). cache_threshold: _CAL_PROFILE_TH (1000).
 * paths[p] (nullable) must hold ref.aln_len + qry.aln_len bytes. */
int twl_align_level(twl_ctx *ctx, const twl_level_pair *pairs, int n_pairs, int current_task, float gappy_threshold,
                    int32_t cache_threshold, int8_t *const *paths, twl_level_result *results);

/* Intermediates of the last twl_align_level call, for parity tests and for the caller's msaFreq bookkeeping. */
#define TWL_F_PROFILE_RAW_REF 0   /* float [aln_len][P] after calculateProfile */
#define TWL_F_PROFILE_RAW_QRY 1
#define TWL_F_CONSENSUS_REF 2     /* char [aln_len] */
#define TWL_F_CONSENSUS_QRY 3
#define TWL_F_RUNS_REF 4          /* int32 (start,len) pairs of removed gappy-column runs */
#define TWL_F_RUNS_QRY 5
#define TWL_F_PATH_WO 6           /* int8 path before addGappyColumnsBack */
#define TWL_F_FREQ_REF 7          /* float [aln_len][P] msaFreq cached for the ref node by this call */
#define TWL_F_FREQ_QRY 8
#define TWL_F_FREQ_MERGED 9       /* float [path_len][P] merged msaFreq (updateFrequency) */
#define TWL_F_DP_PROFILE_REF 10   /* float [ref_len_dp][P+2]: compacted columns + gapOpen + gapExtend (row-major view of the packed layout) */
#define TWL_F_DP_PROFILE_QRY 11
int twl_level_fetch(twl_ctx *ctx, int pair, int what, void *dst, size_t cap_bytes, size_t *out_bytes);

/* Device time of the last twl_align_level call by phase, milliseconds: [0] profile build, [1] gappy/PSGP/pack,
 * [2] DP chain, [3] row update + frequency merge. */
int twl_level_phase_ms(twl_ctx *ctx, float out[4]);
/* Phase [3] split in two: out[0] = gappy-column restore (addGappyColumnsBack: a sequential merge per pair, latency bound),
 * out[1] = path prefix counts + row rewrite + frequency merge (updateAlignment / updateFrequency: the HBM-bound part). */
int twl_level_update_split_ms(twl_ctx *ctx, float out[2]);

/* addGappyColumnsBack (alignment-helper.cpp:324-375) runs on the device. When removed runs of BOTH nodes start at the same
 * path position the reference aligns their consensus substrings (pairwiseGlobal, alignment-helper.cpp:243-322); the kernel
 * keeps that alignment in shared memory, and a pair with a run pair of more than 8192 matrix cells is redone by a second
 * pass of the same kernel with its matrices in a global scratch buffer. Returns how many pairs took the second pass since
 * twl_init (diagnostics; 0 on ordinary data). */
int twl_level_large_restores(const twl_ctx *ctx);

/* Device time (CUDA events on the context's stream) of the kernels of the last run(), in milliseconds, and the number
 * of kernel launches it issued. */
float twl_last_kernel_ms(const twl_ctx *ctx);
int twl_last_launch_count(const twl_ctx *ctx);

/* Tuning / diagnostics switches; every setting produces identical results. "force_generic" = 1 routes nucleotide batches
 * through the wide-band generic kernel instead of the register-resident wavefront kernel (A/B parity tests). "wide_workers"
 * (default 8; 0 = run the wide-band kernel after the narrow one instead of beside it), "latency_mode" (-1 auto, 0 off,
 * 1 always: the one-CTA-per-SM shape for levels with few pairs), "latency_shape" (2 = 512 threads x 2 rows, default;
 * 3 = 512 x 1 first and 512 x 2 for pairs whose band outgrows 512 rows), "max_ctas_per_sm" (occupancy experiments),
 * "protein_sim" (1 default: proteins compute every pair's similarity matrix with a dependency-free kernel and run the recurrence
 * on the register-resident wavefront kernel; 0 = the generic kernel scores on the fly), "sim_budget_mb" (default 16384: a chain
 * whose matrices need more falls back to the generic kernel),
 * "dp_trace" = 1 prints per-stage times to stderr. "inject_nomem" = n makes the n-th following twl_align_level call fail with
 * TWL_E_NOMEM after its kernels ran (tests of the all-or-nothing rollback and of the adapter's spill-and-retry path).
 * The environment variable TWL_OPTIONS="name=value,name=value" applies the same switches at twl_init. Environment only:
 * TWL_RESTORE_JOBS=0 aligns coinciding gappy runs in line during the restore walk instead of putting them off to the parallel job
 * kernel (A/B); TWL_LEVEL_BUDGET_MB caps the scratch of a level chunk; TWL_TRACE=1 prints the host-side steps of every level chunk. */
int twl_set_option(twl_ctx *ctx, const char *name, int value);

/* Device self-test: evaluates the reciprocal-based exact division used by the DP kernels and the IEEE divide on n
 * operand pairs and reports on how many they differ bit-wise (must be 0). */
int twl_selftest_division(twl_ctx *ctx, const float *num, const float *den, int n, int *mismatches);

/* Library build information (arch string etc.). */
const char *twl_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TWILIGHT_B200_H */

/* oracle/twl_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the TWILIGHT per-level alignment path. It exists to check the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it. The product
 * (twilight_b200/) never links or calls it.
 *
 * Parity status: PINNED. tests/test_oracle_vs_reference.py compares every function here with the unmodified
 * reference compiled into oracle/_ref/libtalco_ref.so (oracle/ref_shim.cpp) on the bundled sars_20 / RNASim
 * data and on seeded synthetic inputs; tests/golden/ holds committed vectors generated from that library by
 * tests/golden/make_golden.py.
 *
 * Floating point: the port reproduces the operation order of the x86 reference build (TALCO_SIMD, GCC
 * -ffp-contract=fast): products of the 5x5 / 21x21 contraction are rounded individually, the gap-character terms
 * and the tile-0 edge term are fused multiply-adds. It must be compiled with -ffp-contract=off; fused operations are
 * written out as fmaf().
 */
#ifndef TWL_ORACLE_H
#define TWL_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t P;             /* profile width: 6 (nucleotide) or 22 (protein); matrix is (P-1)x(P-1) row-major */
    const float *score;    /* substitution matrix, msa::Params::scoringMatrix (scoring-matrix.cpp:81-137) */
    float gapOpen;         /* Talco_xdrop::Params (TALCO-XDrop.cpp:36-53) */
    float gapExtend;
    float gapBoundary;
    float gapCharScore;    /* gapExtend, or 0 for tasks 1/2 or >10000 sequences (alignment-cpu.cpp:88) */
    int32_t xdrop;         /* 1000*|gapExtend| = 5000 by default */
    int32_t fLen;          /* anti-diagonal cap, 4096 */
    int32_t marker;        /* tile marker, 1024 */
} twlo_talco_params;

/* Talco_xdrop::Align_freq + Tile + Traceback + Reduction_tree (TALCO-XDrop.cpp:62-689).
 * freq* are [len][P] row-major, gap* are [len]. aln must hold refLen+qryLen bytes.
 * Returns the path length (0 on error); *errorType as in the reference (0 ok, 1 x-drop died, 2 band > fLen,
 * 3 index overrun). cells/tiles/diagonals (nullable) receive the work counters of SURVEY.md §8(d). */
int twlo_talco_align(const twlo_talco_params *p, int refLen, int qryLen, const float *freqRef, const float *freqQry,
                     const float *gapOpRef, const float *gapExRef, const float *gapOpQry, const float *gapExQry,
                     float refNum, float qryNum, int8_t *aln, int *errorType, uint64_t *cells, int32_t *tiles,
                     uint64_t *diagonals);

/* letterIdx (scoring-matrix.cpp:26-79) after toupper. type 'n' or 'p'. */
int twlo_letter_index(char type, char c);

/* calculateProfile, row-accumulating branch (alignment-helper.cpp:23-34): profile[t][letter] += w for each member
 * row in order, w = weight/nodeWeight*alnNum. profile is [alnLen][P], must be zeroed by the caller. */
void twlo_profile_from_rows(char type, int nSeq, const char *const *rows, const float *weights, int alnLen,
                            int alnNum, float nodeWeight, float *profile);
/* calculateProfile, cached branch (alignment-helper.cpp:16-21): profile = msaFreq / nodeWeight * alnNum. */
void twlo_profile_from_freq(int P, const float *msaFreq, int alnLen, int alnNum, float nodeWeight, float *profile);
/* the msaFreq cache written by calculateProfile (alignment-helper.cpp:35-40): profile / alnNum * nodeWeight. */
void twlo_freq_from_profile(int P, const float *profile, int alnLen, int alnNum, float nodeWeight, float *msaFreq);
/* getConsensus (alignment-helper.cpp:221-241). */
void twlo_consensus(int P, const float *profile, int len, char *out);
/* removeGappyColumns for one side (alignment-helper.cpp:74-166). profile is compacted in place, tail zeroed.
 * runs receives (start,len) pairs, returns the number of runs; *newLen the compacted length. */
int twlo_remove_gappy(int P, float *profile, int len, int alnNum, float threshold, int32_t *runs, int *newLen);
/* calculatePSGP for one side (alignment-helper.cpp:168-219). */
void twlo_psgp(int P, const float *profile, int len, int alnNum, float gapOpen, float gapExtend, float *gapOp,
               float *gapEx);
/* pairwiseGlobal (alignment-helper.cpp:243-322) on two consensus strings; returns path length. */
int twlo_pairwise_global(char type, const float *score, float gapOpen, float gapExtend, const char *s1, int m,
                         const char *s2, int n, int8_t *path);
/* addGappyColumnsBack (alignment-helper.cpp:324-375). out must hold refLen+qryLen bytes (original lengths). */
int twlo_add_gappy_back(char type, const float *score, float gapOpen, float gapExtend, const int8_t *aln, int alnLen,
                        const int32_t *runsRef, int nRunsRef, const int32_t *runsQry, int nRunsQry,
                        const char *consRef, const char *consQry, int8_t *out);
/* updateFrequency (alignment-helper.cpp:506-539): merged [pathLen][P]. */
void twlo_merge_freq(int P, const float *freqRef, const float *freqQry, const int8_t *path, int pathLen,
                     float refWeight, float qryWeight, float *merged);
/* updateAlignment row rewrite (alignment-helper.cpp:381-401, 428-448): side 0 = ref member (copies on 0/2),
 * side 1 = qry member (copies on 0/1). */
void twlo_update_row(int side, const char *row, const int8_t *path, int pathLen, char *out);

#ifdef __cplusplus
}
#endif
#endif

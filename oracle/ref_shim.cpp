// oracle/ref_shim.cpp — TEST INFRASTRUCTURE ONLY (never on the product path).
//
// A flat C entry surface over the *unmodified* reference functions, compiled together with the reference sources
// (where they lie under /root/reference/src) into oracle/_ref/libtalco_ref.so by oracle/Makefile. It is used to
//   * pin the CPU restatement (oracle/twl_oracle.cpp) and the CUDA path at function boundaries, and
//   * time the reference's own DP (`bench.py --impl reference`, cpu_baseline.kind == "reference").
// Every function below only marshals flat arrays into the reference's own types and calls the reference:
//   ref_talco_align     -> Talco_xdrop::Align_freq                      (src/TALCO-XDrop.cpp:62)
//   ref_pair_pipeline   -> the per-pair sequence of src/alignment-cpu.cpp:49-170 built from
//                          alignment_helper::{calculateProfile,getConsensus,removeGappyColumns,calculatePSGP,
//                          addGappyColumnsBack,updateFrequency,updateAlignment} (src/alignment-helper.cpp)
//   ref_level_cpu       -> msa::progressive::cpu::parallelAlignmentCPU  (src/alignment-cpu.cpp:36) on a level
#include "msa.hpp"
#include "TALCO-XDrop.hpp"
#include <tbb/parallel_for.h>

#include <chrono>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

// msa::Params can only be constructed from a variables_map (scoring-matrix.cpp:81); build one with the defaults the
// CLI declares (twilight-main.cpp:62-72) and then overwrite the public fields with the caller's values.
msa::Params *makeParams(char type, const float *score, float gapOpen, float gapExtend, float gapBoundary) {
    po::variables_map vm;
    auto put = [&](const char *k, float v) { vm.set(k, po::variable_value(std::make_shared<float>(v), &typeid(float))); };
    put("gap-open", gapOpen);
    put("gap-extend", gapExtend);
    put("gap-ends", gapBoundary);
    put("xdrop", 600.0f);
    put("match", 18.0f);
    put("mismatch", -8.0f);
    put("transition", -4.0f);
    vm.set("blosum", po::variable_value(std::make_shared<int>(62), &typeid(int)));
    msa::Params *p = new msa::Params(vm, type);
    if (score) {
        for (int i = 0; i < p->matrixSize; ++i)
            for (int j = 0; j < p->matrixSize; ++j) p->scoringMatrix[i][j] = score[i * p->matrixSize + j];
    }
    return p;
}

// msa::Option's only constructor parses the command line and touches the file system (option.cpp:12); the hot path
// reads just five POD fields (type, gappyVertical, alnMode, printDetail, noFilter), so hand it zeroed storage.
struct OptionBox {
    alignas(msa::Option) unsigned char raw[sizeof(msa::Option)];
    msa::Option *get() { return reinterpret_cast<msa::Option *>(raw); }
    OptionBox(char type, float gappyVertical, int alnMode) {
        std::memset(raw, 0, sizeof(raw));
        get()->type = type;
        get()->gappyVertical = gappyVertical;
        get()->alnMode = alnMode;
        get()->printDetail = false;
        get()->noFilter = true;
        get()->cpuNum = 1;
    }
};

struct SideIn {
    int nSeq;
    const char *const *rows; // nSeq rows of alnLen chars
    const float *weights;    // per-sequence weight
    int alnLen, alnNum;
    float alnWeight;
    const float *msaFreq; // nullable [alnLen][P]
};

void fillNode(msa::Node *node, msa::SequenceDB *db, const SideIn &s, int P, int &nextId, bool debug) {
    for (int n = 0; n < s.nSeq; ++n) {
        std::string row(s.rows[n], s.alnLen);
        std::string name = node->identifier + "_" + std::to_string(n);
        db->addSequence(nextId, name, row, -1, s.weights[n], debug, 0);
        node->seqsIncluded.push_back(nextId);
        ++nextId;
    }
    node->alnLen = s.alnLen;
    node->alnNum = s.alnNum;
    node->alnWeight = s.alnWeight;
    if (s.msaFreq) {
        node->msaFreq.assign(s.alnLen, std::vector<float>(P, 0.f));
        for (int t = 0; t < s.alnLen; ++t)
            for (int v = 0; v < P; ++v) node->msaFreq[t][v] = s.msaFreq[t * P + v];
    }
}

} // namespace

extern "C" {

// Talco_xdrop::Align_freq (TALCO-XDrop.cpp:62). aln must hold refLen+qryLen bytes. Returns the path length
// (0 when the reference cleared the path), *errorType as set by the reference.
int ref_talco_align(int P, int refLen, int qryLen, const float *freqRef, const float *freqQry, const float *gapOpRef,
                    const float *gapExRef, const float *gapOpQry, const float *gapExQry, float refNum, float qryNum,
                    const float *score, float gapOpen, float gapExtend, float gapBoundary, float gapCharScore,
                    int xdrop, int fLen, int marker, int8_t *aln, int *errorType) {
    char type = (P == 6) ? 'n' : 'p';
    msa::Params *mp = makeParams(type, score, gapOpen, gapExtend, gapBoundary);
    Talco_xdrop::Params tp(*mp);
    tp.gapCharScore = gapCharScore;
    if (xdrop > 0) tp.xdrop = xdrop;
    if (fLen > 0) tp.fLen = fLen;
    if (marker > 0) tp.marker = marker;
    std::vector<std::vector<float>> fr(refLen, std::vector<float>(P)), fq(qryLen, std::vector<float>(P));
    for (int s = 0; s < refLen; ++s) for (int t = 0; t < P; ++t) fr[s][t] = freqRef[s * P + t];
    for (int s = 0; s < qryLen; ++s) for (int t = 0; t < P; ++t) fq[s][t] = freqQry[s * P + t];
    std::vector<std::vector<float>> gOp(2), gEx(2);
    gOp[0].assign(gapOpRef, gapOpRef + refLen);
    gEx[0].assign(gapExRef, gapExRef + refLen);
    gOp[1].assign(gapOpQry, gapOpQry + qryLen);
    gEx[1].assign(gapExQry, gapExQry + qryLen);
    std::vector<int8_t> path;
    int16_t err = 0;
    Talco_xdrop::Align_freq(&tp, fr, fq, gOp, gEx, std::make_pair(refNum, qryNum), path, err);
    *errorType = err;
    for (size_t i = 0; i < path.size(); ++i) aln[i] = path[i];
    delete mp;
    return static_cast<int>(path.size());
}

// The per-pair pipeline of alignment-cpu.cpp:49-170 with every intermediate exposed.
// Output buffers (caller-allocated; memLen = max(refLen, qryLen), cap = refLen + qryLen):
//   profileRaw   [2][memLen][P]  after calculateProfile
//   consensus    [2][memLen]     after getConsensus (bytes, not NUL-terminated)
//   profile      [2][memLen][P]  after removeGappyColumns
//   lensOut      [2]             lens after removal
//   gappyRuns    [2][memLen][2]  (start,len) runs; gappyCount[2]
//   gapOp/gapEx  [2][memLen]
//   alnWo/alnW   [cap]           path without / with gappy columns; alnWoLen/alnWLen
//   newRows      [(nRef+nQry)][cap] rows after updateAlignment (ref members first), newLen
//   cachedFreq   [2][memLen][P] msaFreq cached by calculateProfile (valid iff cachedFlag[s])
//   mergedFreq   [cap][P]        first->msaFreq after updateFrequency (valid iff mergedFlag)
int ref_pair_pipeline(char type, int currentTask, float gappyVertical, const float *score, float gapOpen,
                      float gapExtend, float gapBoundary, int marker,
                      int nRef, const char *const *rowsRef, const float *wRef, int refLen, int refNum, float refWeight, const float *refFreq,
                      int nQry, const char *const *rowsQry, const float *wQry, int qryLen, int qryNum, float qryWeight, const float *qryFreq,
                      float *profileRaw, char *consensus, float *profile, int *lensOut, int *gappyRuns, int *gappyCount,
                      float *gapOp, float *gapEx, int8_t *alnWo, int *alnWoLen, int8_t *alnW, int *alnWLen,
                      int *errorTypeOut, char *newRows, int *newLen, float *cachedFreq, int *cachedFlag,
                      float *mergedFreq, int *mergedFlag) {
    using namespace msa;
    const int P = (type == 'n') ? 6 : 22;
    Params *mp = makeParams(type, score, gapOpen, gapExtend, gapBoundary);
    OptionBox ob(type, gappyVertical, DEFAULT_ALN);
    Option *option = ob.get();
    SequenceDB *db = new SequenceDB();
    db->currentTask = currentTask;
    Node *a = new Node("refnode", 0.f), *b = new Node("qrynode", 0.f);
    int nextId = 0;
    SideIn sr{nRef, rowsRef, wRef, refLen, refNum, refWeight, refFreq};
    SideIn sq{nQry, rowsQry, wQry, qryLen, qryNum, qryWeight, qryFreq};
    fillNode(a, db, sr, P, nextId, false);
    fillNode(b, db, sq, P, nextId, false);
    NodePair pair(a, b);

    const int memLen = std::max(refLen, qryLen);
    float *hostFreq, *hostGapOp, *hostGapEx;
    progressive::cpu::allocateMemory_and_Initialize(hostFreq, hostGapOp, hostGapEx, memLen, P);
    std::pair<IntPairVec, IntPairVec> gappyColumns;
    stringPair cons({"", ""});
    IntPair lens = {refLen, qryLen};
    alignment_helper::calculateProfile(hostFreq, pair, db, option, memLen);
    std::memcpy(profileRaw, hostFreq, sizeof(float) * 2 * memLen * P);
    cachedFlag[0] = (!refFreq && !a->msaFreq.empty());
    cachedFlag[1] = (!qryFreq && !b->msaFreq.empty());
    for (int s = 0; s < 2; ++s) {
        Node *nd = s ? b : a;
        if (cachedFlag[s])
            for (int t = 0; t < nd->alnLen; ++t)
                for (int v = 0; v < P; ++v) cachedFreq[(s * memLen + t) * P + v] = nd->msaFreq[t][v];
    }
    alignment_helper::getConsensus(option, hostFreq, cons.first, refLen);
    alignment_helper::getConsensus(option, hostFreq + P * memLen, cons.second, qryLen);
    std::memcpy(consensus, cons.first.data(), refLen);
    std::memcpy(consensus + memLen, cons.second.data(), qryLen);
    alignment_helper::removeGappyColumns(hostFreq, pair, option, gappyColumns, memLen, lens, currentTask);
    alignment_helper::calculatePSGP(hostFreq, hostGapOp, hostGapEx, pair, db, option, memLen, {0, 0}, lens, *mp);
    std::memcpy(profile, hostFreq, sizeof(float) * 2 * memLen * P);
    std::memcpy(gapOp, hostGapOp, sizeof(float) * 2 * memLen);
    std::memcpy(gapEx, hostGapEx, sizeof(float) * 2 * memLen);
    lensOut[0] = lens.first;
    lensOut[1] = lens.second;
    gappyCount[0] = gappyColumns.first.size();
    gappyCount[1] = gappyColumns.second.size();
    for (size_t g = 0; g < gappyColumns.first.size(); ++g) {
        gappyRuns[2 * g] = gappyColumns.first[g].first;
        gappyRuns[2 * g + 1] = gappyColumns.first[g].second;
    }
    for (size_t g = 0; g < gappyColumns.second.size(); ++g) {
        gappyRuns[2 * (memLen + g)] = gappyColumns.second[g].first;
        gappyRuns[2 * (memLen + g) + 1] = gappyColumns.second[g].second;
    }

    // alignment-cpu.cpp:70-133 (single attempt; the retry ladder is the caller's business)
    std::vector<int8_t> aln_wo_gc;
    Profile freqRef(lens.first, std::vector<float>(P, 0.0)), freqQry(lens.second, std::vector<float>(P, 0.0));
    Profile gOp(2), gEx(2);
    for (int s = 0; s < lens.first; s++) for (int t = 0; t < P; ++t) freqRef[s][t] = hostFreq[P * s + t];
    for (int s = 0; s < lens.second; s++) for (int t = 0; t < P; ++t) freqQry[s][t] = hostFreq[P * (memLen + s) + t];
    for (int r = 0; r < lens.first; ++r) { gOp[0].push_back(hostGapOp[r]); gEx[0].push_back(hostGapEx[r]); }
    for (int q = 0; q < lens.second; ++q) { gOp[1].push_back(hostGapOp[memLen + q]); gEx[1].push_back(hostGapEx[memLen + q]); }
    progressive::cpu::freeMemory(hostFreq, hostGapOp, hostGapEx);
    Talco_xdrop::Params tp(*mp);
    if (marker > 0) tp.marker = marker;
    if (currentTask == 1 || currentTask == 2 || refNum > 10000 || qryNum > 10000) tp.gapCharScore = 0;
    int16_t err = 0;
    if (lens.first == 0) for (int j = 0; j < lens.second; ++j) aln_wo_gc.push_back(1);
    if (lens.second == 0) for (int j = 0; j < lens.first; ++j) aln_wo_gc.push_back(2);
    if (aln_wo_gc.empty())
        Talco_xdrop::Align_freq(&tp, freqRef, freqQry, gOp, gEx, std::make_pair((float)refNum, (float)qryNum), aln_wo_gc, err);
    *errorTypeOut = err;
    *alnWoLen = aln_wo_gc.size();
    for (size_t i = 0; i < aln_wo_gc.size(); ++i) alnWo[i] = aln_wo_gc[i];
    *alnWLen = 0;
    *newLen = 0;
    *mergedFlag = 0;
    if (!aln_wo_gc.empty()) {
        alnPath aln_w_gc;
        int alnRef = 0, alnQry = 0;
        for (auto op : aln_wo_gc) { if (op == 0) { alnRef++; alnQry++; } if (op == 1) alnQry++; if (op == 2) alnRef++; }
        alignment_helper::addGappyColumnsBack(aln_wo_gc, aln_w_gc, gappyColumns, *mp, {alnRef, alnQry}, cons);
        *alnWLen = aln_w_gc.size();
        for (size_t i = 0; i < aln_w_gc.size(); ++i) alnW[i] = aln_w_gc[i];
        float rw = a->alnWeight, qw = b->alnWeight;
        alignment_helper::updateFrequency(pair, db, aln_w_gc, {rw, qw});
        alignment_helper::updateAlignment(pair, db, option, aln_w_gc);
        const int cap = refLen + qryLen;
        *newLen = a->alnLen;
        for (int n = 0; n < nRef + nQry; ++n) {
            auto *si = db->sequences[n];
            std::memcpy(newRows + (size_t)n * cap, si->alnStorage[si->storage], si->len);
        }
        if (!a->msaFreq.empty()) {
            *mergedFlag = 1;
            for (size_t t = 0; t < a->msaFreq.size(); ++t)
                for (int v = 0; v < P; ++v) mergedFreq[t * P + v] = a->msaFreq[t][v];
        }
    }
    for (auto *si : db->sequences) delete si;
    delete db;
    delete a;
    delete b;
    delete mp;
    return 0;
}

// One guide-tree level through the reference's own level entry point, msa::progressive::cpu::alignmentKernel_CPU ->
// parallelAlignmentCPU (src/alignment-cpu.cpp:32-183): the stock call the scheduler makes at src/progressive.cpp:180, on a
// real NodePairVec / SequenceDB built from flat inputs. Its tbb::parallel_for over the pairs runs on `threads` workers.
//   nSeq/alnLen/alnNum/alnWeight  [2*nPairs]  (ref side at 2p, qry side at 2p+1); rows/weights: all member rows in that order
// Outputs: newLen[p] = first->alnLen after the call (0 for a deferred pair); *seconds = wall time of the level call alone
// (node / DB construction excluded). Returns the number of pairs the reference deferred (db->fallback_nodes).
int ref_level_cpu(char type, int currentTask, float gappyVertical, const float *score, float gapOpen, float gapExtend,
                  float gapBoundary, int nPairs, const int *nSeq, const char *const *rows, const float *weights,
                  const int *alnLen, const int *alnNum, const float *alnWeight, int threads, int *newLen, double *seconds) {
    using namespace msa;
    const int P = (type == 'n') ? 6 : 22;
    Params *mp = makeParams(type, score, gapOpen, gapExtend, gapBoundary);
    OptionBox ob(type, gappyVertical, 0);
    ob.get()->cpuNum = threads;
    SequenceDB *db = new SequenceDB();
    db->currentTask = currentTask;
    std::vector<Node *> owned;
    NodePairVec level;
    int nextId = 0;
    size_t at = 0;
    for (int p = 0; p < nPairs; ++p) {
        Node *nd[2];
        for (int s = 0; s < 2; ++s) {
            const int k = 2 * p + s;
            nd[s] = new Node((s ? "q" : "r") + std::to_string(p), 0.f);
            owned.push_back(nd[s]);
            SideIn in{nSeq[k], rows + at, weights + at, alnLen[k], alnNum[k], alnWeight[k], nullptr};
            fillNode(nd[s], db, in, P, nextId, false);
            at += static_cast<size_t>(nSeq[k]);
        }
        level.push_back(std::make_pair(nd[0], nd[1]));
    }
    const int before = tbb::compat_detail::parallelism_cap();
    tbb::compat_detail::parallelism_cap() = threads < 1 ? 1 : threads;
    const auto t0 = std::chrono::steady_clock::now();
    progressive::cpu::alignmentKernel_CPU(nullptr, level, db, ob.get(), *mp);
    const auto t1 = std::chrono::steady_clock::now();
    tbb::compat_detail::parallelism_cap() = before;
    if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
    const int deferred = static_cast<int>(db->fallback_nodes.size());
    for (int p = 0; p < nPairs; ++p)
        if (newLen) newLen[p] = level[p].second->seqsIncluded.empty() ? level[p].first->alnLen : 0;
    for (auto *si : db->sequences) delete si;
    delete db;
    for (Node *n : owned) delete n;
    delete mp;
    return deferred;
}

// Worker cap for the compat parallel_for (the reference does this with tbb::global_control, twilight-main.cpp:117).
void ref_set_threads(int n) { tbb::compat_detail::parallelism_cap() = n < 1 ? 1 : n; }

} // extern "C"

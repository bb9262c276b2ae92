// alignment_b200.cpp — the reference-side binding: an `msa::alnFunction` (src/msa.hpp:175) that runs a guide-tree
// level through libtwilight_b200.so. It is compiled together with the UNCHANGED TWILIGHT host sources (option parsing,
// tree, partitioning, sequence DB, I/O, progressive scheduler) and passed to msaOnSubtree() exactly where the
// reference passes cpu::alignmentKernel_CPU (src/twilight-main.cpp:148,183,201,220,261,299; src/progressive.cpp:291).
//
// Division of labour in this revision: the profile/PSGP preparation and the row update use the reference's own
// alignment_helper functions on the host (as the reference GPU build does, src/cuda/alignment-gpu.cu:261-288,
// 335-420, except that removeGappyColumns IS applied, as on the CPU path); the TALCO-XDrop DP + traceback of every
// pair of the level runs on the B200 through twl_align_profiles(). The per-pair protocol (trivial pairs, low-quality
// singletons, errorType handling, retry ladder for tasks 1/2, fallback2cpu) follows src/alignment-cpu.cpp:86-181.
//
// There is no CPU alignment fallback: if the CUDA library cannot initialise, the process aborts with a message.
#ifndef MSA_HPP
#include "msa.hpp"
#endif
#include "twilight_b200.h"

#include <tbb/parallel_for.h>
#include <tbb/spin_rw_mutex.h>

#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

namespace msa {
namespace progressive {
namespace b200 {

namespace {

struct Device {
    twl_ctx *ctx = nullptr;
    int M = 0;
    float gapOpen = 0, gapExtend = 0, gapBoundary = 0;
    std::vector<float> score;
};

Device &device() {
    static Device d;
    return d;
}

[[noreturn]] void die(const char *what, const char *detail) {
    std::cerr << "twilight-b200: " << what << ": " << detail << "\n";
    std::exit(1);
}

void ensureContext(Params &param) {
    Device &d = device();
    if (!d.ctx) {
        const char *env = std::getenv("TWL_DEVICE");
        const int dev = env ? std::atoi(env) : 0;
        if (twl_init(dev, &d.ctx) != TWL_OK) die("cannot initialise the CUDA device", twl_last_error(nullptr));
        std::atexit([] { if (device().ctx) { twl_destroy(device().ctx); device().ctx = nullptr; } });
    }
    const int M = param.matrixSize;
    std::vector<float> flat(static_cast<size_t>(M) * M);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) flat[i * M + j] = param.scoringMatrix[i][j];
    if (d.M != M || d.gapOpen != param.gapOpen || d.gapExtend != param.gapExtend || d.gapBoundary != param.gapBoundary || d.score != flat) {
        if (twl_set_params(d.ctx, flat.data(), M, param.gapOpen, param.gapExtend, param.gapBoundary) != TWL_OK)
            die("twl_set_params", twl_last_error(d.ctx));
        d.M = M; d.gapOpen = param.gapOpen; d.gapExtend = param.gapExtend; d.gapBoundary = param.gapBoundary; d.score = flat;
    }
}

// Everything one pair needs between the host-side preparation and the host-side update.
struct PairWork {
    float *freq = nullptr, *gapOp = nullptr, *gapEx = nullptr;
    int memLen = 0;
    std::pair<IntPairVec, IntPairVec> gappyColumns;
    stringPair consensus{"", ""};
    IntPair lens{0, 0};
    int32_t refLen = 0, qryLen = 0, refNum = 0, qryNum = 0;
    bool lowQ = false;
    bool onDevice = false;      // has a DP to run
    int32_t xdrop = 0, fLen = 0; // Talco_xdrop::Params state of the retry ladder (0 = default)
    float gapCharScore = 0;
    std::vector<int8_t> aln;    // aln_wo_gc
    void release() {
        if (freq) cpu::freeMemory(freq, gapOp, gapEx);
        freq = gapOp = gapEx = nullptr;
    }
};

} // namespace

void alignmentKernel_B200(Tree *tree, NodePairVec &nodes, SequenceDB *database, Option *option, Params &param) {
    ensureContext(param);
    twl_ctx *ctx = device().ctx;
    const int profileSize = param.matrixSize + 1;
    const int task = database->currentTask;
    tbb::spin_rw_mutex fallbackMutex;
    std::vector<int> fallbackPairs;

    // Bound host memory: prepare/align/apply the level in chunks of pairs (profiles are 2*memLen*P floats per pair).
    const size_t chunkBudgetBytes = static_cast<size_t>(3) << 30;
    size_t begin = 0;
    while (begin < nodes.size()) {
        size_t end = begin, bytes = 0;
        while (end < nodes.size()) {
            const size_t memLen = std::max(nodes[end].first->getAlnLen(task), nodes[end].second->getAlnLen(task));
            const size_t need = memLen * (profileSize + 2) * 2 * sizeof(float);
            if (end > begin && bytes + need > chunkBudgetBytes) break;
            bytes += need;
            ++end;
        }
        const int count = static_cast<int>(end - begin);
        std::vector<PairWork> work(count);

        // ---- host preparation, alignment-cpu.cpp:49-92 -------------------------------------------------------
        tbb::parallel_for(tbb::blocked_range<int>(0, count), [&](tbb::blocked_range<int> range) {
        for (int w = range.begin(); w < range.end(); ++w) {
            const int nIdx = static_cast<int>(begin) + w;
            PairWork &pw = work[w];
            pw.refLen = nodes[nIdx].first->getAlnLen(task);
            pw.qryLen = nodes[nIdx].second->getAlnLen(task);
            pw.refNum = nodes[nIdx].first->getAlnNum(task);
            pw.qryNum = nodes[nIdx].second->getAlnNum(task);
            pw.memLen = std::max(pw.refLen, pw.qryLen);
            cpu::allocateMemory_and_Initialize(pw.freq, pw.gapOp, pw.gapEx, pw.memLen, profileSize);
            pw.lens = {pw.refLen, pw.qryLen};
            alignment_helper::calculateProfile(pw.freq, nodes[nIdx], database, option, pw.memLen);
            alignment_helper::getConsensus(option, pw.freq, pw.consensus.first, pw.refLen);
            alignment_helper::getConsensus(option, pw.freq + profileSize * pw.memLen, pw.consensus.second, pw.qryLen);
            alignment_helper::removeGappyColumns(pw.freq, nodes[nIdx], option, pw.gappyColumns, pw.memLen, pw.lens, task);
            alignment_helper::calculatePSGP(pw.freq, pw.gapOp, pw.gapEx, nodes[nIdx], database, option, pw.memLen, {0, 0}, pw.lens, param);
            pw.gapCharScore = (task == 1 || task == 2 || pw.refNum > 10000 || pw.qryNum > 10000) ? 0.0f : param.gapExtend;
            // NB: the reference tests the ORIGINAL lengths here (alignment-cpu.cpp:89-90)
            if (pw.refLen == 0) pw.aln.assign(pw.qryLen, 1);
            if (pw.qryLen == 0) pw.aln.assign(pw.refLen, 2);
            const bool lowQ_r = (option->alnMode == MERGE_MSA) ? false : ((pw.refNum > 1) ? false : database->sequences[nodes[nIdx].first->seqsIncluded[0]]->lowQuality);
            const bool lowQ_q = (option->alnMode == MERGE_MSA) ? false : ((pw.qryNum > 1) ? false : database->sequences[nodes[nIdx].second->seqsIncluded[0]]->lowQuality);
            pw.lowQ = lowQ_r || lowQ_q;
            pw.onDevice = !pw.lowQ && pw.aln.empty();
        }
        });

        // ---- TALCO-XDrop on the device, with the retry ladder of alignment-cpu.cpp:95-130 ---------------------
        std::vector<int> pending;
        for (int w = 0; w < count; ++w) if (work[w].onDevice) pending.push_back(w);
        while (!pending.empty()) {
            std::vector<twl_profile_pair> batch(pending.size());
            std::vector<int8_t *> paths(pending.size());
            std::vector<twl_pair_result> results(pending.size());
            for (size_t b = 0; b < pending.size(); ++b) {
                PairWork &pw = work[pending[b]];
                if (pw.lens.first < 1 || pw.lens.second < 1) die("empty profile after gappy-column removal", "unsupported input");
                twl_profile_pair &tp = batch[b];
                tp.freq_ref = pw.freq;
                tp.freq_qry = pw.freq + static_cast<size_t>(profileSize) * pw.memLen;
                tp.gap_open_ref = pw.gapOp;
                tp.gap_ext_ref = pw.gapEx;
                tp.gap_open_qry = pw.gapOp + pw.memLen;
                tp.gap_ext_qry = pw.gapEx + pw.memLen;
                tp.ref_len = pw.lens.first;
                tp.qry_len = pw.lens.second;
                tp.ref_num = static_cast<float>(pw.refNum);
                tp.qry_num = static_cast<float>(pw.qryNum);
                tp.gap_char_score = pw.gapCharScore;
                tp.xdrop = pw.xdrop;
                tp.flen = pw.fLen;
                pw.aln.assign(static_cast<size_t>(pw.lens.first) + pw.lens.second, 0);
                paths[b] = pw.aln.data();
            }
            if (twl_align_profiles(ctx, batch.data(), static_cast<int>(batch.size()), paths.data(), results.data()) != TWL_OK)
                die("twl_align_profiles", twl_last_error(ctx));
            std::vector<int> again;
            for (size_t b = 0; b < pending.size(); ++b) {
                const int w = pending[b];
                PairWork &pw = work[w];
                const int err = results[b].status;
                pw.aln.resize(err == 0 ? results[b].path_len : 0);
                if (err == 0) continue;
                if (task == 0) {                                            // :108-115 defer the pair
                    fallbackPairs.push_back(static_cast<int>(begin) + w);
                    continue;
                }
                const int32_t curX = pw.xdrop > 0 ? pw.xdrop : static_cast<int32_t>(1000 * -1 * param.gapExtend);
                const int32_t curF = pw.fLen > 0 ? pw.fLen : (1 << 12);
                const int32_t minLen = std::min(pw.lens.first, pw.lens.second);
                if (err == 2) {                                             // :116-119
                    if (option->printDetail) std::cout << "Updated anti-diagonal limit on No. " << begin + w << '\n';
                    pw.fLen = std::min(static_cast<int32_t>(curF * 1.2) << 1, minLen);
                    pw.xdrop = curX;
                } else if (err == 3) {                                      // :120-123
                    std::cout << "There might be some bugs in the code!\n";
                    std::exit(1);
                } else {                                                    // :124-129
                    pw.xdrop = static_cast<int32_t>(curX * 2);
                    pw.fLen = std::min(static_cast<int32_t>(pw.xdrop * 4) << 1, minLen);
                    if (option->printDetail) std::cout << "Updated x-drop value on No. " << begin + w << "\tNew Xdrop: " << pw.xdrop << '\n';
                }
                again.push_back(w);
            }
            pending.swap(again);
        }

        // ---- host update, alignment-cpu.cpp:135-175 ----------------------------------------------------------
        tbb::parallel_for(tbb::blocked_range<int>(0, count), [&](tbb::blocked_range<int> range) {
        for (int w = range.begin(); w < range.end(); ++w) {
            const int nIdx = static_cast<int>(begin) + w;
            PairWork &pw = work[w];
            pw.release();
            if (task == 0 && (pw.refNum == 1 || pw.qryNum == 1) && pw.lowQ) {
                pw.aln.clear();
                tbb::spin_rw_mutex::scoped_lock lock(fallbackMutex);
                fallbackPairs.push_back(nIdx);
            }
            if (pw.aln.empty()) continue;
            alnPath aln_w_gc;
            int alnRef = 0, alnQry = 0;
            for (auto a : pw.aln) {
                if (a == 0) { alnRef += 1; alnQry += 1; }
                if (a == 1) { alnQry += 1; }
                if (a == 2) { alnRef += 1; }
            }
            alignment_helper::addGappyColumnsBack(pw.aln, aln_w_gc, pw.gappyColumns, param, {alnRef, alnQry}, pw.consensus);
            alnRef = 0, alnQry = 0;
            for (auto a : aln_w_gc) {
                if (a == 0) { alnRef += 1; alnQry += 1; }
                if (a == 1) { alnQry += 1; }
                if (a == 2) { alnRef += 1; }
            }
            const float refWeight = nodes[nIdx].first->alnWeight, qryWeight = nodes[nIdx].second->alnWeight;
            if (alnRef != pw.refLen) std::cout << "R: Post " << nodes[nIdx].first->identifier << "(" << alnRef << "/" << nodes[nIdx].first->getAlnLen(task) << ")\n";
            if (alnQry != pw.qryLen) std::cout << "Q: Post " << nodes[nIdx].second->identifier << "(" << alnQry << "/" << nodes[nIdx].second->getAlnLen(task) << ")\n";
            if (option->alnMode != PLACE_WO_TREE) {
                alignment_helper::updateFrequency(nodes[nIdx], database, aln_w_gc, {refWeight, qryWeight});
                alignment_helper::updateAlignment(nodes[nIdx], database, option, aln_w_gc);
            } else {
                tbb::spin_rw_mutex::scoped_lock lock(database->mapMutex);
                database->subtreeAln[nodes[nIdx].second->seqsIncluded[0]] = aln_w_gc;
            }
        }
        });
        begin = end;
    }
    if (fallbackPairs.empty()) return;
    alignment_helper::fallback2cpu(fallbackPairs, nodes, database, option);
}

} // namespace b200

// Build-time hook used by the drop-in CLI (twilight_b200/host/Makefile): the unchanged twilight-main.cpp and
// progressive.cpp are compiled with -DalignmentKernel_CPU=alignmentKernel_B200_entry, so every place where the
// reference passes cpu::alignmentKernel_CPU resolves to this symbol instead.
namespace cpu {
void alignmentKernel_B200_entry(Tree *T, NodePairVec &alnPairs, SequenceDB *database, Option *option, Params &param) {
    b200::alignmentKernel_B200(T, alnPairs, database, option, param);
}
} // namespace cpu

} // namespace progressive
} // namespace msa

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

void alignmentKernel_B200(Tree *tree, NodePairVec &nodes, SequenceDB *database, Option *option, Params &param);

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

// ---------------------------------------------------------------------------------------------------------------
// Level pipeline: the whole per-pair body of parallelAlignmentCPU on the device through twl_align_level. Rows live in
// HBM (twl_rows_*); the host SequenceDB is kept in sync after every level (rows are copied back into
// SequenceInfo::alnStorage exactly where updateAlignment would have written them), so every other part of the host
// (final write-out, storeSubtreeProfile, --check, parked-sequence materialisation) keeps working unchanged.
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Resident {            // what the device holds for a row id, to detect rows the host replaced (new subtree, new DB)
    const void *owner = nullptr;
    int len = -1;
    bool storage = false;
    uint64_t sig = 0;        // content signature of the host copy the device row corresponds to
};

// 64-bit multiplicative hash over the row bytes (one pass, 8 bytes per step)
uint64_t rowSignature(const char *p, int len) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ static_cast<uint64_t>(len);
    int k = 0;
    for (; k + 8 <= len; k += 8) {
        uint64_t w;
        std::memcpy(&w, p + k, 8);
        h = (h ^ w) * 0xFF51AFD7ED558CCDull;
        h ^= h >> 29;
    }
    uint64_t tail = 0;
    if (k < len) std::memcpy(&tail, p + k, len - k);
    h = (h ^ tail) * 0xC4CEB9FE1A85EC53ull;
    return h ^ (h >> 32);
}
std::vector<Resident> &resident() {
    static std::vector<Resident> r;
    return r;
}

void flatten(const Profile &f, std::vector<float> &out) {
    out.clear();
    for (const auto &col : f) out.insert(out.end(), col.begin(), col.end());
}

void unflatten(const std::vector<float> &in, int P, Profile &f) {
    const size_t n = in.size() / P;
    f.assign(n, std::vector<float>(P, 0.f));
    for (size_t t = 0; t < n; ++t)
        for (int v = 0; v < P; ++v) f[t][v] = in[t * P + v];
}

bool fetchFreq(twl_ctx *ctx, int pair, int what, std::vector<float> &out) {
    size_t bytes = 0;
    if (twl_level_fetch(ctx, pair, what, nullptr, 0, &bytes) != TWL_OK) return false;
    out.resize(bytes / sizeof(float));
    return twl_level_fetch(ctx, pair, what, out.data(), bytes, &bytes) == TWL_OK;
}

} // namespace

void alignmentKernel_B200_level(Tree *tree, NodePairVec &nodes, SequenceDB *database, Option *option, Params &param) {
    ensureContext(param);
    twl_ctx *ctx = device().ctx;
    const int P = param.matrixSize + 1;
    const int task = database->currentTask;
    const int nPairs = static_cast<int>(nodes.size());

    // levels with empty nodes take the profile-batch path (alignment-cpu.cpp:89-90 special case)
    for (auto &pr : nodes)
        if (pr.first->getAlnLen(task) == 0 || pr.second->getAlnLen(task) == 0) { alignmentKernel_B200(tree, nodes, database, option, param); return; }

    std::vector<twl_level_pair> lp(nPairs);
    std::vector<std::vector<int32_t>> ids(2 * nPairs);
    std::vector<std::vector<float>> freqs(2 * nPairs);
    std::vector<char> lowQ(nPairs, 0);
    const bool updateRows = (option->alnMode != PLACE_WO_TREE) && (task != 2);

    // rows that must be (re)sent: unknown to the device or replaced on the host since
    std::vector<int32_t> upIds, upLens;
    std::vector<const char *> upRows;
    std::vector<float> upW;
    auto &res = resident();
    for (int n = 0; n < nPairs; ++n) {
        Node *nd[2] = {nodes[n].first, nodes[n].second};
        twl_node_side *side[2] = {&lp[n].ref, &lp[n].qry};
        for (int s = 0; s < 2; ++s) {
            auto &v = ids[2 * n + s];
            const bool cached = !nd[s]->msaFreq.empty();
            for (int sIdx : nd[s]->seqsIncluded) {
                if (sIdx < 0) continue;                                    // parked group / subtree id: path composition only
                if (cached && !updateRows) continue;                       // rows neither read nor written
                v.push_back(sIdx);
                auto *seq = database->sequences[sIdx];
                if (static_cast<size_t>(sIdx) >= res.size()) res.resize(sIdx + 1);
                Resident &r = res[sIdx];
                const uint64_t sig = rowSignature(seq->alnStorage[seq->storage], seq->len);
                if (r.owner != seq || r.len != seq->len || r.storage != seq->storage || r.sig != sig) {
                    upIds.push_back(sIdx); upLens.push_back(seq->len); upRows.push_back(seq->alnStorage[seq->storage]); upW.push_back(seq->weight);
                    r.owner = seq; r.len = seq->len; r.storage = seq->storage; r.sig = sig;
                }
            }
            if (cached) flatten(nd[s]->msaFreq, freqs[2 * n + s]);
            side[s]->seq_ids = v.data();
            side[s]->n_ids = static_cast<int32_t>(v.size());
            side[s]->aln_len = nd[s]->getAlnLen(task);
            side[s]->aln_num = nd[s]->getAlnNum(task);
            side[s]->aln_weight = nd[s]->alnWeight;
            side[s]->msa_freq = cached ? freqs[2 * n + s].data() : nullptr;
        }
        const int refNum = lp[n].ref.aln_num, qryNum = lp[n].qry.aln_num;
        const bool lowQ_r = (option->alnMode == MERGE_MSA) ? false : ((refNum > 1) ? false : database->sequences[nodes[n].first->seqsIncluded[0]]->lowQuality);
        const bool lowQ_q = (option->alnMode == MERGE_MSA) ? false : ((qryNum > 1) ? false : database->sequences[nodes[n].second->seqsIncluded[0]]->lowQuality);
        lowQ[n] = lowQ_r || lowQ_q;
        lp[n].flags = (lowQ[n] ? TWL_PAIR_PROFILE_ONLY : 0) | (updateRows ? 0 : TWL_PAIR_NO_ROW_UPDATE);
        lp[n].reserved = 0;
    }
    if (!upIds.empty() && twl_rows_upload(ctx, static_cast<int>(upIds.size()), upIds.data(), upRows.data(), upLens.data(), upW.data()) != TWL_OK)
        die("twl_rows_upload", twl_last_error(ctx));

    std::vector<twl_level_result> out(nPairs);
    std::vector<std::vector<int8_t>> pathBuf(nPairs);
    std::vector<int8_t *> pathPtr(nPairs);
    for (int n = 0; n < nPairs; ++n) {
        pathBuf[n].resize(static_cast<size_t>(lp[n].ref.aln_len) + lp[n].qry.aln_len + 1);
        pathPtr[n] = pathBuf[n].data();
    }
    if (twl_align_level(ctx, lp.data(), nPairs, task, option->gappyVertical, alignment_helper::_CAL_PROFILE_TH, pathPtr.data(), out.data()) != TWL_OK)
        die("twl_align_level", twl_last_error(ctx));

    std::vector<int> fallbackPairs;
    std::vector<int32_t> downIds;
    std::vector<char *> downDst;
    for (int n = 0; n < nPairs; ++n) {
        Node *first = nodes[n].first, *second = nodes[n].second;
        std::vector<float> flat;
        // calculateProfile's msaFreq cache (helper.cpp:35-40) happens whether or not the pair aligns
        if ((out[n].cached & 1) && fetchFreq(ctx, n, TWL_F_FREQ_REF, flat)) unflatten(flat, P, first->msaFreq);
        if ((out[n].cached & 2) && fetchFreq(ctx, n, TWL_F_FREQ_QRY, flat)) unflatten(flat, P, second->msaFreq);
        const int refNum = lp[n].ref.aln_num, qryNum = lp[n].qry.aln_num;
        if (!lowQ[n] && out[n].status != 0) {
            if (out[n].status == 3) { std::cout << "There might be some bugs in the code!\n"; std::exit(1); }
            fallbackPairs.push_back(n);                                    // only task 0 leaves a status behind (alignment-cpu.cpp:108-115)
            continue;
        }
        if (lowQ[n]) {
            if (task == 0 && (refNum == 1 || qryNum == 1)) fallbackPairs.push_back(n);   // :135-144
            continue;
        }
        alnPath aln(pathBuf[n].begin(), pathBuf[n].begin() + out[n].path_len);
        if (option->alnMode == PLACE_WO_TREE) {
            database->subtreeAln[second->seqsIncluded[0]] = aln;          // :172-174
            continue;
        }
        // updateFrequency, helper.cpp:506-539
        if (!first->msaFreq.empty() && !second->msaFreq.empty()) {
            if (!(out[n].cached & 4) || !fetchFreq(ctx, n, TWL_F_FREQ_MERGED, flat)) die("twl_level_fetch", "merged msaFreq missing");
            second->msaFreq.clear();
            unflatten(flat, P, first->msaFreq);
            first->alnLen = static_cast<int>(first->msaFreq.size());
        }
        // updateAlignment, helper.cpp:377-503: resident rows were rewritten on the device; mirror them into the host DB
        const int totalLen = static_cast<int>(aln.size());
        Node *nd[2] = {first, second};
        for (int s = 0; s < 2; ++s) {
            const int8_t own = (s == 0) ? 2 : 1;
            for (int sIdx : nd[s]->seqsIncluded) {
                if (task != 2 && sIdx >= 0) {
                    auto *seq = database->sequences[sIdx];
                    seq->memCheck(totalLen);
                    downIds.push_back(sIdx);
                    downDst.push_back(seq->alnStorage[1 - seq->storage]);
                    seq->len = totalLen;
                    seq->changeStorage();
                    Resident &r = res[sIdx];
                    r.owner = seq; r.len = totalLen; r.storage = seq->storage;
                } else {
                    alnPath &org = database->subtreeAln[sIdx];
                    alnPath updated(aln.size());
                    int orgIdx = 0;
                    for (size_t k = 0; k < aln.size(); ++k) updated[k] = (aln[k] == 0 || aln[k] == own) ? org[orgIdx++] : static_cast<int8_t>(1);
                    org = updated;
                }
            }
        }
        first->alnNum += second->alnNum;
        first->alnLen = totalLen;
        first->alnWeight += second->alnWeight;
        for (auto idx : second->seqsIncluded) first->seqsIncluded.push_back(idx);
        second->seqsIncluded.clear();
        // parking of >1000 sequences behind one group id, helper.cpp:479-500
        if (first->seqsIncluded.size() > alignment_helper::_UPDATE_SEQ_TH && !first->msaFreq.empty() && task != 2) {
            int seqCount = 0, firstSeqID = 0;
            for (auto idx : first->seqsIncluded)
                if (idx > 1) { if (firstSeqID == 0) firstSeqID = -idx; seqCount++; }
            if (seqCount >= alignment_helper::_UPDATE_SEQ_TH) {
                database->subtreeAln[firstSeqID] = alnPath(totalLen, 0);
                std::vector<int> kept;
                kept.push_back(firstSeqID);
                for (auto idx : first->seqsIncluded) {
                    if (idx >= 0) database->sequences[idx]->subtreeIdx = firstSeqID;
                    else kept.push_back(idx);
                }
                first->seqsIncluded = kept;
            }
        }
    }
    if (!downIds.empty() && twl_rows_download(ctx, static_cast<int>(downIds.size()), downIds.data(), downDst.data(), nullptr) != TWL_OK)
        die("twl_rows_download", twl_last_error(ctx));
    for (int32_t id : downIds) {
        auto *seq = database->sequences[id];
        res[id].sig = rowSignature(seq->alnStorage[seq->storage], seq->len);
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
    // TWL_PIPELINE=dp keeps profile preparation and row update on the host (reference helpers) and runs only the DP on
    // the device; the default runs the whole per-pair pipeline on the device.
    static const bool dpOnly = [] { const char *e = std::getenv("TWL_PIPELINE"); return e && std::string(e) == "dp"; }();
    if (dpOnly) b200::alignmentKernel_B200(T, alnPairs, database, option, param);
    else b200::alignmentKernel_B200_level(T, alnPairs, database, option, param);
}
} // namespace cpu

} // namespace progressive
} // namespace msa

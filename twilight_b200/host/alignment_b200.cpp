// alignment_b200.cpp — the reference-side binding: an `msa::alnFunction` (src/msa.hpp:175) that runs a guide-tree
// level through libtwilight_b200.so. It is compiled together with the UNCHANGED TWILIGHT host sources (option parsing,
// tree, partitioning, sequence DB, I/O, progressive scheduler) and passed to msaOnSubtree() exactly where the
// reference passes cpu::alignmentKernel_CPU (src/twilight-main.cpp:148,183,201,220,261,299; src/progressive.cpp:291).
//
// Division of labour: the whole per-pair body of parallelAlignmentCPU (src/alignment-cpu.cpp:46-176: calculateProfile,
// getConsensus, removeGappyColumns, calculatePSGP, Talco_xdrop::Align_freq, addGappyColumnsBack, updateFrequency and the
// row rewrite of updateAlignment) runs on the B200 in ONE call per level, twl_align_level. The member rows live in HBM
// for the whole progressive alignment of a subtree: a row is uploaded once, when its leaf is first aligned, rewritten
// on the device at every level, and copied back to SequenceInfo::alnStorage once — when its node is parked behind a
// group id (helper.cpp:479-500; from then on only its path is composed) or when msaOnSubtree ends. Per level the host
// does O(pairs + member ids) bookkeeping: node fields, deferral (fallback2cpu), parking, path composition for
// negative ids. No alignment arithmetic runs on the host, and there is no CPU fallback: if the CUDA library cannot
// initialise, the process aborts with a message.
//
// Hooks: `alignmentKernel_B200_entry` is the alnFunction; `msa::progressive::msaOnSubtree` is defined
// HERE as a three-line wrapper (new row generation -> the reference's msaOnSubtree -> rows back to the host) around
// the reference's own function, which twilight_b200/host/Makefile compiles under the name msaOnSubtree_stock.
#ifndef MSA_HPP
#include "msa.hpp"
#endif
#include "twilight_b200.h"

#include <tbb/parallel_for.h>
#include <tbb/spin_rw_mutex.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

namespace msa {
namespace progressive {

// the reference's msaOnSubtree (src/progressive.cpp:232), compiled with -DmsaOnSubtree=msaOnSubtree_stock
void msaOnSubtree_stock(Tree *T, SequenceDB *database, Option *option, Params &param, alnFunction alignmentKernel, int subtree);

namespace b200 {

namespace {

struct Device {
    twl_ctx *ctx = nullptr;
    int M = 0;
    float gapOpen = 0, gapExtend = 0, gapBoundary = 0;
    std::vector<float> score;
};

// TWL_STATS=1: one JSON line on stderr at exit with what the device did for this run (bench.py reads it for the
// BASELINE.json configs that go through the CLI): device milliseconds by phase, DP cell updates, pairs, levels,
// bytes moved each way, and the wall time spent inside the level calls (host bookkeeping included).
struct Stats {
    bool on = false;
    double phaseMs[4] = {0, 0, 0, 0}, wallMs = 0;
    unsigned long long cells = 0, pairs = 0, levels = 0, h2dBytes = 0, d2hBytes = 0, deferred = 0;
};
Stats &stats() {
    static Stats s;
    return s;
}
void printStats() {
    const Stats &s = stats();
    std::fprintf(stderr, "[twl-stats] {\"levels\": %llu, \"pairs\": %llu, \"deferred_pairs\": %llu, \"cells\": %llu, \"device_ms\": %.3f, "
                 "\"phase_ms\": {\"profile_build\": %.3f, \"gappy_psgp_pack\": %.3f, \"dp_chain\": %.3f, \"row_update_freq_merge\": %.3f}, "
                 "\"level_calls_wall_ms\": %.3f, \"h2d_row_bytes\": %llu, \"d2h_row_bytes\": %llu}\n",
                 s.levels, s.pairs, s.deferred, s.cells, s.phaseMs[0] + s.phaseMs[1] + s.phaseMs[2] + s.phaseMs[3], s.phaseMs[0], s.phaseMs[1],
                 s.phaseMs[2], s.phaseMs[3], s.wallMs, s.h2dBytes, s.d2hBytes);
}

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
        stats().on = std::getenv("TWL_STATS") != nullptr;
        std::atexit([] {
            if (stats().on) printStats();
            if (device().ctx) { twl_destroy(device().ctx); device().ctx = nullptr; }
        });
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

// Where the current text of a row lives. Within one msaOnSubtree call the host never writes a member row between
// levels (the only host writer, progressive::updateAlignment, runs after the last level and only touches parked rows),
// so residency is tracked by generation: beginSubtree() forgets everything, a row is uploaded the first time a level
// needs it, and `dirty` means the device holds a newer text than SequenceInfo::alnStorage.
struct Resident {
    bool onDevice = false, dirty = false;
};
struct RowState {
    SequenceDB *db = nullptr;
    std::vector<Resident> rows;
    size_t nDirty = 0;
};
RowState &rowState() {
    static RowState r;
    return r;
}

// device rows -> SequenceInfo::alnStorage[storage] for the given ids (all of them dirty), in bounded slices
void downloadRows(SequenceDB *database, const std::vector<int32_t> &ids) {
    if (ids.empty()) return;
    twl_ctx *ctx = device().ctx;
    RowState &rs = rowState();
    constexpr size_t kSliceBytes = static_cast<size_t>(512) << 20;
    size_t at = 0;
    while (at < ids.size()) {
        size_t end = at, bytes = 0;
        std::vector<char *> dst;
        while (end < ids.size() && (end == at || bytes < kSliceBytes)) {
            auto *seq = database->sequences[ids[end]];
            seq->memCheck(seq->len);                                       // sequencedb.cpp:57-76 (host len already tracks the device)
            dst.push_back(seq->alnStorage[seq->storage]);
            bytes += static_cast<size_t>(seq->len);
            ++end;
        }
        stats().d2hBytes += bytes;
        if (twl_rows_download(ctx, static_cast<int>(end - at), ids.data() + at, dst.data(), nullptr) != TWL_OK)
            die("twl_rows_download", twl_last_error(ctx));
        at = end;
    }
    for (int32_t id : ids)
        if (rs.rows[id].dirty) { rs.rows[id].dirty = false; --rs.nDirty; }
}

void downloadAllDirty(SequenceDB *database) {
    RowState &rs = rowState();
    if (!rs.nDirty || rs.db != database) return;
    std::vector<int32_t> ids;
    for (size_t id = 0; id < rs.rows.size() && id < database->sequences.size(); ++id)
        if (rs.rows[id].dirty) ids.push_back(static_cast<int32_t>(id));
    downloadRows(database, ids);
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

// A new SequenceDB generation: the device row store is emptied (its pools are kept for reuse) and nothing is assumed
// resident. Called by the msaOnSubtree wrapper below, i.e. once per subtree / merge pass.
void beginSubtree(SequenceDB *database) {
    RowState &rs = rowState();
    if (device().ctx && twl_rows_clear(device().ctx) != TWL_OK) die("twl_rows_clear", twl_last_error(device().ctx));
    rs.db = database;
    rs.rows.clear();
    rs.nDirty = 0;
}

// Every row the device rewrote and the host has not seen yet goes back into the SequenceDB, so that everything after
// msaOnSubtree (storeSubtreeProfile, writeSubAlignments, writeFinalMSA, --check) reads the same bytes as after the CPU path.
void endSubtree(SequenceDB *database) { downloadAllDirty(database); }

void alignmentKernel_B200_level(Tree *tree, NodePairVec &nodes, SequenceDB *database, Option *option, Params &param) {
    ensureContext(param);
    twl_ctx *ctx = device().ctx;
    RowState &rs = rowState();
    const auto wall0 = std::chrono::steady_clock::now();
    if (rs.db != database) beginSubtree(database);                         // called outside the msaOnSubtree wrapper
    const int P = param.matrixSize + 1;
    const int task = database->currentTask;
    const int nPairs = static_cast<int>(nodes.size());

    std::vector<twl_level_pair> lp(nPairs);
    std::vector<std::vector<int32_t>> ids(2 * nPairs);
    std::vector<std::vector<float>> freqs(2 * nPairs);
    std::vector<char> lowQ(nPairs, 0), needPath(nPairs, 0);
    const bool updateRows = (option->alnMode != PLACE_WO_TREE) && (task != 2);

    // rows the device does not hold yet
    std::vector<int32_t> upIds, upLens;
    std::vector<const char *> upRows;
    std::vector<float> upW;
    for (int n = 0; n < nPairs; ++n) {
        Node *nd[2] = {nodes[n].first, nodes[n].second};
        twl_node_side *side[2] = {&lp[n].ref, &lp[n].qry};
        for (int s = 0; s < 2; ++s) {
            auto &v = ids[2 * n + s];
            const bool cached = !nd[s]->msaFreq.empty();
            for (int sIdx : nd[s]->seqsIncluded) {
                if (sIdx < 0) { needPath[n] = 1; continue; }               // parked group / subtree id: path composition only
                if (cached && !updateRows) continue;                       // rows neither read nor written
                v.push_back(sIdx);
                if (static_cast<size_t>(sIdx) >= rs.rows.size()) rs.rows.resize(std::max<size_t>(sIdx + 1, database->sequences.size()));
                Resident &r = rs.rows[sIdx];
                if (!r.onDevice) {
                    auto *seq = database->sequences[sIdx];
                    upIds.push_back(sIdx); upLens.push_back(seq->len); upRows.push_back(seq->alnStorage[seq->storage]); upW.push_back(seq->weight);
                    r.onDevice = true;
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
        if (option->alnMode == PLACE_WO_TREE || task == 2) needPath[n] = 1;
    }

    // the final path comes back to the host only where the host composes paths with it (negative ids, PLACE_WO_TREE, merges)
    std::vector<twl_level_result> out(nPairs);
    std::vector<std::vector<int8_t>> pathBuf(nPairs);
    std::vector<int8_t *> pathPtr(nPairs, nullptr);
    for (int n = 0; n < nPairs; ++n)
        if (needPath[n]) {
            pathBuf[n].resize(static_cast<size_t>(lp[n].ref.aln_len) + lp[n].qry.aln_len + 1);
            pathPtr[n] = pathBuf[n].data();
        }

    // One retry when the device runs out of memory: everything the device holds goes back to the host, the row store is
    // emptied and the level's rows are sent again (inputs larger than HBM degrade to per-level staging instead of failing).
    for (int attempt = 0;; ++attempt) {
        int rc = TWL_OK;
        if (!upIds.empty()) rc = twl_rows_upload(ctx, static_cast<int>(upIds.size()), upIds.data(), upRows.data(), upLens.data(), upW.data());
        if (rc == TWL_OK) rc = twl_align_level(ctx, lp.data(), nPairs, task, option->gappyVertical, alignment_helper::_CAL_PROFILE_TH, pathPtr.data(), out.data());
        if (rc == TWL_OK) break;
        if (rc != TWL_E_NOMEM || attempt > 0) die("twl_align_level", twl_last_error(ctx));
        std::cerr << "twilight-b200: device memory exhausted, spilling the row store to the host and retrying the level\n";
        downloadAllDirty(database);
        beginSubtree(database);
        upIds.clear(); upLens.clear(); upRows.clear(); upW.clear();
        for (auto &v : ids)
            for (int32_t sIdx : v) {
                if (static_cast<size_t>(sIdx) >= rs.rows.size()) rs.rows.resize(std::max<size_t>(sIdx + 1, database->sequences.size()));
                if (rs.rows[sIdx].onDevice) continue;
                auto *seq = database->sequences[sIdx];
                upIds.push_back(sIdx); upLens.push_back(seq->len); upRows.push_back(seq->alnStorage[seq->storage]); upW.push_back(seq->weight);
                rs.rows[sIdx].onDevice = true;
            }
    }

    if (stats().on) {
        Stats &st = stats();
        float ph[4] = {0, 0, 0, 0};
        twl_level_phase_ms(ctx, ph);
        for (int i = 0; i < 4; ++i) st.phaseMs[i] += ph[i];
        st.levels += 1;
        st.pairs += static_cast<unsigned long long>(nPairs);
        for (int n = 0; n < nPairs; ++n) st.cells += out[n].cells;
        for (int32_t len : upLens) st.h2dBytes += static_cast<unsigned long long>(len);
    }
    std::vector<int> fallbackPairs;
    std::vector<int32_t> parkedIds;
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
        const int totalLen = out[n].path_len;
        if (option->alnMode == PLACE_WO_TREE) {
            database->subtreeAln[second->seqsIncluded[0]] = alnPath(pathBuf[n].begin(), pathBuf[n].begin() + totalLen);   // :172-174
            continue;
        }
        // updateFrequency, helper.cpp:506-539
        if (!first->msaFreq.empty() && !second->msaFreq.empty()) {
            if (!(out[n].cached & 4) || !fetchFreq(ctx, n, TWL_F_FREQ_MERGED, flat)) die("twl_level_fetch", "merged msaFreq missing");
            second->msaFreq.clear();
            unflatten(flat, P, first->msaFreq);
            first->alnLen = static_cast<int>(first->msaFreq.size());
        }
        // updateAlignment, helper.cpp:377-503: resident rows were rewritten on the device and stay there; the host keeps the
        // length. Parked groups / subtree ids (negative) have their stored path composed with this pair's path.
        Node *nd[2] = {first, second};
        for (int s = 0; s < 2; ++s) {
            const int8_t own = (s == 0) ? 2 : 1;
            for (int sIdx : nd[s]->seqsIncluded) {
                if (task != 2 && sIdx >= 0) {
                    database->sequences[sIdx]->len = totalLen;
                    Resident &r = rs.rows[sIdx];
                    if (!r.dirty) { r.dirty = true; ++rs.nDirty; }
                } else {
                    const int8_t *aln = pathBuf[n].data();
                    alnPath &org = database->subtreeAln[sIdx];
                    alnPath updated(totalLen);
                    int orgIdx = 0;
                    for (int k = 0; k < totalLen; ++k) updated[k] = (aln[k] == 0 || aln[k] == own) ? org[orgIdx++] : static_cast<int8_t>(1);
                    org.swap(updated);
                }
            }
        }
        first->alnNum += second->alnNum;
        first->alnLen = totalLen;
        first->alnWeight += second->alnWeight;
        for (auto idx : second->seqsIncluded) first->seqsIncluded.push_back(idx);
        second->seqsIncluded.clear();
        // parking of >1000 sequences behind one group id, helper.cpp:479-500: from now on only the group's path is composed,
        // the rows themselves are final until progressive::updateAlignment expands them -> they go back to the host now
        if (first->seqsIncluded.size() > alignment_helper::_UPDATE_SEQ_TH && !first->msaFreq.empty() && task != 2) {
            int seqCount = 0, firstSeqID = 0;
            for (auto idx : first->seqsIncluded)
                if (idx > 1) { if (firstSeqID == 0) firstSeqID = -idx; seqCount++; }
            if (seqCount >= alignment_helper::_UPDATE_SEQ_TH) {
                database->subtreeAln[firstSeqID] = alnPath(totalLen, 0);
                std::vector<int> kept;
                kept.push_back(firstSeqID);
                for (auto idx : first->seqsIncluded) {
                    if (idx >= 0) {
                        database->sequences[idx]->subtreeIdx = firstSeqID;
                        if (rs.rows[idx].dirty) parkedIds.push_back(idx);
                    } else kept.push_back(idx);
                }
                first->seqsIncluded = kept;
            }
        }
    }
    downloadRows(database, parkedIds);
    if (!fallbackPairs.empty()) alignment_helper::fallback2cpu(fallbackPairs, nodes, database, option);
    if (stats().on) {
        stats().deferred += fallbackPairs.size();
        stats().wallMs += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
    }
}

} // namespace b200

// The subtree driver the CLI calls (src/twilight-main.cpp:148,183,201,220,261,299): the reference's own msaOnSubtree
// between a "forget the device rows" and a "rows back to the host".
void msaOnSubtree(Tree *T, SequenceDB *database, Option *option, Params &param, alnFunction alignmentKernel, int subtree) {
    b200::beginSubtree(database);
    msaOnSubtree_stock(T, database, option, param, alignmentKernel, subtree);
    b200::endSubtree(database);
}

// Build-time hook used by the drop-in CLI (twilight_b200/host/Makefile): the unchanged twilight-main.cpp and
// progressive.cpp are compiled with -DalignmentKernel_CPU=alignmentKernel_B200_entry, so every place where the
// reference passes cpu::alignmentKernel_CPU resolves to this symbol instead.
namespace cpu {
void alignmentKernel_B200_entry(Tree *T, NodePairVec &alnPairs, SequenceDB *database, Option *option, Params &param) {
    b200::alignmentKernel_B200_level(T, alnPairs, database, option, param);
}
} // namespace cpu

} // namespace progressive
} // namespace msa

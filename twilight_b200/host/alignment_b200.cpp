// alignment_b200.cpp — the reference-side binding: an `msa::alnFunction` (src/msa.hpp:175) that runs a guide-tree
// level through libtwilight_b200.so. It is compiled together with the UNCHANGED TWILIGHT host sources (option parsing,
// tree, partitioning, sequence DB, I/O, progressive scheduler) and passed to msaOnSubtree() exactly where the
// reference passes cpu::alignmentKernel_CPU (src/twilight-main.cpp:148,183,201,220,261,299; src/progressive.cpp:291).
//
// Division of labour: the whole per-pair body of parallelAlignmentCPU (src/alignment-cpu.cpp:46-176: calculateProfile,
// getConsensus, removeGappyColumns, calculatePSGP, Talco_xdrop::Align_freq, addGappyColumnsBack, updateFrequency and the
// row rewrite of updateAlignment) runs on the B200 in ONE call per level, twl_align_level. The member rows live in HBM
// for the whole progressive alignment of a subtree: a row is uploaded once, when its leaf is first aligned, rewritten
// on the device at every level — including the rows of nodes the reference would "park" behind a group id
// (helper.cpp:479-500), so the final MSA is materialised in HBM — and copied back to SequenceInfo::alnStorage once, when
// msaOnSubtree ends (with TWL_PARK=1: when its node is parked). Per level the host
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
#include <string>
#include <thread>
#include <vector>

namespace msa {
namespace progressive {

// the reference's msaOnSubtree (src/progressive.cpp:232), compiled with -DmsaOnSubtree=msaOnSubtree_stock
void msaOnSubtree_stock(Tree *T, SequenceDB *database, Option *option, Params &param, alnFunction alignmentKernel, int subtree);

namespace b200 {

namespace {

// One context per GPU the run may use: TWL_DEVICES="0,1,2,3" (or "all"); default one device, TWL_DEVICE or 0. This is the
// reference GPU build's --gpu-index / one-host-thread-per-GPU scheme (src/cuda/gpu-info.cu:6-94, alignment-gpu.cu:226-253)
// with device-resident rows: every row lives on exactly one device at a time.
struct Device {
    twl_ctx *ctx = nullptr;
    int index = 0;
};
struct Devices {
    std::vector<Device> dev;
    int M = 0;
    float gapOpen = 0, gapExtend = 0, gapBoundary = 0;
    std::vector<float> score;
};

// TWL_STATS=1: one JSON line on stderr at exit with what the device did for this run (bench.py reads it for the
// BASELINE.json configs that go through the CLI): device milliseconds by phase (per level the slowest device counts), DP
// cell updates, pairs, levels, bytes moved each way, and the wall time spent inside the level calls (host bookkeeping included).
struct Stats {
    bool on = false;
    double phaseMs[4] = {0, 0, 0, 0}, wallMs = 0;
    unsigned long long cells = 0, pairs = 0, levels = 0, h2dBytes = 0, d2hBytes = 0, deferred = 0, migratedRows = 0;
    std::vector<unsigned long long> pairsOnDevice;
};
Stats &stats() {
    static Stats s;
    return s;
}
void printStats() {
    const Stats &s = stats();
    std::string per = "[";
    for (size_t d = 0; d < s.pairsOnDevice.size(); ++d) per += (d ? ", " : "") + std::to_string(s.pairsOnDevice[d]);
    per += "]";
    std::fprintf(stderr, "[twl-stats] {\"levels\": %llu, \"pairs\": %llu, \"deferred_pairs\": %llu, \"cells\": %llu, \"device_ms\": %.3f, "
                 "\"phase_ms\": {\"profile_build\": %.3f, \"gappy_psgp_pack\": %.3f, \"dp_chain\": %.3f, \"row_update_freq_merge\": %.3f}, "
                 "\"level_calls_wall_ms\": %.3f, \"h2d_row_bytes\": %llu, \"d2h_row_bytes\": %llu, \"devices\": %zu, \"pairs_per_device\": %s, "
                 "\"rows_migrated_between_devices\": %llu}\n",
                 s.levels, s.pairs, s.deferred, s.cells, s.phaseMs[0] + s.phaseMs[1] + s.phaseMs[2] + s.phaseMs[3], s.phaseMs[0], s.phaseMs[1],
                 s.phaseMs[2], s.phaseMs[3], s.wallMs, s.h2dBytes, s.d2hBytes, s.pairsOnDevice.size(), per.c_str(), s.migratedRows);
}

Devices &devices() {
    static Devices d;
    return d;
}

[[noreturn]] void die(const char *what, const char *detail) {
    std::cerr << "twilight-b200: " << what << ": " << detail << "\n";
    std::exit(1);
}

void ensureContexts(Params &param) {
    Devices &D = devices();
    if (D.dev.empty()) {
        std::vector<int> want;
        if (const char *env = std::getenv("TWL_DEVICES")) {
            const std::string all(env);
            if (all == "all") {
                for (int i = 0; i < twl_device_count(); ++i) want.push_back(i);
            } else {
                size_t at = 0;
                while (at < all.size()) {
                    const size_t end = std::min(all.find(',', at), all.size());
                    if (end > at) want.push_back(std::atoi(all.substr(at, end - at).c_str()));
                    at = end + 1;
                }
            }
        }
        if (want.empty()) {
            const char *env = std::getenv("TWL_DEVICE");
            want.push_back(env ? std::atoi(env) : 0);
        }
        for (int idx : want) {
            Device d;
            d.index = idx;
            if (twl_init(idx, &d.ctx) != TWL_OK) die("cannot initialise the CUDA device", twl_last_error(nullptr));
            D.dev.push_back(d);
        }
        stats().on = std::getenv("TWL_STATS") != nullptr;
        stats().pairsOnDevice.assign(D.dev.size(), 0);
        std::atexit([] {
            if (stats().on) printStats();
            for (Device &d : devices().dev) if (d.ctx) { twl_destroy(d.ctx); d.ctx = nullptr; }
        });
    }
    const int M = param.matrixSize;
    std::vector<float> flat(static_cast<size_t>(M) * M);
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) flat[i * M + j] = param.scoringMatrix[i][j];
    if (D.M != M || D.gapOpen != param.gapOpen || D.gapExtend != param.gapExtend || D.gapBoundary != param.gapBoundary || D.score != flat) {
        for (Device &d : D.dev)
            if (twl_set_params(d.ctx, flat.data(), M, param.gapOpen, param.gapExtend, param.gapBoundary) != TWL_OK)
                die("twl_set_params", twl_last_error(d.ctx));
        D.M = M; D.gapOpen = param.gapOpen; D.gapExtend = param.gapExtend; D.gapBoundary = param.gapBoundary; D.score = flat;
    }
}

// Where the current text of a row lives. Within one msaOnSubtree call the host never writes a member row between
// levels (the only host writer, progressive::updateAlignment, runs after the last level and only touches parked rows),
// so residency is tracked by generation: beginSubtree() forgets everything, a row is uploaded the first time a level
// needs it, and `dirty` means a device holds a newer text than SequenceInfo::alnStorage.
struct Resident {
    int8_t dev = -1;         // device slot that holds the row, -1: host only
    int8_t home = -1;        // preferred device slot of a leaf (contiguous blocks of the guide tree's leaf order), -1: none
    bool dirty = false;
};
struct RowState {
    SequenceDB *db = nullptr;
    std::vector<Resident> rows;
    size_t nDirty = 0;
    bool homesReady = false;
};
RowState &rowState() {
    static RowState r;
    return r;
}
Resident &residentOf(int id, SequenceDB *database) {
    RowState &rs = rowState();
    if (static_cast<size_t>(id) >= rs.rows.size()) rs.rows.resize(std::max<size_t>(id + 1, database->sequences.size()));
    return rs.rows[id];
}

// device rows -> SequenceInfo::alnStorage[storage] for the given ids (all resident), per device and in bounded slices
void downloadRows(SequenceDB *database, const std::vector<int32_t> &ids) {
    if (ids.empty()) return;
    RowState &rs = rowState();
    Devices &D = devices();
    constexpr size_t kSliceBytes = static_cast<size_t>(512) << 20;
    // SequenceInfo::memCheck (sequencedb.cpp:57-76) reallocates and zero-fills both buffers of a row that outgrew them: for a
    // final MSA that is several GB of host writes, so it runs on all host threads like the reference's own update loops
    tbb::parallel_for(tbb::blocked_range<size_t>(0, ids.size()), [&](tbb::blocked_range<size_t> range) {
        for (size_t i = range.begin(); i < range.end(); ++i) {
            auto *seq = database->sequences[ids[i]];
            seq->memCheck(seq->len);
        }
    });
    for (size_t d = 0; d < D.dev.size(); ++d) {
        std::vector<int32_t> mine;
        for (int32_t id : ids) if (rs.rows[id].dev == static_cast<int8_t>(d)) mine.push_back(id);
        size_t at = 0;
        while (at < mine.size()) {
            size_t end = at, bytes = 0;
            std::vector<char *> dst;
            while (end < mine.size() && (end == at || bytes < kSliceBytes)) {
                auto *seq = database->sequences[mine[end]];
                dst.push_back(seq->alnStorage[seq->storage]);
                bytes += static_cast<size_t>(seq->len);
                ++end;
            }
            stats().d2hBytes += bytes;
            if (twl_rows_download(D.dev[d].ctx, static_cast<int>(end - at), mine.data() + at, dst.data(), nullptr) != TWL_OK)
                die("twl_rows_download", twl_last_error(D.dev[d].ctx));
            at = end;
        }
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

// Leaves of the guide tree in depth-first order get contiguous blocks of devices as their "home": sibling subtrees then
// meet on one device and only the few joins near the root move rows between GPUs (subtree-affine placement).
void assignHomes(Tree *tree, SequenceDB *database) {
    RowState &rs = rowState();
    rs.homesReady = true;
    const size_t nDev = devices().dev.size();
    if (nDev < 2 || !tree || !tree->root) return;
    std::vector<int> leafIds;
    std::vector<Node *> stack{tree->root};
    while (!stack.empty()) {
        Node *nd = stack.back();
        stack.pop_back();
        if (nd->children.empty()) {
            auto it = database->name_map.find(nd->identifier);
            if (it != database->name_map.end()) leafIds.push_back(it->second->id);
            continue;
        }
        for (auto c = nd->children.rbegin(); c != nd->children.rend(); ++c) stack.push_back(*c);
    }
    for (size_t k = 0; k < leafIds.size(); ++k) residentOf(leafIds[k], database).home = static_cast<int8_t>(k * nDev / leafIds.size());
}

} // namespace

// A new SequenceDB generation: the device row stores are emptied (their pools are kept for reuse) and nothing is assumed
// resident. Called by the msaOnSubtree wrapper below, i.e. once per subtree / merge pass.
void beginSubtree(SequenceDB *database) {
    RowState &rs = rowState();
    for (Device &d : devices().dev)
        if (d.ctx && twl_rows_clear(d.ctx) != TWL_OK) die("twl_rows_clear", twl_last_error(d.ctx));
    rs.db = database;
    rs.rows.clear();
    rs.nDirty = 0;
    rs.homesReady = false;
}

// Every row a device rewrote and the host has not seen yet goes back into the SequenceDB, so that everything after
// msaOnSubtree (storeSubtreeProfile, writeSubAlignments, writeFinalMSA, --check) reads the same bytes as after the CPU path.
void endSubtree(SequenceDB *database) { downloadAllDirty(database); }

void alignmentKernel_B200_level(Tree *tree, NodePairVec &nodes, SequenceDB *database, Option *option, Params &param) {
    ensureContexts(param);
    Devices &D = devices();
    const int nDev = static_cast<int>(D.dev.size());
    RowState &rs = rowState();
    const auto wall0 = std::chrono::steady_clock::now();
    if (rs.db != database) beginSubtree(database);                         // called outside the msaOnSubtree wrapper
    if (!rs.homesReady) assignHomes(tree, database);
    const int P = param.matrixSize + 1;
    const int task = database->currentTask;
    const int nPairs = static_cast<int>(nodes.size());

    std::vector<twl_level_pair> lp(nPairs);
    std::vector<std::vector<int32_t>> ids(2 * nPairs);
    std::vector<std::vector<float>> freqs(2 * nPairs);
    std::vector<char> lowQ(nPairs, 0), needPath(nPairs, 0);
    std::vector<int> devOf(nPairs, 0), localIdx(nPairs, 0);
    const bool updateRows = (option->alnMode != PLACE_WO_TREE) && (task != 2);

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

    // ---- which device aligns which pair: where most of its rows already are; for rows nobody holds yet, the home of the
    // pair's first leaf; otherwise the device with the least work so far in this level
    std::vector<double> load(nDev, 0.0);
    if (nDev > 1) {
        for (int n = 0; n < nPairs; ++n) {
            std::vector<double> bytes(nDev, 0.0);
            int home = -1;
            bool any = false;
            for (int s = 0; s < 2; ++s)
                for (int32_t id : ids[2 * n + s]) {
                    const Resident &r = residentOf(id, database);
                    if (r.dev >= 0) { bytes[r.dev] += database->sequences[id]->len + 1; any = true; }
                    else if (home < 0) home = r.home;
                }
            int d = 0;
            if (any) d = static_cast<int>(std::max_element(bytes.begin(), bytes.end()) - bytes.begin());
            else if (home >= 0) d = home;
            else d = static_cast<int>(std::min_element(load.begin(), load.end()) - load.begin());
            devOf[n] = d;
            load[d] += static_cast<double>(lp[n].ref.aln_len) + lp[n].qry.aln_len;
        }
    }

    // the final path comes back to the host only where the host composes paths with it (negative ids, PLACE_WO_TREE, merges)
    std::vector<twl_level_result> out(nPairs);
    std::vector<std::vector<int8_t>> pathBuf(nPairs);
    for (int n = 0; n < nPairs; ++n)
        if (needPath[n]) pathBuf[n].resize(static_cast<size_t>(lp[n].ref.aln_len) + lp[n].qry.aln_len + 1);

    struct PerDevice {
        std::vector<int> pairs;                       // indices into `nodes`
        std::vector<twl_level_pair> lp;
        std::vector<int8_t *> paths;
        std::vector<twl_level_result> out;
        std::vector<int32_t> upIds, upLens;
        std::vector<const char *> upRows;
        std::vector<float> upW;
        int rc = TWL_OK;
        float phase[4] = {0, 0, 0, 0};
    };
    // One retry when a device runs out of memory: everything the devices hold goes back to the host, the row stores are
    // emptied and the level's rows are sent again (inputs larger than HBM degrade to per-level staging instead of failing).
    for (int attempt = 0;; ++attempt) {
        std::vector<PerDevice> pd(nDev);
        std::vector<std::vector<std::vector<int32_t>>> moves(nDev, std::vector<std::vector<int32_t>>(nDev));   // [from][to] -> row ids
        for (int n = 0; n < nPairs; ++n) {
            PerDevice &p = pd[devOf[n]];
            localIdx[n] = static_cast<int>(p.pairs.size());
            p.pairs.push_back(n);
            p.lp.push_back(lp[n]);
            p.paths.push_back(needPath[n] ? pathBuf[n].data() : nullptr);
            for (int s = 0; s < 2; ++s)
                for (int32_t id : ids[2 * n + s]) {
                    Resident &r = residentOf(id, database);
                    if (r.dev == devOf[n]) continue;
                    if (r.dev >= 0) moves[r.dev][devOf[n]].push_back(id);
                    else {
                        auto *seq = database->sequences[id];
                        p.upIds.push_back(id); p.upLens.push_back(seq->len); p.upRows.push_back(seq->alnStorage[seq->storage]); p.upW.push_back(seq->weight);
                    }
                    r.dev = static_cast<int8_t>(devOf[n]);
                }
        }
        // rows that change device: GPU to GPU (twl_rows_migrate: cudaMemcpyPeer over NVLink), at most a few joins per level
        for (int a = 0; a < nDev; ++a)
            for (int b = 0; b < nDev; ++b)
                if (!moves[a][b].empty()) {
                    if (twl_rows_migrate(D.dev[a].ctx, D.dev[b].ctx, static_cast<int>(moves[a][b].size()), moves[a][b].data()) != TWL_OK)
                        die("twl_rows_migrate", twl_last_error(D.dev[b].ctx));
                    stats().migratedRows += moves[a][b].size();
                }
        auto runDevice = [&](int d) {
            PerDevice &p = pd[d];
            if (p.pairs.empty()) return;
            p.out.resize(p.pairs.size());
            twl_ctx *ctx = D.dev[d].ctx;
            if (!p.upIds.empty()) p.rc = twl_rows_upload(ctx, static_cast<int>(p.upIds.size()), p.upIds.data(), p.upRows.data(), p.upLens.data(), p.upW.data());
            if (p.rc == TWL_OK)
                p.rc = twl_align_level(ctx, p.lp.data(), static_cast<int>(p.lp.size()), task, option->gappyVertical, alignment_helper::_CAL_PROFILE_TH, p.paths.data(), p.out.data());
            if (p.rc == TWL_OK) twl_level_phase_ms(ctx, p.phase);
        };
        // one host thread per GPU, as the reference GPU build (alignment-gpu.cu:247)
        std::vector<std::thread> workers;
        for (int d = 1; d < nDev; ++d) if (!pd[d].pairs.empty()) workers.emplace_back(runDevice, d);
        runDevice(0);
        for (auto &w : workers) w.join();
        bool nomem = false, failed = false;
        for (int d = 0; d < nDev; ++d) {
            if (pd[d].rc == TWL_E_NOMEM) nomem = true;
            else if (pd[d].rc != TWL_OK) { failed = true; std::cerr << "twilight-b200: device " << D.dev[d].index << ": " << twl_last_error(D.dev[d].ctx) << "\n"; }
        }
        // (with several devices a level may have completed on some of them: no clean retry; twl_align_level itself is all or nothing)
        if (failed || (nomem && (attempt > 0 || nDev > 1))) die("twl_align_level", nomem ? "out of device memory" : "failed");
        if (nomem) {
            std::cerr << "twilight-b200: device memory exhausted, spilling the row stores to the host and retrying the level\n";
            // rows of devices whose call did not run keep their content; everything goes back to the host and is forgotten
            downloadAllDirty(database);
            beginSubtree(database);
            rs.homesReady = true;
            continue;
        }
        for (int d = 0; d < nDev; ++d)
            for (size_t k = 0; k < pd[d].pairs.size(); ++k) out[pd[d].pairs[k]] = pd[d].out[k];
        if (stats().on) {
            Stats &st = stats();
            for (int i = 0; i < 4; ++i) {
                float m = 0.f;
                for (int d = 0; d < nDev; ++d) m = std::max(m, pd[d].phase[i]);
                st.phaseMs[i] += m;
            }
            st.levels += 1;
            st.pairs += static_cast<unsigned long long>(nPairs);
            for (int n = 0; n < nPairs; ++n) st.cells += out[n].cells;
            for (int d = 0; d < nDev; ++d) {
                st.pairsOnDevice[d] += pd[d].pairs.size();
                for (int32_t len : pd[d].upLens) st.h2dBytes += static_cast<unsigned long long>(len);
            }
        }
        break;
    }

    std::vector<int> fallbackPairs;
    std::vector<int32_t> parkedIds;
    for (int n = 0; n < nPairs; ++n) {
        Node *first = nodes[n].first, *second = nodes[n].second;
        twl_ctx *ctx = D.dev[devOf[n]].ctx;
        const int ln = localIdx[n];
        std::vector<float> flat;
        // calculateProfile's msaFreq cache (helper.cpp:35-40) happens whether or not the pair aligns
        if ((out[n].cached & 1) && fetchFreq(ctx, ln, TWL_F_FREQ_REF, flat)) unflatten(flat, P, first->msaFreq);
        if ((out[n].cached & 2) && fetchFreq(ctx, ln, TWL_F_FREQ_QRY, flat)) unflatten(flat, P, second->msaFreq);
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
            if (!(out[n].cached & 4) || !fetchFreq(ctx, ln, TWL_F_FREQ_MERGED, flat)) die("twl_level_fetch", "merged msaFreq missing");
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
        // Parking of >1000 sequences behind one group id (helper.cpp:479-500) is a host-memory-bandwidth optimisation of the
        // reference: the rows of a big node stop being rewritten, only one path per group is composed, and
        // progressive::updateAlignment (progressive.cpp:194-230) expands the rows through the composed path at the end. On the
        // device rewriting rows is an HBM-speed kernel, so by default nothing is parked: every member row is rewritten at every
        // level and the final MSA is materialised in HBM when the last level returns (the host's expansion loop finds nothing
        // to do, no path ever crosses PCIe for it). The final bytes are the same — composing paths and applying them one after
        // the other are the same function. TWL_PARK=1 restores the reference's parking (rows of a parked node go back to the
        // host at that moment), which keeps device memory at ~4x the *current* row lengths instead of the final ones.
        static const bool park = [] { const char *e = std::getenv("TWL_PARK"); return e && std::atoi(e) != 0; }();
        if (park && first->seqsIncluded.size() > alignment_helper::_UPDATE_SEQ_TH && !first->msaFreq.empty() && task != 2) {
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

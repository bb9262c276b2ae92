// oracle/twl_oracle.cpp — TEST INFRASTRUCTURE ONLY. See twl_oracle.h for the role and the parity status.
//
// A from-scratch CPU restatement of the reference's per-pair alignment path. Every function cites the reference
// lines it follows. Compile with -ffp-contract=off (oracle/Makefile does): every fused operation of the reference
// build is spelled fmaf() here, everything else must stay un-fused.
#include "twl_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace {

constexpr int32_t kInsBoundary = -2; // I_BOUNDARY, TALCO-XDrop.cpp:33
constexpr int32_t kDelBoundary = -3; // D_BOUNDARY, TALCO-XDrop.cpp:34

// ---------------------------------------------------------------------------------------------------------------
// Column-pair score numerator, TALCO-XDrop.cpp:372-444 in the order the x86 (TALCO_SIMD, contraction on) build
// evaluates it. r and q are profile columns of width P; S is the (P-1)x(P-1) matrix; g = gapCharScore.
// ---------------------------------------------------------------------------------------------------------------
inline float numeratorNt(const float *r, const float *q, const float *S, float g) {
    float num = 0.0f;
    for (int l = 0; l < 5; ++l) { // :378-392 — lanes (q[m]*S[l][m])*r[l], summed left to right
        float t0 = (q[0] * S[l * 5 + 0]) * r[l];
        float t1 = (q[1] * S[l * 5 + 1]) * r[l];
        float t2 = (q[2] * S[l * 5 + 2]) * r[l];
        float t3 = (q[3] * S[l * 5 + 3]) * r[l];
        float t4 = (q[4] * S[l * 5 + 4]) * r[l];
        num = num + ((((t0 + t1) + t2) + t3) + t4);
    }
    for (int l = 0; l < 5; ++l) num = fmaf(r[l] * q[5], g, num); // :393 (contracted)
    for (int m = 0; m < 5; ++m) num = fmaf(r[5] * q[m], g, num); // :394 (contracted)
    return num;
}

inline float numeratorAa(const float *r, const float *q, const float *S, float g) {
    float num = 0.0f;
    for (int l = 0; l < 21; ++l) { // :409-430
        const float *row = S + l * 21;
        float v[8];
        for (int m = 0; m < 8; ++m) v[m] = fmaf(r[l], q[8 + m] * row[8 + m], (q[m] * row[m]) * r[l]);
        for (int m = 16; m < 21; ++m) num = fmaf(r[l] * q[m], row[m], num); // :423-425 (contracted)
        float h = ((((((v[0] + v[1]) + v[2]) + v[3]) + v[4]) + v[5]) + v[6]) + v[7];
        num = h + num;
    }
    // :432-433 — the oracle build vectorises the first 20 terms of each loop (two rounded products, plain adds) and
    // leaves the 21st as a scalar fused multiply-add.
    for (int l = 0; l < 20; ++l) num = num + (r[l] * q[21]) * g;
    num = fmaf(r[20] * q[21], g, num);
    for (int m = 0; m < 20; ++m) num = num + (q[m] * r[21]) * g;
    num = fmaf(g, r[21] * q[20], num);
    return num;
}

struct TalcoInput {
    const twlo_talco_params *prm;
    int P;
    int refTotal, qryTotal;
    const float *freqRef, *freqQry;
    const float *gapOpRef, *gapExRef, *gapOpQry, *gapExQry;
    float refNum, qryNum;
};

struct TileResult {
    std::vector<int8_t> ops; // in the order the reference pushes them (tail runs, then traceback back to front)
    bool lastTile = false;
    int error = 0;
    uint64_t cells = 0, diagonals = 0;
};

// All-equal scan, Talco_xdrop::Reduction_tree (TALCO-XDrop.cpp:110-119): value at `start` if the next `length`
// entries match it, else -1.
inline int32_t allEqualOrMinusOne(const int32_t *c, int32_t start, int32_t length) {
    int32_t v = c[start];
    for (int32_t t = 1; t <= length; ++t)
        if (c[start + t] != v) return -1;
    return v;
}

// Talco_xdrop::Traceback (TALCO-XDrop.cpp:134-231), expressed in (diagonal, row) coordinates: the reference's flat
// address arithmetic (:201-217) lands on (k-2,i-1) for a match step, (k-1,i-1) for an insertion, (k-1,i) for a
// deletion.
void tracebackTile(const std::vector<int32_t> &width, const std::vector<int32_t> &lower, const std::vector<int8_t> &tb,
                   int32_t startDiag, int32_t startRow, int32_t startRefCol, int8_t startState, bool firstTile,
                   std::vector<int8_t> &ops) {
    std::vector<int64_t> base(width.size() + 1, 0);
    for (size_t k = 0; k < width.size(); ++k) base[k + 1] = base[k] + width[k];
    int32_t k = startDiag;
    int16_t row = static_cast<int16_t>(startRow);
    int16_t qi = static_cast<int16_t>(startRow);
    int16_t ri = static_cast<int16_t>(startRefCol);
    int8_t state = startState;
    while (k >= 0) {
        int64_t addr = base[k] + (row - lower[k]);
        int8_t cell = (addr >= 0 && addr < static_cast<int64_t>(tb.size())) ? tb[addr] : 0;
        int8_t dir;
        if (state == 0) {
            int8_t p = cell & 0x03;
            if (p == 0) { dir = 0; state = 0; }
            else if (p == 1) { dir = 1; state = (cell & 0x04) ? 1 : 0; }
            else { dir = 2; state = (cell & 0x08) ? 2 : 0; }
        } else if (state == 1) {
            dir = 1;
            state = (cell & 0x04) ? 1 : 0;
        } else {
            dir = 2;
            state = (cell & 0x08) ? 2 : 0;
        }
        if (dir == 0) { k -= 2; row -= 1; qi--; ri--; }
        else if (dir == 1) { k -= 1; row -= 1; qi--; }
        else { k -= 1; ri--; }
        ops.push_back(dir);
        if (firstTile && (ri < 0 || qi < 0)) break;
    }
    if (firstTile) { // :221-230 leading gap runs of the global alignment
        while (ri > -1) { ops.push_back(2); ri--; }
        while (qi > -1) { ops.push_back(1); qi--; }
    }
}

// Talco_xdrop::Tile (TALCO-XDrop.cpp:233-689).
void runTile(const TalcoInput &in, int32_t &refOff, int32_t &qryOff, int tile, TileResult &out) {
    const twlo_talco_params &p = *in.prm;
    const int P = in.P, M = P - 1;
    const float negInf = -static_cast<float>(2.0 * p.xdrop + 1.0); // :252
    const float xdropF = static_cast<float>(p.xdrop);
    const int32_t marker = p.marker;
    int32_t refLen = in.refTotal - refOff, qryLen = in.qryTotal - qryOff;
    const int32_t cap = std::min(p.fLen, std::min(refLen, qryLen)); // :258
    const float denom = in.refNum * in.qryNum;                       // :269
    const float edgeOpen = p.gapOpen, edgeExtend = p.gapExtend;     // alnType 0, :274-275

    if (refLen < 0 || qryLen < 0) { out.error = 3; return; } // :311-318

    const int pad = std::max(cap, 0) + 4;
    std::vector<float> S[3], I[2], D[2];
    std::vector<int32_t> CS[3], CI[2], CD[2];
    for (int b = 0; b < 3; ++b) { S[b].assign(pad, -1.0f); CS[b].assign(pad, -1); }                 // :301-308
    for (int b = 0; b < 2; ++b) { I[b].assign(pad, -1.0f); D[b].assign(pad, -1.0f); CI[b].assign(pad, kInsBoundary); CD[b].assign(pad, kDelBoundary); }
    int32_t lo[3] = {0, 1, 2}, hi[3] = {0, -1, -2}; // :296-297

    std::vector<int8_t> tb;
    std::vector<int32_t> width, lower;
    float maxScore = 0.0f, maxScorePrime = negInf, convScore = 0.0f;
    bool converged = false, stoppedOnConvergence = false;
    int32_t convValue = 0, prevConvS = -1, lastK = 0;

    for (int32_t k = 0; k < refLen + qryLen - 1; ++k) {
        const int c0 = k % 3, c1 = (k + 2) % 3, c2 = (k + 1) % 3; // this, previous, before-previous diagonal
        const int g0 = k % 2, g1 = (k + 1) % 2;
        const int32_t L0 = lo[c0], U0 = hi[c0], L1 = lo[c1], U1 = hi[c1], L2 = lo[c2], U2 = hi[c2];
        if (L0 >= U0 + 1) { out.lastTile = true; out.error = 1; out.ops.clear(); return; }  // :323-329
        if (U0 - L0 + 1 > cap) { out.lastTile = true; out.error = 2; out.ops.clear(); return; } // :331-338
        if (k <= marker) { width.push_back(U0 - L0 + 1); lower.push_back(L0); }
        out.cells += static_cast<uint64_t>(U0 - L0 + 1);
        out.diagonals += 1;

        for (int32_t i = L0; i <= U0; ++i) {
            const int32_t j = k - i; // :358-359 reduces to this
            const int32_t off = i - L0, offDiag = i - 1 - L2, offUp = i - L1, offLeft = offUp - 1;
            float match = negInf, insOpen = negInf, insExt = negInf, delOpen = negInf, delExt = negInf;
            const bool diagIn = offDiag >= 0 && offDiag <= U2 - L2;
            const bool onEdge0 = (tile == 0) && (i == 0 || j == 0);
            if (k == 0 || diagIn || onEdge0) {
                const float *r = in.freqRef + static_cast<size_t>(refOff + j) * P;
                const float *q = in.freqQry + static_cast<size_t>(qryOff + i) * P;
                const float num = (P == 6) ? numeratorNt(r, q, p.score, p.gapCharScore) : numeratorAa(r, q, p.score, p.gapCharScore);
                const float sim = num / denom;
                if (onEdge0) {
                    if (i == 0 && j == 0) match = sim;
                    else match = fmaf(edgeExtend, static_cast<float>(std::max(0, std::max(refOff + j, qryOff + i) - 1)), sim + edgeOpen); // :448 (contracted)
                } else if (offDiag < 0) match = sim;
                else match = S[c2][offDiag] + sim;
            }
            if (offUp >= 0 && offUp <= U1 - L1) {
                delOpen = S[c1][offUp] + in.gapOpRef[refOff + j];
                delExt = D[g1][offUp] + in.gapExRef[refOff + j];
            }
            if (offLeft >= 0 && offLeft <= U1 - L1) {
                insOpen = S[c1][offLeft] + in.gapOpQry[qryOff + i];
                insExt = I[g1][offLeft] + in.gapExQry[qryOff + i];
            }
            const bool insFromIns = insExt >= insOpen, delFromDel = delExt >= delOpen; // extension wins ties
            const float insBest = insFromIns ? insExt : insOpen, delBest = delFromDel ? delExt : delOpen;
            I[g0][off] = insBest;
            D[g0][off] = delBest;
            int8_t ptr;
            float s;
            if (match >= insBest) {
                if (match >= delBest) { s = match; ptr = 0; }
                else { s = delBest; ptr = 2; }
            } else if (insBest > delBest) { s = insBest; ptr = 1; }
            else { s = delBest; ptr = 2; }
            if (s < maxScore - xdropF) s = negInf; // :495
            S[c0][off] = s;
            if (maxScorePrime < s) maxScorePrime = s;

            if (k == marker - 1) {
                CS[c0][off] = (3 << 16) | (i & 0xFFFF);
            } else if (k == marker) {
                CS[c0][off] = (0 << 16) | (i & 0xFFFF);
                CI[g0][off] = (1 << 16) | (i & 0xFFFF);
                CD[g0][off] = (2 << 16) | (i & 0xFFFF);
            } else if (k >= marker + 1) { // :527-547; upper bounds deliberately unchecked, stale slots are visible
                if (insFromIns) CI[g0][off] = (offLeft >= 0) ? CI[g1][offLeft] : kInsBoundary;
                else CI[g0][off] = (offLeft >= 0 && CS[c1][offLeft] != -1) ? CS[c1][offLeft] : kInsBoundary;
                if (delFromDel) CD[g0][off] = (offUp >= 0) ? CD[g1][offUp] : kDelBoundary;
                else CD[g0][off] = (offUp >= 0 && CS[c1][offUp] != -1) ? CS[c1][offUp] : kDelBoundary;
                if (ptr == 0) CS[c0][off] = (offDiag >= 0) ? CS[c2][offDiag] : -1; // offDiag < 0 is an out-of-bounds read upstream (never hit on real data)
                else if (ptr == 1) CS[c0][off] = CI[g0][off];
                else CS[c0][off] = CD[g0][off];
            }
            if (k <= marker) tb.push_back(static_cast<int8_t>(ptr | (insFromIns ? 0x04 : 0) | (delFromDel ? 0x08 : 0)));
        }

        int32_t newL = L0, newU = U0; // :560-583 trim dead ends
        while (newL <= U0 && S[c0][newL - L0] <= negInf) newL++;
        while (newU >= L0 && S[c0][newU - L0] <= negInf) newU--;

        if (!converged && k < refLen + qryLen - 2) { // :585-595
            const int32_t cI = allEqualOrMinusOne(CI[g0].data(), newL - L0, newU - newL);
            const int32_t cD = allEqualOrMinusOne(CD[g0].data(), newL - L0, newU - newL);
            const int32_t cS = allEqualOrMinusOne(CS[c0].data(), newL - L0, newU - newL);
            if (cI == cD && cI == cS && prevConvS == cS && cI != -1) {
                converged = true;
                convValue = prevConvS;
                convScore = maxScorePrime;
            }
            prevConvS = cS;
        }
        lo[c2] = std::max(newL, std::max(0, k + 2 - refLen)); // diagonal k+1 reuses the slot of k-2
        hi[c2] = std::min(qryLen - 1, newU + 1);
        maxScore = (maxScorePrime < 0) ? 0.0f : maxScorePrime; // :607
        lastK = k;
        if (converged && maxScore > convScore) { stoppedOnConvergence = true; break; }
    }

    // :614-652 locate the cell the traceback starts from
    int32_t convQry, convRef, startDiag;
    int32_t tbState;
    const int32_t nStored = static_cast<int32_t>(width.size());
    if (stoppedOnConvergence || lastK >= marker) {
        const int32_t v = stoppedOnConvergence ? convValue : CS[lastK % 3][0];
        convQry = v & 0xFFFF;
        tbState = static_cast<int8_t>((v >> 16) & 0xFFFF);
        convRef = marker - convQry - ((tbState == 3) ? 1 : 0);
        startDiag = (tbState == 3) ? nStored - 2 : nStored - 1;
    } else {
        convQry = qryLen - 1;
        convRef = refLen - 1;
        startDiag = lastK;
        tbState = 0;
        out.lastTile = true;
    }
    // Boundary codes (:645-652): the tile slid `marker` cells along one profile only. Upstream then starts the
    // traceback from an out-of-range address (undefined); here the start cell is clamped instead.
    if (convQry == (kDelBoundary & 0xFFFF)) { convQry = 0; convRef = marker; }
    else if (convQry == (kInsBoundary & 0xFFFF)) { convQry = marker; convRef = 0; }
    const int32_t tbRow = convQry, tbRefCol = convRef;

    refOff += convRef;
    qryOff += convQry;
    refLen = in.refTotal - refOff;
    qryLen = in.qryTotal - qryOff;
    if (refLen < 0 || qryLen < 0) { out.error = 3; out.ops.clear(); return; }

    if (refOff == in.refTotal - 1 && qryOff < in.qryTotal - 1) { // :671-674
        out.ops.insert(out.ops.end(), in.qryTotal - qryOff - 1, 1);
        out.lastTile = true;
    }
    if (qryOff == in.qryTotal - 1 && refOff < in.refTotal - 1) { // :675-678
        out.ops.insert(out.ops.end(), in.refTotal - refOff - 1, 2);
        out.lastTile = true;
    }
    if (refOff == in.refTotal - 1 && qryOff == in.qryTotal - 1) out.lastTile = true;

    int8_t startState = static_cast<int8_t>(static_cast<int8_t>(tbState) % 3);
    tracebackTile(width, lower, tb, startDiag, static_cast<int16_t>(tbRow), static_cast<int16_t>(tbRefCol), startState, tile == 0, out.ops);
}

} // namespace

extern "C" {

int twlo_talco_align(const twlo_talco_params *p, int refLen, int qryLen, const float *freqRef, const float *freqQry,
                     const float *gapOpRef, const float *gapExRef, const float *gapOpQry, const float *gapExQry,
                     float refNum, float qryNum, int8_t *aln, int *errorType, uint64_t *cells, int32_t *tiles,
                     uint64_t *diagonals) {
    TalcoInput in{p, p->P, refLen, qryLen, freqRef, freqQry, gapOpRef, gapExRef, gapOpQry, gapExQry, refNum, qryNum};
    int32_t refOff = 0, qryOff = 0;
    int tile = 0, n = 0;
    uint64_t nCells = 0, nDiag = 0;
    *errorType = 0;
    bool last = false;
    while (!last) { // Align_freq, TALCO-XDrop.cpp:77-106
        TileResult tr;
        runTile(in, refOff, qryOff, tile, tr);
        nCells += tr.cells;
        nDiag += tr.diagonals;
        last = tr.lastTile;
        if (tr.error) *errorType = tr.error;
        if (tr.ops.empty()) { n = 0; break; }
        const int count = static_cast<int>(tr.ops.size());
        for (int t = count - 1; t >= 0; --t) {
            if (t == count - 1 && tile > 0) continue; // the re-aligned tile origin, :99
            aln[n++] = tr.ops[t];
        }
        ++tile;
    }
    if (cells) *cells = nCells;
    if (tiles) *tiles = tile;
    if (diagonals) *diagonals = nDiag;
    return n;
}

int twlo_letter_index(char type, char c) {
    if (type == 'p') {
        static const char *aa = "ACDEFGHIKLMNPQRSTVWY";
        if (c == '-' || c == '.') return 21;
        const char *hit = (c != '\0') ? std::strchr(aa, c) : nullptr;
        return hit ? static_cast<int>(hit - aa) : 20;
    }
    switch (c) {
    case 'A': return 0;
    case 'C': return 1;
    case 'G': return 2;
    case 'T': case 'U': return 3;
    case '-': case '.': return 5;
    default: return 4;
    }
}

void twlo_profile_from_rows(char type, int nSeq, const char *const *rows, const float *weights, int alnLen,
                            int alnNum, float nodeWeight, float *profile) {
    const int P = (type == 'n') ? 6 : 22;
    for (int s = 0; s < nSeq; ++s) {
        const float w = weights[s] / nodeWeight * alnNum; // helper.cpp:24
        for (int t = 0; t < alnLen; ++t) {
            const int letter = twlo_letter_index(type, static_cast<char>(std::toupper(static_cast<unsigned char>(rows[s][t]))));
            float &slot = profile[static_cast<size_t>(t) * P + letter];
            slot = static_cast<float>(slot + 1.0 * w); // helper.cpp:30 (double intermediate, as written upstream)
        }
    }
}

void twlo_profile_from_freq(int P, const float *msaFreq, int alnLen, int alnNum, float nodeWeight, float *profile) {
    for (size_t n = 0; n < static_cast<size_t>(alnLen) * P; ++n) profile[n] = msaFreq[n] / nodeWeight * alnNum; // helper.cpp:19
}

void twlo_freq_from_profile(int P, const float *profile, int alnLen, int alnNum, float nodeWeight, float *msaFreq) {
    for (size_t n = 0; n < static_cast<size_t>(alnLen) * P; ++n) msaFreq[n] = profile[n] / alnNum * nodeWeight; // helper.cpp:39
}

void twlo_consensus(int P, const float *profile, int len, char *out) {
    static const char nt[] = "ACGTN";
    static const char aa[] = "ACDEFGHIKLMNPQRSTVWYX";
    const char *lut = (P == 6) ? nt : aa;
    for (int t = 0; t < len; ++t) {
        int best = P - 2;
        float top = 0.0f;
        for (int v = 0; v < P - 2; ++v)
            if (profile[static_cast<size_t>(t) * P + v] > top) { top = profile[static_cast<size_t>(t) * P + v]; best = v; }
        out[t] = lut[best];
    }
}

int twlo_remove_gappy(int P, float *profile, int len, int alnNum, float threshold, int32_t *runs, int *newLen) {
    *newLen = len;
    if (threshold == 1.0f) return 0; // helper.cpp:77
    int nRuns = 0, start = -1;
    for (int t = 0; t < len; ++t) {
        const bool gappy = profile[static_cast<size_t>(t) * P + P - 1] / alnNum > threshold; // helper.cpp:84 (float / int)
        if (gappy) {
            if (start < 0) start = t;
        } else if (start >= 0) {
            runs[2 * nRuns] = start; runs[2 * nRuns + 1] = t - start; ++nRuns; start = -1;
        }
    }
    if (start >= 0) { runs[2 * nRuns] = start; runs[2 * nRuns + 1] = len - start; ++nRuns; }
    if (nRuns == 0) return 0;
    int dst = 0, src = 0, g = 0;
    while (src < len) {
        if (g < nRuns && src == runs[2 * g]) { src += runs[2 * g + 1]; ++g; continue; }
        if (dst != src) std::memmove(profile + static_cast<size_t>(dst) * P, profile + static_cast<size_t>(src) * P, sizeof(float) * P);
        ++dst; ++src;
    }
    *newLen = dst;
    std::fill(profile + static_cast<size_t>(dst) * P, profile + static_cast<size_t>(len) * P, 0.0f);
    return nRuns;
}

void twlo_psgp(int P, const float *profile, int len, int alnNum, float gapOpen, float gapExtend, float *gapOp, float *gapEx) {
    const float scale = (P == 6) ? 0.5f : 1.0f;    // helper.cpp:179
    const float minExtend = gapExtend * 0.2;        // helper.cpp:180 (double product rounded to float)
    const float minOpen = gapOpen * 0.1;
    for (int s = 0; s < len; ++s) {
        const float g = profile[static_cast<size_t>(s) * P + P - 1];
        if (g > 0) { // helper.cpp:188-189: ((num - g) * 1.0 / num) is evaluated in double
            const double keep = (alnNum - g) * 1.0 / alnNum;
            gapOp[s] = std::min(minOpen, static_cast<float>(gapOpen * scale * keep));
            gapEx[s] = std::min(minExtend, static_cast<float>(gapExtend * keep));
        } else {
            gapOp[s] = gapOpen;
            gapEx[s] = gapExtend;
        }
    }
}

int twlo_pairwise_global(char type, const float *score, float gapOpen, float gapExtend, const char *s1, int m,
                         const char *s2, int n, int8_t *path) {
    const int Msz = (type == 'n') ? 5 : 21;
    const size_t W = static_cast<size_t>(n) + 1;
    std::vector<float> Mm((m + 1) * W, 0.0f), X((m + 1) * W, 0.0f), Y((m + 1) * W, 0.0f);
    std::vector<int8_t> tb((m + 1) * W, 0);
    for (int i = 1; i <= m; ++i) { Mm[i * W] = 0; X[i * W] = 0; Y[i * W] = -1e9; tb[i * W] = 2; } // helper.cpp:258-265
    for (int j = 1; j <= n; ++j) { Mm[j] = 0; Y[j] = 0; X[j] = -1e9; tb[j] = 1; }                     // helper.cpp:266-273
    for (int i = 1; i <= m; ++i) {
        for (int j = 1; j <= n; ++j) {
            const int a = twlo_letter_index(type, static_cast<char>(std::toupper(static_cast<unsigned char>(s1[i - 1]))));
            const int b = twlo_letter_index(type, static_cast<char>(std::toupper(static_cast<unsigned char>(s2[j - 1]))));
            const float base = score[a * Msz + b];
            const size_t c = i * W + j, up = (i - 1) * W + j, left = i * W + j - 1, dg = (i - 1) * W + j - 1;
            Mm[c] = base + std::max({Mm[dg], X[dg], Y[dg]});
            X[c] = std::max(Mm[up] + gapOpen, X[up] + gapExtend);
            Y[c] = std::max(Mm[left] + gapOpen, Y[left] + gapExtend);
            const float best = std::max({Mm[c], X[c], Y[c]});
            tb[c] = (best == Mm[c]) ? 0 : ((best == Y[c]) ? 1 : 2);
        }
    }
    int len = 0, i = m, j = n;
    while (i > 0 || j > 0) {
        const int8_t d = tb[i * W + j];
        path[len++] = d;
        if (d == 0) { --i; --j; } else if (d == 1) { --j; } else { --i; }
    }
    std::reverse(path, path + len);
    return len;
}

int twlo_add_gappy_back(char type, const float *score, float gapOpen, float gapExtend, const int8_t *aln, int alnLen,
                        const int32_t *runsRef, int nRunsRef, const int32_t *runsQry, int nRunsQry,
                        const char *consRef, const char *consQry, int8_t *out) {
    int n = 0, r = 0, q = 0, gr = 0, gq = 0;
    for (int a = 0; a <= alnLen; ++a) { // helper.cpp:328 visits one position past the path
        const bool hitR = gr < nRunsRef && r == runsRef[2 * gr];
        const bool hitQ = gq < nRunsQry && q == runsQry[2 * gq];
        if (hitR && hitQ) {
            const int lr = runsRef[2 * gr + 1], lq = runsQry[2 * gq + 1];
            std::vector<int8_t> sub(lr + lq + 1);
            const int sl = twlo_pairwise_global(type, score, gapOpen, gapExtend, consRef + r, lr, consQry + q, lq, sub.data());
            for (int t = 0; t < sl; ++t) out[n++] = sub[t];
            ++gr; ++gq; r += lr; q += lq;
        } else {
            if (hitR) { const int l = runsRef[2 * gr + 1]; for (int t = 0; t < l; ++t) out[n++] = 2; r += l; ++gr; }
            if (hitQ) { const int l = runsQry[2 * gq + 1]; for (int t = 0; t < l; ++t) out[n++] = 1; q += l; ++gq; }
        }
        if (a < alnLen) {
            out[n++] = aln[a];
            if (aln[a] == 0) { ++r; ++q; } else if (aln[a] == 1) { ++q; } else if (aln[a] == 2) { ++r; }
        }
    }
    return n;
}

void twlo_merge_freq(int P, const float *freqRef, const float *freqQry, const int8_t *path, int pathLen,
                     float refWeight, float qryWeight, float *merged) {
    int r = 0, q = 0;
    for (int t = 0; t < pathLen; ++t) {
        float *dst = merged + static_cast<size_t>(t) * P;
        if (path[t] == 0) {
            for (int v = 0; v < P; ++v) dst[v] = freqRef[static_cast<size_t>(r) * P + v] + freqQry[static_cast<size_t>(q) * P + v];
            ++r; ++q;
        } else if (path[t] == 1) {
            for (int v = 0; v < P - 1; ++v) dst[v] = freqQry[static_cast<size_t>(q) * P + v];
            dst[P - 1] = static_cast<float>(freqQry[static_cast<size_t>(q) * P + P - 1] + 1.0 * refWeight);
            ++q;
        } else if (path[t] == 2) {
            for (int v = 0; v < P - 1; ++v) dst[v] = freqRef[static_cast<size_t>(r) * P + v];
            dst[P - 1] = static_cast<float>(freqRef[static_cast<size_t>(r) * P + P - 1] + 1.0 * qryWeight);
            ++r;
        } else {
            for (int v = 0; v < P; ++v) dst[v] = 0.0f;
        }
    }
}

void twlo_update_row(int side, const char *row, const int8_t *path, int pathLen, char *out) {
    const int8_t own = (side == 0) ? 2 : 1;
    int src = 0;
    for (int t = 0; t < pathLen; ++t) out[t] = (path[t] == 0 || path[t] == own) ? row[src++] : '-';
}

} // extern "C"

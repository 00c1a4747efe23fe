// k4_device.cuh — MVH "total predicted peaks" of one peptide row (optionally with one extra modification, Step 5).
// See k4_predict.cu for the reference citations.
#pragma once
#include "k1_device.cuh"

namespace pq {
namespace dev {

constexpr int kMaxLoss = 3;

// MVHMatch.calcMVH (PSMMatch/MVHMatch.java:150-227): every b/y ion the ion factory emits with fewer than two neutral losses
// and no H2O/NH3 loss — plain ions plus one single-loss variant per modification-defined loss carried by the peptide (A8) —
// at fragment charges 1..maxc, kept when 200 <= m/z <= 2000, made unique, then removeFuzzyValues(set, itol) (:104-119)
// counts the values whose successor is at least itol away, plus the last one.  Each (ion type, loss, charge) series is
// ascending in the ion number, so the sorted unique stream is an S-way merge.  `ex` = one more modification at a site
// (Step 5; Unimod-imported modifications carry no neutral loss, A8), site -1 = none.
__device__ inline int32_t predict_peaks(const uint8_t* __restrict__ row, const ModTable* __restrict__ M, double itol, int maxc, Extra ex) {
    const int L = row[0];
    double b[kMaxLen], y[kMaxLen];
    double loss[kMaxLoss]; int nloss = 0;
    auto add_loss = [&](double l) {
        if (l == 0.0) return;
        for (int i = 0; i < nloss; ++i) if (loss[i] == l) return;
        if (nloss < kMaxLoss) loss[nloss++] = l;
    };
    add_loss(M->nterm_loss[row[1]]); add_loss(M->cterm_loss[row[2]]);
    double acc = M->nterm[row[1]];
    if (ex.site == 0) acc = __dadd_rn(acc, ex.delta);
    for (int k = 1; k <= L - 1; ++k) {
        uint32_t code = row[kRowHeader + k - 1];
        acc = __dadd_rn(acc, M->aa[code]);
        double d = M->delta[code];
        if (d != 0.0) acc = __dadd_rn(acc, d);
        if (k == ex.site) acc = __dadd_rn(acc, ex.delta);
        b[k] = acc;
    }
    acc = M->cterm[row[2]];
    if (ex.site == L + 1) acc = __dadd_rn(acc, ex.delta);
    for (int k = 1; k <= L - 1; ++k) {
        uint32_t code = row[kRowHeader + L - k];
        acc = __dadd_rn(acc, M->aa[code]);
        double d = M->delta[code];
        if (d != 0.0) acc = __dadd_rn(acc, d);
        if (L - k + 1 == ex.site) acc = __dadd_rn(acc, ex.delta);
        y[k] = __dadd_rn(acc, kWater);
    }
    for (int k = 0; k < L; ++k) add_loss(M->loss[row[kRowHeader + k]]);

    // series s = ((type * (nloss+1)) + l) * maxc + (c-1)
    const int S = 2 * (nloss + 1) * maxc;
    int idx[2 * (kMaxLoss + 1) * 2];
    double head[2 * (kMaxLoss + 1) * 2];
    auto value = [&](int s, int k) -> double {
        int c = s % maxc + 1; int q = s / maxc; int l = q % (nloss + 1); int t = q / (nloss + 1);
        double m = t ? y[k] : b[k];
        if (l > 0) m = __dsub_rn(m, loss[l - 1]);
        // (m + c p) / c: for c = 1 and 2 — all MVH ever asks for — c p and the division are exact, so no divide sequence is needed
        if (c == 1) return __dadd_rn(m, kProton);
        if (c == 2) return __dmul_rn(__dadd_rn(m, 2.0 * kProton), 0.5);
        return __ddiv_rn(__dadd_rn(m, __dmul_rn((double)c, kProton)), (double)c);
    };
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    auto advance = [&](int s) {          // next in-range value of series s, or +inf
        for (;;) {
            int k = ++idx[s];
            if (k > L - 1) { head[s] = inf; return; }
            double v = value(s, k);
            if (v > 2000.0) { head[s] = inf; idx[s] = L; return; }
            if (v >= 200.0) { head[s] = v; return; }
        }
    };
    for (int s = 0; s < S; ++s) { idx[s] = 0; advance(s); }
    int count = 0; bool have = false; double last = 0.0;
    for (;;) {
        int best = -1; double bv = inf;
        for (int s = 0; s < S; ++s) if (head[s] < bv) { bv = head[s]; best = s; }
        if (best < 0) break;
        if (!have) { have = true; last = bv; }
        else if (bv != last) { if (__dsub_rn(bv, last) >= itol) ++count; last = bv; }
        advance(best);
    }
    if (have) ++count;
    return count;
}

}  // namespace dev
}  // namespace pq

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
// The two fragment ladders (b[k], y[k] for k = 1..L-1, before the charge is applied) and the distinct neutral losses of one row.
__device__ inline void predict_ladders(const uint8_t* __restrict__ row, const ModTable* __restrict__ M, Extra ex, double* b, double* y, double* loss, int& nloss) {
    const int L = row[0];
    nloss = 0;
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
}

// The S-way merge with the number of series (2 ion types x NL <= 2 loss variants x MAXC <= 2 charges) known at compile time.
// series s = ((type * NL) + l) * MAXC + (c-1).  Heads and positions stay in registers; WHICH series advances differs from lane to
// lane, so the advance is driven by data (selects over all series), not by a branch per series: a warp of rows walks the merge in
// lock-step (the first form of this — one predicated block per series — ran with 8 of 32 lanes active).
// (m + c p) / c: for c = 1 and 2 — all MVH ever asks for — c p and the division are exact; subtracting a loss of +0.0 is exact too.
template <int NL, int MAXC>
__device__ __forceinline__ int32_t predict_count_fixed(const double* b, const double* y, int L, const double* loss, double itol) {
    static_assert(NL >= 1 && NL <= 2 && MAXC >= 1 && MAXC <= 2, "fixed merge: at most one loss and two charges");
    constexpr int S = 2 * NL * MAXC;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double loss0 = NL > 1 ? loss[0] : 0.0;
    int idx[S]; double head[S];
    auto advance = [&](int s) {                     // series s moves to its next in-range value (or +inf)
        int k = 0;
#pragma unroll
        for (int j = 0; j < S; ++j) k = s == j ? idx[j] : k;
        const int c = s % MAXC, l = (s / MAXC) % NL, t = s / (MAXC * NL);
        const double* lad = t ? y : b;
        const double lv = l ? loss0 : 0.0;
        double v;
        for (;;) {
            ++k;
            if (k > L - 1) { v = inf; break; }
            const double m = __dsub_rn(lad[k], lv);
            const double v1 = __dadd_rn(m, kProton), v2 = __dmul_rn(__dadd_rn(m, 2.0 * kProton), 0.5);
            v = (MAXC == 1 || c == 0) ? v1 : v2;
            if (v > 2000.0) { v = inf; k = L; break; }
            if (v >= 200.0) break;
        }
#pragma unroll
        for (int j = 0; j < S; ++j) { idx[j] = s == j ? k : idx[j]; head[j] = s == j ? v : head[j]; }
    };
#pragma unroll
    for (int j = 0; j < S; ++j) { idx[j] = 0; head[j] = inf; }
#pragma unroll
    for (int j = 0; j < S; ++j) advance(j);
    int count = 0; bool have = false; double last = 0.0;
    for (;;) {
        int best = -1; double bv = inf;
#pragma unroll
        for (int j = 0; j < S; ++j) if (head[j] < bv) { bv = head[j]; best = j; }
        if (best < 0) break;
        if (!have) { have = true; last = bv; }
        else if (bv != last) { if (__dsub_rn(bv, last) >= itol) ++count; last = bv; }
        advance(best);
    }
    if (have) ++count;
    return count;
}

// any number of losses / charges: the generic merge
__device__ inline int32_t predict_count_generic(const double* b, const double* y, int L, const double* loss, int nloss, double itol, int maxc) {
    const int S = 2 * (nloss + 1) * maxc;
    int idx[2 * (kMaxLoss + 1) * 2];
    double head[2 * (kMaxLoss + 1) * 2];
    auto value = [&](int s, int k) -> double {
        int c = s % maxc + 1; int q = s / maxc; int l = q % (nloss + 1); int t = q / (nloss + 1);
        double m = t ? y[k] : b[k];
        if (l > 0) m = __dsub_rn(m, loss[l - 1]);
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
__device__ inline int32_t predict_count(const double* b, const double* y, int L, const double* loss, int nloss, double itol, int maxc) {
    if (maxc == 1) { if (nloss == 0) return predict_count_fixed<1, 1>(b, y, L, loss, itol); if (nloss == 1) return predict_count_fixed<2, 1>(b, y, L, loss, itol); }
    if (maxc == 2) { if (nloss == 0) return predict_count_fixed<1, 2>(b, y, L, loss, itol); if (nloss == 1) return predict_count_fixed<2, 2>(b, y, L, loss, itol); }
    return predict_count_generic(b, y, L, loss, nloss, itol, maxc);
}

__device__ inline int32_t predict_peaks(const uint8_t* __restrict__ row, const ModTable* __restrict__ M, double itol, int maxc, Extra ex) {
    double b[kMaxLen], y[kMaxLen], loss[kMaxLoss]; int nloss;
    predict_ladders(row, M, ex, b, y, loss, nloss);
    return predict_count(b, y, (int)row[0], loss, nloss, itol, maxc);
}

}  // namespace dev
}  // namespace pq

// k1_device.cuh — device-side pieces of the pair scorer shared by K1 (k1_score.cu) and the Step-5 kernel (k5_ptm.cu):
// TMA/mbarrier helpers, the prepared-spectrum view, the annotator's peak pick, the fragment-ladder walk and the
// precursor-window predicate.  See k1_score.cu for the reference citations.
#pragma once
#include "pq_internal.h"
#include "k_cells.cuh"

namespace pq {
namespace dev {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// pulls a global range towards L2 (no destination in the SM): the bulk copy that follows finds it there
__device__ __forceinline__ void prefetch_l2_bulk(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

struct BlobView {
    const double* ev; const uint16_t* ans; const double* pmz; const double* pin; const uint16_t* pcode; const double* val; const uint16_t* cells; const uint32_t* cbits;
    uint32_t E; uint32_t nsurv; uint32_t flags; int32_t cls[4];
};
__device__ __forceinline__ BlobView view_blob(const uint8_t* b) {
    const BlobHeader* h = reinterpret_cast<const BlobHeader*>(b);
    BlobView v;
    v.E = h->n_peaks; v.nsurv = h->n_survivors; v.flags = h->flags;
    for (int c = 0; c < 4; ++c) v.cls[c] = h->cls_count[c];
    const BlobLayout L = blob_layout(v.E);
    v.ev = reinterpret_cast<const double*>(b + L.ev);
    v.ans = reinterpret_cast<const uint16_t*>(b + L.ans);
    v.pmz = reinterpret_cast<const double*>(b + L.pmz);
    v.pin = reinterpret_cast<const double*>(b + L.pin);
    v.pcode = reinterpret_cast<const uint16_t*>(b + L.pcode);
    v.val = reinterpret_cast<const double*>(b + L.val);
    v.cells = reinterpret_cast<const uint16_t*>(b + L.cells);
    v.cbits = reinterpret_cast<const uint32_t*>(b + L.cbits);
    return v;
}

// The annotator's peak pick for one theoretical m/z (A2-A4): most intense matchable peak with |mz - t| <= itol, equal
// intensity -> smaller |error|.  K0 has inverted the predicate into exact break points (pq_common.h): find the interval
// of t in the event list and read its answer.  Returns the claim code (kAnsNone = nothing to claim).
// METHOD 1: code = survivor id; METHOD 2: code = (class << 12) | survivor id, m/z range already applied.
// Tie scan: two in-window peaks have equal intensity, the reference's scan decides (the first peak wins unless a later
// one is strictly closer to t).  Rare, so it stays out of line: the hot loop keeps a small instruction footprint.
template <int METHOD>
__device__ __noinline__ uint32_t match_ion_tie(const double* pmz, const double* pin, const uint16_t* pcode, uint32_t first, double t, double itol) {
    uint32_t code = kNoSurvivor; double best_in = __longlong_as_double(0xfff0000000000000LL), best_abs = 0.0, best_mz = 0.0;   // -inf
    for (uint32_t j = first;; ++j) {
        double m = pmz[j];                                 // +inf after the last peak ends the scan
        double err = __dsub_rn(m, t);
        if (err > itol) break;
        double ae = fabs(err);
        if (ae <= itol) {
            double v = pin[j];
            if (v > best_in || (v == best_in && ae < best_abs)) { best_in = v; best_abs = ae; best_mz = m; code = pcode[j]; }
        }
    }
    if ((code & 0xFFFu) == kNoSurvivor) return kAnsNone;
    if (METHOD == 2 && (best_mz < 200.0 || best_mz > 2000.0)) return kAnsNone;      // MVHMatch.java:248-250
    return code;
}

template <int METHOD>
__device__ __forceinline__ uint32_t match_ion(const BlobView& B, double t, double itol) {
    uint32_t k = B.cells[cell_of(t)] & 0x7FFFu;
    while (B.ev[k + 1] <= t) ++k;
    uint32_t a = B.ans[k];
    if (a < kAnsTie || a == kAnsNone) return a;
    return match_ion_tie<METHOD>(B.pmz, B.pin, B.pcode, a & 0xFFFu, t, itol);
}
// Most theoretical m/z fall into cells where nothing can be claimed (kCellClean): one table load answers them.
__device__ __forceinline__ bool ion_may_match(const BlobView& B, double t) { return reinterpret_cast<const int16_t*>(B.cells)[cell_filter(t)] >= 0; }   // kCellClean = sign bit
// the same answer from the bit map (a blob read from global memory: 320 bytes instead of the 5 KB table)
__device__ __forceinline__ bool ion_may_match_bits(const BlobView& B, double t) { const uint32_t c = cell_filter(t); return !((__ldg(B.cbits + (c >> 5)) >> (c & 31u)) & 1u); }

struct Extra { int32_t site; double delta; };   // Step 5: one extra modification (site -1 = none)

struct PairScore { double score; int32_t a, b; };

// The score from what the claims left behind (shared by the ladder-walking scorer below and the table-driven one of k1_ladder.cu).
// METHOD 1: HyperscoreMatch.java:150-179; METHOD 2: MVHMatch.java:281-299.
template <int METHOD>
__device__ __forceinline__ PairScore finish_pair(const BlobView& B, int ca, int cb, double sum, int key0, int key1, int key2, int32_t predicted,
                                                 const double* log10_fact, const double* ln_fact) {
    PairScore r;
    if (METHOD == 1) {
        if (ca > 64) ca = 64;
        if (cb > 64) cb = 64;
        r.a = ca; r.b = cb;
        if (ca == 0 && cb == 0) { r.score = 0.0; return r; }
        double hs = jlog10(sum);
        if (ca > 0 && cb > 0) hs = __dadd_rn(__dadd_rn(hs, log10_fact[ca]), log10_fact[cb]);
        else if (cb > 0) hs = __dadd_rn(hs, log10_fact[cb]);
        else hs = __dadd_rn(hs, log10_fact[ca]);
        r.score = __dmul_rn(4.0, hs);
    } else {
        int matched = key0 + key1 + key2;
        r.a = matched; r.b = predicted;
        r.score = 0.0;
        if (!(B.flags & 1u)) { r.a = 0; r.b = 0; return r; }     // fewer than 7 peaks survived: calcMVH returns before scoring
        if (predicted <= 0) return r;
        int key[4] = {key0, key1, key2, predicted - matched};
        int N = (int)B.nsurv;
        double mvh = 0.0; bool died = false;
        for (int i = 0; i < 4; ++i) {
            int n = B.cls[i], k = key[i];
            if (n < 0 || k < 0) continue;
            if (n < k) { died = true; continue; }
            mvh = __dadd_rn(mvh, __dsub_rn(__dsub_rn(ln_fact[n], ln_fact[n - k]), ln_fact[k]));
        }
        int tb = B.cls[3] + N;
        if (tb >= 0 && predicted >= 0) {
            if (tb < predicted) died = true;
            else mvh = __dsub_rn(mvh, __dsub_rn(__dsub_rn(ln_fact[tb], ln_fact[tb - predicted]), ln_fact[predicted]));
        }
        mvh = __dsub_rn(0.0, mvh);
        r.score = died ? 0.0 : mvh;
    }
    return r;
}

template <int BLOCK>
__device__ __forceinline__ uint32_t row_byte(const uint32_t* codes, int idx) {
    uint32_t w = codes[(idx >> 2) * BLOCK];
    return (w >> ((idx & 3) * 8)) & 0xFFu;
}

// Walks the fragment ladder of one peptide row against the staged spectrum.
// METHOD 1: hyperscore (count_b, count_y, sum of normalised intensities); METHOD 2: MVH class hits.
// NCH = number of fragment charges (precursor charge - 1, JMatch.java:398-407) when it is 1, 2 or 3: the charge loop
// unrolls and charges 1 and 2 need no divide; NCH = 0 takes the charge count at run time.
// tab[code * TS] = {residue mass, modification mass} (one 16-byte shared-memory load per residue).  TS = 8: the table holds 8
// copies of every entry, copy (lane & 7) at tab[code * 8 + (lane & 7)], and `tab` arrives offset by (lane & 7): the eight lanes
// of a quarter-warp then hit eight different 16-byte bank groups whatever their residues are (no bank conflicts).
template <int BLOCK, int METHOD, int NCH, int TS = 1, bool BITS = false>
__device__ __forceinline__ PairScore score_pair_n(const uint32_t* codes /* this thread's column */, const BlobView& B,
                                                  const double2* tab, const double* ntd, const double* ctd,
                                                  int z, double itol, Extra ex,
                                                  const double* log10_fact, const double* ln_fact, int32_t predicted,
                                                  uint32_t* usedw /* MVH: this thread's column of the used bitmap */,
                                                  double cutoff /* > 0: only "score >= cutoff" matters (Step 4) */) {
    const uint32_t head = codes[0];
    const int L = (int)(head & 0xFF);
    const int nch = NCH > 0 ? NCH : (z <= 1 ? 1 : z - 1);
    uint64_t used = 0ull;
    int ca = 0, cb = 0;
    int key0 = 0, key1 = 0, key2 = 0;
    double sum = 0.0;
    if (METHOD == 2) for (int w = 0; w < 10; ++w) usedw[w * BLOCK] = 0u;

    // Claim in annotator order (A5): a surviving peak counts once, for the first ion that reaches it.
    // (A branch-free variant — a miss adding +0.0 — was measured slower on the Step-4 launch: most lanes miss.)
    auto claim_value = [&](uint32_t code, bool is_b, double val_of_code) {      // METHOD 1 with B.val[code] already in hand
        if (code == kAnsNone) return;
        if (METHOD == 1) {
            uint64_t bit = 1ull << code;
            if (used & bit) return;
            used |= bit;
            if (is_b) ++ca; else ++cb;
            sum = __dadd_rn(sum, val_of_code);
        } else {
            uint32_t sid = code & 0xFFFu;
            uint32_t w = usedw[(sid >> 5) * BLOCK], bit = 1u << (sid & 31);
            if (w & bit) return;
            usedw[(sid >> 5) * BLOCK] = w | bit;
            uint32_t c = (code >> 12) & 3u;
            key0 += c == 0; key1 += c == 1; key2 += c == 2;
        }
    };
    auto claim = [&](uint32_t code, bool is_b) {
        if (code == kAnsNone) return;
        if (METHOD == 1) {
            uint64_t bit = 1ull << code;
            if (used & bit) return;
            used |= bit;
            if (is_b) ++ca; else ++cb;
            sum = __dadd_rn(sum, B.val[code]);
        } else {
            uint32_t sid = code & 0xFFFu;
            uint32_t w = usedw[(sid >> 5) * BLOCK], bit = 1u << (sid & 31);
            if (w & bit) return;
            usedw[(sid >> 5) * BLOCK] = w | bit;
            uint32_t c = (code >> 12) & 3u;
            key0 += c == 0; key1 += c == 1; key2 += c == 2;
        }
    };
    // All charge states of one ion, charge ascending (A5).  m/z at charge c = (m + c*proton) / c: c*proton is exact for
    // c = 1, 2, 4 and so is the division by a power of two, so only c = 3 (and >= 5) needs a real FP64 divide.
    // A6 chargeValidated: c > 1 needs c <= ion number k.
    // Only ~1 theoretical m/z in 10 lands in a cell where something can be claimed.  Those are queued (per thread, in ion
    // order) and resolved after the ladder walk: the warp runs the long look-up as often as its busiest lane needs it,
    // not once per ion.  Claims happen in queue order = annotator order, and a miss claims nothing, so nothing changes.
    // When only "score >= cutoff" matters (Step 4 counts shuffles that reach the target's score), both series are held in
    // the queue: hyperscore is monotone in the two counts and the intensity sum, the counts cannot exceed the queued
    // possible matches (nor the survivor count together) and a normalised intensity cannot exceed 100, so
    // 4 (log10(100 n) + log10 nb! + log10 ny!) bounds the score from above.  A pair whose bound stays below the cutoff is
    // reported as 0 without resolving a single match (it could never count).
    constexpr int kQueue = 32;
    double q[kQueue]; int qn = 0;
    bool in_b = true; int nb_q = 0;
    bool holding = METHOD == 1 && cutoff > 0.0;
    auto flush = [&]() {                                       // resolve everything queued, in ion order
        const int split = in_b ? qn : nb_q;
        if (BITS) {
            // The blob sits in global memory (Step 2: a thread per spectrum): one ion's look-up is a chain of dependent misses
            // (cell table -> events -> answer -> intensity), ~1 us each.  Four ions go through each stage together, so the misses
            // of a stage overlap; the claims still happen one after the other in ion order.
#pragma unroll 1
            for (int i0 = 0; i0 < qn; i0 += 4) {
                double t[4]; uint32_t k[4]; double e[4]; uint32_t a[4]; double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { t[u] = q[min(i0 + u, qn - 1)]; k[u] = __ldg(B.cells + cell_of(t[u])) & 0x7FFFu; }
#pragma unroll
                for (int u = 0; u < 4; ++u) e[u] = __ldg(B.ev + k[u] + 1);
#pragma unroll
                for (int u = 0; u < 4; ++u) { while (e[u] <= t[u]) { ++k[u]; e[u] = __ldg(B.ev + k[u] + 1); } a[u] = __ldg(B.ans + k[u]); }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (a[u] >= kAnsTie && a[u] != kAnsNone) a[u] = match_ion_tie<METHOD>(B.pmz, B.pin, B.pcode, a[u] & 0xFFFu, t[u], itol);
                    v[u] = (METHOD == 1 && a[u] != kAnsNone) ? __ldg(B.val + a[u]) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) if (i0 + u < qn) claim_value(a[u], i0 + u < split, v[u]);
            }
        } else {
            for (int i = 0; i < qn; ++i) claim(match_ion<METHOD>(B, q[i], itol), i < split);
        }
        qn = 0; nb_q = 0;
    };
    auto push = [&](double t) { if (BITS ? ion_may_match_bits(B, t) : ion_may_match(B, t)) q[qn++] = t; };
    auto ion = [&](double m, int k) {
        push(__dadd_rn(m, kProton));
        if (nch >= 2 && k >= 2) push(__dmul_rn(__dadd_rn(m, 2.0 * kProton), 0.5));
        if (nch >= 3 && k >= 3) push(div3_rn(__dadd_rn(m, __dmul_rn(3.0, kProton))));
        if (NCH == 0) {
            if (nch >= 4 && k >= 4) push(__dmul_rn(__dadd_rn(m, 4.0 * kProton), 0.25));
            for (int c = 5; c <= nch && c <= k; ++c) {
                if (qn >= kQueue) { flush(); holding = false; }
                push(__ddiv_rn(__dadd_rn(m, __dmul_rn((double)c, kProton)), (double)c));
            }
        }
    };

    // ---- b ions: forward sum, residue mass then its modification, N-term modification opens the sum ----
    // (a plain residue has modification mass +0.0: adding it leaves the positive running sum unchanged, so no branch)
    // Residue i (0-based) is byte i & 3 of row word 1 + i / 4: one shared-memory word feeds four ladder steps.
    // The inner loops stay rolled and the tie scan out of line: a small instruction footprint measured faster than
    // unrolled ladders, and separate b / y loops faster than one shared body.
    double acc = ntd[(head >> 8) & 0xFF];
    if (ex.site == 0) acc = __dadd_rn(acc, ex.delta);
    for (int i0 = 0; i0 < L - 1; i0 += 4) {
        if (qn > kQueue - 16) { flush(); holding = false; }        // room for the next four ions (up to four charges each)
        uint32_t word = codes[(1 + (i0 >> 2)) * BLOCK];
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int k = i0 + j + 1;
            if (k > L - 1) break;
            const double2 t = tab[(word & 0xFFu) * TS];
            word >>= 8;
            acc = __dadd_rn(__dadd_rn(acc, t.x), t.y);
            if (k == ex.site) acc = __dadd_rn(acc, ex.delta);
            ion(acc, k);
        }
    }
    if (holding) nb_q = qn; else flush();
    in_b = false;
    // ---- y ions: rewind sum from the C-terminus, + H2O ----
    acc = ctd[(head >> 16) & 0xFF];
    if (ex.site == L + 1) acc = __dadd_rn(acc, ex.delta);
    for (int i = L - 1; i >= 1;) {
        if (qn > kQueue - 16) { flush(); holding = false; }
        uint32_t word = codes[(1 + (i >> 2)) * BLOCK] << (8 * (3 - (i & 3)));        // residue i in the top byte
#pragma unroll 1
        for (int j = i & 3; j >= 0 && i >= 1; --j, --i) {
            const double2 t = tab[(word >> 24) * TS];
            word <<= 8;
            acc = __dadd_rn(__dadd_rn(acc, t.x), t.y);
            if (i + 1 == ex.site) acc = __dadd_rn(acc, ex.delta);
            ion(__dadd_rn(acc, kWater), L - i);
        }
    }
    if (holding) {
        const int nb = min(nb_q, 64), ny = min(qn - nb_q, 64), n = min(min(nb_q + (qn - nb_q), (int)B.nsurv), 64);
        double ub = 0.0;
        if (n > 0) ub = 4.0 * (log10_fact[65 + n] + log10_fact[nb] + log10_fact[ny]);        // log10_fact[65 + n] = log10(100 n)
        if (ub * (1.0 + 1e-12) + 1e-9 < cutoff) { PairScore r0; r0.score = 0.0; r0.a = -1; r0.b = 0; return r0; }      // a = -1: decided by the bound
    }
    flush();

    return finish_pair<METHOD>(B, ca, cb, sum, key0, key1, key2, predicted, log10_fact, ln_fact);
}

template <int BLOCK, int METHOD, int TS = 1, bool BITS = false>
__device__ __forceinline__ PairScore score_pair(const uint32_t* codes, const BlobView& B, const double2* tab, const double* ntd, const double* ctd,
                                                int z, double itol, Extra ex, const double* log10_fact, const double* ln_fact, int32_t predicted,
                                                uint32_t* usedw, double cutoff = 0.0) {
    switch (z) {                                  // fragment charges = precursor charge - 1, at least 1 (JMatch.java:398-407)
        case 0: case 1: case 2: return score_pair_n<BLOCK, METHOD, 1, TS, BITS>(codes, B, tab, ntd, ctd, z, itol, ex, log10_fact, ln_fact, predicted, usedw, cutoff);
        case 3: return score_pair_n<BLOCK, METHOD, 2, TS, BITS>(codes, B, tab, ntd, ctd, z, itol, ex, log10_fact, ln_fact, predicted, usedw, cutoff);
        case 4: return score_pair_n<BLOCK, METHOD, 3, TS, BITS>(codes, B, tab, ntd, ctd, z, itol, ex, log10_fact, ln_fact, predicted, usedw, cutoff);
        default: return score_pair_n<BLOCK, METHOD, 0, TS, BITS>(codes, B, tab, ntd, ctd, z, itol, ex, log10_fact, ln_fact, predicted, usedw, cutoff);
    }
}

// Shared-memory loads by 32-bit shared address.  The count pass below is a few instructions per ion, and converting a generic
// pointer to its shared address inside the unrolled steps (the compiler re-derives it each time rather than hold a register)
// showed up as 10 % of K1's instructions; one conversion per row and explicit ld.shared avoids that.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }   // zero-extended
__device__ __forceinline__ double2 lds_f64x2(uint32_t a) { double2 v; asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a)); return v; }

// Count pass of the score bound (Step 4, no extra modification): walks the same fragment ladder as score_pair_n — the same sums
// in the same order, so the same theoretical m/z — but only counts the ions that land in a cell where something can be
// claimed.  Returns possible b matches | possible y matches << 16.  No queue, no claims: this is all a shuffle costs when
// its bound 4 (log10(100 n) + log10 nb! + log10 ny!) stays below the target's score.
// Four residues per row word.  Words whose four ions all carry every fragment charge (ion number >= NCH) and all exist run
// as straight-line code; the first ions of each series and the ragged ends take the guarded step.  Clean cells are counted
// through the flag bit and subtracted from the number of ion / charge combinations.
template <int BLOCK, int NCH, int TS = 1, bool HAS_EX = false>
__device__ __forceinline__ uint32_t count_possible_n(const uint32_t* codes, const BlobView& B, const double2* tab, const double* ntd,
                                                     const double* ctd, int z, Extra ex = Extra{-1, 0.0}) {
    const uint32_t head = codes[0];
    const int L = (int)(head & 0xFF), Lm1 = L - 1;
    const int nch = NCH > 0 ? NCH : (z <= 1 ? 1 : z - 1);
    const uint32_t cells_s = smem_addr(B.cells), tab_s = smem_addr(tab);
    auto clean = [&](double t) -> uint32_t { return lds_u16(cells_s + 2u * cell_filter(t)) >> 15; };   // 1: nothing can be claimed at t
    auto residue = [&](uint32_t word, int j) -> double2 { return lds_f64x2(tab_s + ((word >> (8 * j)) & 0xFFu) * (16u * TS)); };
    auto ion = [&](double m, int k) -> uint32_t {                        // any ion number (A6: charge c needs c <= k)
        uint32_t c = clean(__dadd_rn(m, kProton));
        if (nch >= 2 && k >= 2) c += clean(__dmul_rn(__dadd_rn(m, 2.0 * kProton), 0.5));
        if (nch >= 3 && k >= 3) c += clean(div3_rn(__dadd_rn(m, __dmul_rn(3.0, kProton))));
        if (NCH == 0) {
            if (nch >= 4 && k >= 4) c += clean(__dmul_rn(__dadd_rn(m, 4.0 * kProton), 0.25));
            for (int f = 5; f <= nch && f <= k; ++f) c += clean(__ddiv_rn(__dadd_rn(m, __dmul_rn((double)f, kProton)), (double)f));
        }
        return c;
    };
    auto ion_full = [&](double m, int k) -> uint32_t {                   // ion number >= NCH: every charge exists
        if (NCH == 0) return ion(m, k);
        uint32_t c = clean(__dadd_rn(m, kProton));
        if (NCH >= 2) c += clean(__dmul_rn(__dadd_rn(m, 2.0 * kProton), 0.5));
        if (NCH >= 3) c += clean(div3_rn(__dadd_rn(m, __dmul_rn(3.0, kProton))));
        return c;
    };
    // ion / charge combinations of one series: sum over k = 1..L-1 of min(nch, k)
    const int full = max(Lm1 - nch, 0), ramp = min(nch, Lm1);
    const int per_side = full * nch + ramp * (ramp + 1) / 2;
    uint32_t cb = 0, cy = 0;                                // minus the clean ones

    // ---- b ions: residue k - 1 closes ion k; word g holds ions 4g+1 .. 4g+4 ----
    double acc = ntd[(head >> 8) & 0xFF];
    if (HAS_EX && ex.site == 0) acc = __dadd_rn(acc, ex.delta);
    {
        const uint32_t word = codes[1 * BLOCK];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j + 1 <= Lm1) {
                const double2 t = residue(word, j);
                acc = __dadd_rn(__dadd_rn(acc, t.x), t.y);
                if (HAS_EX && j + 1 == ex.site) acc = __dadd_rn(acc, ex.delta);
                cb += ion(acc, j + 1);
            }
        }
    }
    const int G = Lm1 >> 2;                                 // words 0 .. G-1 are full
#pragma unroll 1
    for (int g = 1; g < G; ++g) {
        const uint32_t word = codes[(1 + g) * BLOCK];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double2 t = residue(word, j);
            acc = __dadd_rn(__dadd_rn(acc, t.x), t.y);
            if (HAS_EX && 4 * g + j + 1 == ex.site) acc = __dadd_rn(acc, ex.delta);
            cb += ion_full(acc, 4 * g + j + 1);
        }
    }
    if (G >= 1 && (Lm1 & 3)) {
        const uint32_t word = codes[(1 + G) * BLOCK];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (4 * G + j + 1 <= Lm1) {
                const double2 t = residue(word, j);
                acc = __dadd_rn(__dadd_rn(acc, t.x), t.y);
                if (HAS_EX && 4 * G + j + 1 == ex.site) acc = __dadd_rn(acc, ex.delta);
                cb += ion_full(acc, 4 * G + j + 1);
            }
        }
    }
    // ---- y ions: residue i (from L-1 down to 1) closes ion L - i ----
    acc = ctd[(head >> 16) & 0xFF];
    if (HAS_EX && ex.site == L + 1) acc = __dadd_rn(acc, ex.delta);
    int i = Lm1;
#pragma unroll 1
    for (; i >= 1 && ((L - i) < nch || (i & 3) != 3); --i) {           // first ions and the ragged top word
        const double2 t = residue(codes[(1 + (i >> 2)) * BLOCK], i & 3);
        acc = __dadd_rn(__dadd_rn(acc, t.x), t.y);
        if (HAS_EX && i + 1 == ex.site) acc = __dadd_rn(acc, ex.delta);
        cy += ion(__dadd_rn(acc, kWater), L - i);
    }
    if (i >= 3) {                                                       // here i & 3 == 3
#pragma unroll 1
        for (int w = i >> 2; w >= 1; --w) {
            const uint32_t word = codes[(1 + w) * BLOCK];
#pragma unroll
            for (int j = 3; j >= 0; --j) {
                const double2 t = residue(word, j);
                acc = __dadd_rn(__dadd_rn(acc, t.x), t.y);
                if (HAS_EX && 4 * w + j + 1 == ex.site) acc = __dadd_rn(acc, ex.delta);
                cy += ion_full(__dadd_rn(acc, kWater), L - (4 * w + j));
            }
        }
        const uint32_t word = codes[1 * BLOCK];
#pragma unroll
        for (int j = 3; j >= 1; --j) {
            const double2 t = residue(word, j);
            acc = __dadd_rn(__dadd_rn(acc, t.x), t.y);
            if (HAS_EX && j + 1 == ex.site) acc = __dadd_rn(acc, ex.delta);
            cy += ion_full(__dadd_rn(acc, kWater), L - j);
        }
    }
    const uint32_t nb = (uint32_t)per_side - cb, ny = (uint32_t)per_side - cy;
    return min(nb, 0xFFFFu) | (min(ny, 0xFFFFu) << 16);
}

// The decision itself, from the numbers of possible b / y matches.
// Hyperscore: 4 (log10(100 n) + log10 nb! + log10 ny!).  MVH (MVHMatch.java:281-299): score = lnC(bins, predicted) - sum over the
// three intensity classes of lnC(classCount, hits) - lnC(emptyBins, predicted - matched); every lnC is >= 0 and, for
// predicted <= emptyBins / 2, lnC(emptyBins, j) grows with j, so with matched <= m = min(possible, survivors, predicted)
// the score is at most lnC(bins, predicted) - lnC(emptyBins, predicted - m).
template <int METHOD>
__device__ __forceinline__ bool bound_below_cutoff(const BlobView& B, int nbq, int nyq, const double* log10_fact, double cutoff,
                                                   const double* ln_fact, int32_t predicted) {
    if (METHOD == 2) {
        const int N = (int)B.nsurv, empty = B.cls[3], bins = empty + N;
        const int m = min(min(nbq + nyq, N), predicted), j = predicted - m;
        auto lnC = [&](int n, int k) { return __dsub_rn(__dsub_rn(ln_fact[n], ln_fact[n - k]), ln_fact[k]); };
        const double ub = __dsub_rn(lnC(bins, predicted), lnC(empty, j));
        return ub * (1.0 + 1e-12) + 1e-6 < cutoff;
    }
    const int nb = min(nbq, 64), ny = min(nyq, 64), n = min(min(nbq + nyq, (int)B.nsurv), 64);
    double ub = 0.0;
    if (n > 0) ub = 4.0 * (log10_fact[65 + n] + log10_fact[nb] + log10_fact[ny]);        // log10_fact[65 + n] = log10(100 n)
    return ub * (1.0 + 1e-12) + 1e-9 < cutoff;
}
// MVH rows the bound cannot judge, or that the scorer answers with 0 anyway: 1 = below (decided), 0 = score it, -1 = go on and count
template <int METHOD>
__device__ __forceinline__ int bound_precheck(const BlobView& B, int32_t predicted) {
    if (METHOD == 2) {
        if (!(B.flags & 1u) || predicted <= 0) return 1;                // the scorer returns 0 for these
        const int N = (int)B.nsurv, empty = B.cls[3], bins = empty + N;
        if (empty < 0 || 2 * predicted > empty) return 0;               // no bound: score it
        if (bins < predicted) return 1;                                  // "died": the scorer returns 0
    }
    return -1;
}

// true: the score of this row cannot reach `cutoff` (see count_possible_n); false: score it.
template <int BLOCK, int TS = 1, int METHOD = 1, bool HAS_EX = false>
__device__ __forceinline__ bool below_score_bound(const uint32_t* codes, const BlobView& B, const double2* tab, const double* ntd,
                                                  const double* ctd, int z, const double* log10_fact, double cutoff,
                                                  const double* ln_fact = nullptr, int32_t predicted = 0, Extra ex = Extra{-1, 0.0}) {
    if (!(cutoff > 0.0)) return false;
    const int pre = bound_precheck<METHOD>(B, predicted);
    if (pre >= 0) return pre != 0;
    uint32_t c;
    switch (z) {
        case 0: case 1: case 2: c = count_possible_n<BLOCK, 1, TS, HAS_EX>(codes, B, tab, ntd, ctd, z, ex); break;
        case 3: c = count_possible_n<BLOCK, 2, TS, HAS_EX>(codes, B, tab, ntd, ctd, z, ex); break;
        case 4: c = count_possible_n<BLOCK, 3, TS, HAS_EX>(codes, B, tab, ntd, ctd, z, ex); break;
        default: c = count_possible_n<BLOCK, 0, TS, HAS_EX>(codes, B, tab, ntd, ctd, z, ex); break;
    }
    return bound_below_cutoff<METHOD>(B, (int)(c & 0xFFFFu), (int)(c >> 16), log10_fact, cutoff, ln_fact, predicted);
}

// The same decision from the row's cell lists (k_cells.cuh): K2 walked the ladder once when it wrote the row; here every
// (ion, fragment charge) is one 2-byte entry — the byte offset of its cell in the staged spectrum's table, whose top bit says
// "nothing can be claimed in this cell" — K1 spreads those bits into a byte per cell once per job, so the test is one byte
// load.  Whole 16-byte units of eight entries; slots that are no ion point at the always-clean sentinel cell behind that table.
// Both sides at once: the b and the y unit of the same index are loaded together (and two indices per trip), so four 16-byte
// loads are in flight per thread instead of one — the lists come from L2, the pass is latency-bound otherwise.
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }   // zero-extended
__device__ __forceinline__ uint2 count_clean_lists(const uint4* __restrict__ pb, const uint4* __restrict__ py, size_t stride, uint32_t n_units, uint32_t cflag_s) {
    uint32_t cb = 0, cy = 0;
    auto flag = [&](uint32_t e) -> uint32_t { return lds_u8(cflag_s + e); };      // the CTA's byte table: 1 = nothing can be claimed in that cell
    auto unit = [&](const uint4& q) -> uint32_t {
        return flag(q.x & 0xFFFFu) + flag(q.x >> 16) + flag(q.y & 0xFFFFu) + flag(q.y >> 16) + flag(q.z & 0xFFFFu) + flag(q.z >> 16) + flag(q.w & 0xFFFFu) + flag(q.w >> 16);
    };
#pragma unroll 2
    for (uint32_t i = 0; i < n_units; ++i) {
        const uint4 qb = __ldg(pb + (size_t)i * stride), qy = __ldg(py + (size_t)i * stride);
        cb += unit(qb); cy += unit(qy);
    }
    return make_uint2(cb, cy);
}
// n_units = streams read x units per stream; n_side = ions x charges of one side; the rest of the 8 * n_units slots are sentinel slots
template <int METHOD>
__device__ __forceinline__ bool below_score_bound_lists(const uint4* __restrict__ row_lists, size_t stride, uint32_t chunks_side, uint32_t n_units, uint32_t n_side,
                                                        const BlobView& B, uint32_t cflag_s, const double* log10_fact, double cutoff, const double* ln_fact, int32_t predicted) {
    if (!(cutoff > 0.0)) return false;
    const int pre = bound_precheck<METHOD>(B, predicted);
    if (pre >= 0) return pre != 0;
    const uint2 clean = count_clean_lists(row_lists, row_lists + (size_t)chunks_side * stride, stride, n_units, cflag_s);
    const uint32_t slots = 8u * n_units;                          // the sentinel slots count as clean, so whatever is not clean is an ion that may match
    const uint32_t nb = slots - clean.x, ny = slots - clean.y;
    (void)n_side;
    return bound_below_cutoff<METHOD>(B, (int)min(nb, 0xFFFFu), (int)min(ny, 0xFFFFu), log10_fact, cutoff, ln_fact, predicted);
}

// Targeted PSM list in Step 2 (FindCandidateSpectraAndScoreWorker.java:60-74): false = the row's peptide names its spectra and this is
// not one of them (the pair is not scored); a named pair on a spectrum that came with a charge is marked "seen" (:69-74).
__device__ __forceinline__ bool targeted_allows(const TargetedView& V, uint32_t row, int32_t spectrum_sorted) {
    const uint32_t seq = V.canon[V.seq_of_form[V.form_of_row[row]]];
    if (!V.seq_listed[seq]) return true;
    const uint64_t key = ((uint64_t)seq << 32) | (uint32_t)V.caller_index[spectrum_sorted];
    uint32_t lo = 0, hi = V.n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (V.keys[mid] < key) lo = mid + 1; else hi = mid; }
    if (lo >= V.n || V.keys[lo] != key) return false;
    if (V.alias[spectrum_sorted] == 0) V.seen[lo] = 1;
    return true;
}

// Precursor-window predicate exactly as DBSearchWorker.java:38-52 / FindCandidate...java:47-58:
// bucket round(10*pepMass) within +-1 of round(10*mass) AND |pepMass - mass| (ppm of pepMass) <= tol.
__device__ __forceinline__ bool in_window(double pm, double mass, long long va, const ScoreConfig& cfg) {
    if (!cfg.no_bucket) {
        long long key = jround(__dmul_rn(pm, 10.0));
        if (key < va - 1 || key > va + 1) return false;
    }
    double del = fabs(__dsub_rn(pm, mass));
    if (cfg.tol_ppm) del = __ddiv_rn(__dmul_rn(__dmul_rn(1.0e6, 1.0), del), pm);
    return del <= cfg.tol;
}


}  // namespace dev
}  // namespace pq

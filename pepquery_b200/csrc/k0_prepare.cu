// k0_prepare.cu — K0: peptide-independent spectrum pre-processing, hoisted out of the pair loop.
//
// The reference redoes this for every (peptide, spectrum) pair inside scorePeptide2Spectrum
// (pg/PeptideSearchMT.java:1144-1251): HyperscoreMatch.SpectrumPreProcessing (PSMMatch/HyperscoreMatch.java:47-59),
// MVHMatch.SpectrumPreProcessing + classifyPeakIntensities (PSMMatch/MVHMatch.java:46-89) over the JMatch filters
// (PSMMatch/JMatch.java:101-292,341-383), plus the annotator's intensity threshold and peak pick (JMatch.java:411-451).
// It depends only on the spectrum, so here it runs once per spectrum that has at least one candidate and writes one
// "prepared spectrum" blob (pq_common.h) that K1 stages into shared memory.
//
// One WARP per spectrum, everything warp-synchronous in shared memory (no CTA barriers):
//   * order-defining steps are bitonic sorts over index arrays (stable through the index); an m/z-ascending input
//     (the normal MGF case) skips the m/z sort;
//   * the isotope chains (removeIsotopes / cleanIsotopes) restart wherever two neighbouring kept peaks are >= gap apart or
//     a peak lies below 200 Th, whatever happened before, so the chain segments between such hard boundaries run in
//     parallel, one lane per segment; only the TIC walk of MVH stays a serial FP64 prefix on one lane;
//   * the annotator's window predicate is inverted into exact break points in t (blob "ev"/"ans", pq_common.h).
#include <cstdlib>

#include "pq_internal.h"

namespace pq {

namespace {


__device__ __forceinline__ double next_up(double x) { return __longlong_as_double(__double_as_longlong(x) + 1); }      // positive finite x
__device__ __forceinline__ double next_down(double x) { return __longlong_as_double(__double_as_longlong(x) - 1); }

struct WarpMem {
    double* mz; double* in; double* da; double* db;            // raw m/z, raw intensity (file order), two work arrays
    uint16_t* ord; uint16_t* ordI; uint16_t* aux; uint16_t* eidx; uint16_t* cx;
    uint8_t* keep; uint8_t* k2;
};
__host__ __device__ inline size_t warp_smem_bytes(uint32_t cap) { return (size_t)cap * (8 * 4 + 2 * 5 + 2) + 64; }      // cap >= 128: the region is reused for the cell histogram (2*kCells bytes) + one bit per event once the events are written

__device__ __forceinline__ bool idx_less(const double* key, uint32_t a, uint32_t b) {
    double ka = key[a], kb = key[b];
    return ka < kb || (ka == kb && a < b);
}
// warp-synchronous bitonic sort of idx[0..n) by (key[idx], idx); n is a power of two >= 32
__device__ __noinline__ void warp_sort_idx(uint16_t* idx, const double* key, uint32_t n, int lane) {
    #pragma unroll 1
    for (uint32_t k = 2; k <= n; k <<= 1) {
        #pragma unroll 1
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = lane; i < n; i += 32) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    uint32_t a = idx[i], b = idx[ixj];
                    bool asc = (i & k) == 0;
                    if (idx_less(key, b, a) == asc) { idx[i] = (uint16_t)b; idx[ixj] = (uint16_t)a; }
                }
            }
            __syncwarp();
        }
    }
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
__device__ __forceinline__ double warp_max(double m) { for (int o = 16; o > 0; o >>= 1) { double x = __shfl_xor_sync(0xffffffffu, m, o); if (m < x) m = x; } return m; }

__device__ __noinline__ uint32_t count_kept(const uint8_t* keep, uint32_t P, int lane) {
    uint32_t c = 0;
    #pragma unroll 1
    for (uint32_t i = lane; i < P; i += 32) c += keep[i];
    return warp_sum(c);
}
__device__ void keep_all(uint8_t* keep, uint32_t P, int lane) { for (uint32_t i = lane; i < P; i += 32) keep[i] = 1; __syncwarp(); }
__device__ __noinline__ double max_kept_p(const uint8_t* keep, const double* in, const uint16_t* ord, uint32_t P, int lane) {
    double m = 0.0;                       // JSpectrum.getMaxIntensity starts at 0 (PSMMatch/JSpectrum.java:108-116)
    #pragma unroll 1
    for (uint32_t i = lane; i < P; i += 32) if (keep[i]) { double v = in[ord[i]]; if (m < v) m = v; }
    return warp_max(m);
}
__device__ __forceinline__ double max_kept(const WarpMem& W, uint32_t P, int lane) { return max_kept_p(W.keep, W.in, W.ord, P, lane); }

// JMatch.removeIsotopes / cleanIsotopes (PSMMatch/JMatch.java:223-280) over the kept entries, in m/z order.
// Reference chain: cur = first; for each next peak p: if p.mz - ref >= gap or p.mz < 200: emit cur, cur = p, ref = p.mz;
// else if p.I > cur.I: cur = p, ref = p.mz.  `ref` is the m/z of a kept peak at or before the previous one, so a peak
// that is >= gap above its kept predecessor (or below 200 Th) always restarts the chain: those are hard boundaries and
// the segments between them are independent.
__device__ __noinline__ void isotope_pass_p(const double* mz, const double* in, const uint16_t* ord, uint8_t* keep, uint8_t* k2, uint16_t* aux, uint32_t P, double gap, int lane) {
    if (count_kept(keep, P, lane) <= 2) return;
    // hard-boundary flags in aux (1 = chain restarts here), result in k2
    int carry = -1;                                                           // last kept position before this chunk
    #pragma unroll 1
    for (uint32_t c0 = 0; c0 < P; c0 += 32) {
        uint32_t i = c0 + lane;
        bool k = i < P && keep[i];
        uint32_t bal = __ballot_sync(0xffffffffu, k);
        if (i < P) {
            k2[i] = 0;
            uint32_t below = bal & ((1u << lane) - 1u);
            int prev = below ? (int)(c0 + 31 - __clz(below)) : carry;
            uint16_t hb = 0;
            if (k) {
                double m = mz[ord[i]];
                hb = (prev < 0 || m < 200.0 || __dsub_rn(m, mz[ord[prev]]) >= gap) ? 1 : 0;
            }
            aux[i] = hb;
        }
        if (bal) carry = (int)(c0 + 31 - __clz(bal));
    }
    __syncwarp();
    #pragma unroll 1
    for (uint32_t h = lane; h < P; h += 32) {
        if (!keep[h] || !aux[h]) continue;
        uint32_t cur = h; double ref = mz[ord[h]]; double cur_in = in[ord[h]];
        #pragma unroll 1
        for (uint32_t i = h + 1; i < P; ++i) {
            if (!keep[i]) continue;
            if (aux[i]) break;
            double m = mz[ord[i]];
            if (__dsub_rn(m, ref) >= gap) { k2[cur] = 1; cur = i; ref = m; cur_in = in[ord[i]]; }
            else { double v = in[ord[i]]; if (v > cur_in) { cur = i; ref = m; cur_in = v; } }
        }
        k2[cur] = 1;
    }
    __syncwarp();
    #pragma unroll 1
    for (uint32_t i = lane; i < P; i += 32) keep[i] = k2[i];
    __syncwarp();
}

__device__ __forceinline__ void isotope_pass(const WarpMem& W, uint32_t P, double gap, int lane) { isotope_pass_p(W.mz, W.in, W.ord, W.keep, W.k2, W.aux, P, gap, lane); }

// rank from the top among the kept entries by (intensity, m/z position): g = #kept j with I_j > I_i or (I_j == I_i and j > i).
// The kept intensities are first compacted (cv, positions in cp; m/z order survives), so the quadratic part runs over the kept
// entries only, one broadcast load per comparison partner, two ranked entries per lane and trip.
__device__ __noinline__ void top_rank_p(const uint8_t* keep, const double* in, const uint16_t* ord, uint32_t P, uint16_t* out, int lane,
                                        double* cv, uint16_t* cp) {
    uint32_t K = 0;
    #pragma unroll 1
    for (uint32_t c0 = 0; c0 < P; c0 += 32) {
        const uint32_t i = c0 + lane;
        const bool k = i < P && keep[i];
        const uint32_t bal = __ballot_sync(0xffffffffu, k);
        if (i < P) out[i] = 0;
        if (k) { const uint32_t pos = K + __popc(bal & ((1u << lane) - 1u)); cv[pos] = in[ord[i]]; cp[pos] = (uint16_t)i; }
        K += __popc(bal);
    }
    __syncwarp();
    #pragma unroll 1
    for (uint32_t k0 = lane; k0 < K; k0 += 64) {
        const uint32_t a = k0, b = k0 + 32;
        const bool hb = b < K;
        const double va = cv[a], vb = hb ? cv[b] : 0.0;
        uint32_t ga = 0, gb = 0;
        #pragma unroll 4
        for (uint32_t j = 0; j < K; ++j) {
            const double vj = cv[j];
            ga += (vj > va || (vj == va && j > a)) ? 1u : 0u;
            gb += (vj > vb || (vj == vb && j > b)) ? 1u : 0u;
        }
        out[cp[a]] = (uint16_t)ga;
        if (hb) out[cp[b]] = (uint16_t)gb;
    }
    __syncwarp();
}
__device__ __forceinline__ void top_rank(const WarpMem& W, uint32_t P, uint16_t* out, int lane) { top_rank_p(W.keep, W.in, W.ord, P, out, lane, W.da, W.cx); }      // da / cx are free whenever a rank is taken
// keep only the `top` largest kept entries: JMatch.filterByPeakCount (:148-174)
__device__ void keep_top(const WarpMem& W, uint32_t P, uint32_t top, int lane) {
    top_rank(W, P, W.aux, lane);
    #pragma unroll 1
    for (uint32_t i = lane; i < P; i += 32) if (W.keep[i] && W.aux[i] >= top) W.keep[i] = 0;
    __syncwarp();
}

__device__ __forceinline__ uint32_t count_le(const double* a, uint32_t n, double v) {      // #{a[i] <= v}, a ascending
    uint32_t lo = 0, hi = n;
    #pragma unroll 1
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (a[mid] <= v) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ uint32_t count_lt(const double* a, uint32_t n, double v) {      // #{a[i] < v}
    uint32_t lo = 0, hi = n;
    #pragma unroll 1
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

template <int kPrepWarps, int METHOD>
__global__ void __launch_bounds__(32 * kPrepWarps)
k0_prepare_kernel(SpectraView sp, const int32_t* __restrict__ list, int32_t n_list, PrepConfig cfg, uint8_t* __restrict__ blobs,
                  const uint64_t* __restrict__ blob_off, uint32_t* __restrict__ blob_len, uint32_t cap, uint32_t p_lo, uint32_t p_hi,
                  const int32_t* __restrict__ blob_index /* null: blob of list item i is blob i; else blob_index[spectrum] */,
                  uint32_t* __restrict__ work_counter /* null: items by static stride; else the next unclaimed item (zeroed by the launcher) */) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ double s_val[kPrepWarps][kValSlots];
    __shared__ double s_d[kPrepWarps][4];
    __shared__ uint32_t s_u[kPrepWarps][8];
    __shared__ int32_t s_cls[kPrepWarps][4];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* base_ptr = smem_raw + (size_t)warp * ((warp_smem_bytes(cap) + 15) & ~(size_t)15);
    WarpMem W;
    W.mz = reinterpret_cast<double*>(base_ptr); W.in = W.mz + cap; W.da = W.in + cap; W.db = W.da + cap;      // mz / in live to the end (lo / rem)
    W.ord = reinterpret_cast<uint16_t*>(W.db + cap); W.ordI = W.ord + cap; W.aux = W.ordI + cap; W.eidx = W.aux + cap; W.cx = W.eidx + cap;
    W.keep = reinterpret_cast<uint8_t*>(W.cx + cap); W.k2 = W.keep + cap;
    double* val = s_val[warp];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double itol = cfg.itol;

    const int32_t total_warps = (int32_t)gridDim.x * kPrepWarps;
    // Work is handed out warp by warp from a counter when there is one: spectra differ 10x in cost and a launch covers one upload
    // chunk, so with a static stride every launch ended on its unluckiest warp.
    auto next_item = [&](int32_t cur) -> int32_t {
        if (!work_counter) return cur < 0 ? (int32_t)blockIdx.x * kPrepWarps + warp : cur + total_warps;
        int32_t v = 0;
        if (lane == 0) v = (int32_t)atomicAdd(work_counter, 1u);
        return __shfl_sync(0xffffffffu, v, 0);
    };
    #pragma unroll 1
    for (int32_t item = next_item(-1); item < n_list; item = next_item(item)) {
        const int32_t s = list[item];
        const uint64_t base = sp.off[s];
        const uint32_t P = (uint32_t)(sp.off[s + 1] - base);
        if (P < p_lo || P > p_hi) continue;                        // the other launch's size class
        uint32_t n = 32; while (n < P) n <<= 1;
        const double prec_mz = sp.prec_mz[s];
        int z = sp.charge[s]; if (z == 0) z = 1;

        bool sorted = true, nonneg = true;
        #pragma unroll 1
        for (uint32_t i = lane; i < n; i += 32) {
            double m = i < P ? sp.mz[base + i] : inf;
            W.mz[i] = m;
            W.in[i] = i < P ? sp.inten[base + i] : inf;
            W.ord[i] = (uint16_t)i; W.ordI[i] = (uint16_t)i;
            if (i > 0 && i < P && !(sp.mz[base + i - 1] <= m)) sorted = false;
            if (i < P && !(W.in[i] >= 0.0)) nonneg = false;
        }
        sorted = __all_sync(0xffffffffu, sorted);
        nonneg = __all_sync(0xffffffffu, nonneg);
        __syncwarp();
        if (!sorted) warp_sort_idx(W.ord, W.mz, n, lane);          // stable m/z order (JSpectrum.sortPeaksByMZ)

        // annotator intensity limit: Spectrum.getIntensityLimit(percentile, 0.05) (JMatch.java:427-430; A2) needs the two
        // order statistics index, index+1 of the raw intensities.  MVH sorts by intensity anyway (the TIC walk needs the
        // order); hyperscore extracts the index+2 smallest values with warp minima instead of sorting.
        double limit;
        {
            const double indexDouble = P > 1 ? __dmul_rn(0.05, (double)(P - 1)) : 0.0;
            const int index = (int)indexDouble;
            double valueAtIndex, valueNext = 0.0;
            const bool extract = METHOD == 1 && P <= 256 && index + 2 <= 24 && nonneg;
            if (!extract) {
                warp_sort_idx(W.ordI, W.in, n, lane);              // stable intensity order (JSpectrum.sortPeaksByIntensity)
                #pragma unroll 1
                for (uint32_t i = lane; i < P; i += 32) W.da[i] = W.in[W.ordI[i]];
                __syncwarp();
                valueAtIndex = W.da[index];
                if (index + 1 < (int)P) valueNext = W.da[index + 1];
            } else {
                // Every lane sorts its own (up to eight) intensities in registers; a round then takes the warp minimum of the lanes'
                // heads — non-negative doubles order like their bit patterns, so two 32-bit warp reductions (high word, then low
                // word among the lanes that tie on it) find it — and the lane that held it moves on to its next value.  Ties may be
                // taken in any order: only the values matter.
                double v0 = inf, v1 = inf, v2 = inf, v3 = inf, v4 = inf, v5 = inf, v6 = inf, v7 = inf;
                if ((uint32_t)lane < P) v0 = W.in[lane];
                if ((uint32_t)lane + 32 < P) v1 = W.in[lane + 32];
                if ((uint32_t)lane + 64 < P) v2 = W.in[lane + 64];
                if ((uint32_t)lane + 96 < P) v3 = W.in[lane + 96];
#define K0_CE(a, b) { const double lo_ = fmin(a, b), hi_ = fmax(a, b); a = lo_; b = hi_; }
                if (P <= 128) { K0_CE(v0, v1) K0_CE(v2, v3) K0_CE(v0, v2) K0_CE(v1, v3) K0_CE(v1, v2) }
                else {
                    if ((uint32_t)lane + 128 < P) v4 = W.in[lane + 128];
                    if ((uint32_t)lane + 160 < P) v5 = W.in[lane + 160];
                    if ((uint32_t)lane + 192 < P) v6 = W.in[lane + 192];
                    if ((uint32_t)lane + 224 < P) v7 = W.in[lane + 224];
                    K0_CE(v0, v1) K0_CE(v2, v3) K0_CE(v4, v5) K0_CE(v6, v7) K0_CE(v0, v2) K0_CE(v1, v3) K0_CE(v4, v6) K0_CE(v5, v7)
                    K0_CE(v1, v2) K0_CE(v5, v6) K0_CE(v0, v4) K0_CE(v3, v7) K0_CE(v1, v5) K0_CE(v2, v6) K0_CE(v1, v4) K0_CE(v3, v6)
                    K0_CE(v2, v4) K0_CE(v3, v5) K0_CE(v3, v4)
                }
#undef K0_CE
                valueAtIndex = 0.0;
                #pragma unroll 1
                for (int r = 0; r <= index + 1 && r < (int)P; ++r) {
                    const uint32_t hi = (uint32_t)__double2hiint(v0);
                    const uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
                    const uint32_t lo = hi == mh ? (uint32_t)__double2loint(v0) : 0xFFFFFFFFu;
                    const uint32_t ml = __reduce_min_sync(0xffffffffu, lo);
                    const uint32_t ball = __ballot_sync(0xffffffffu, hi == mh && lo == ml);
                    if (lane == __ffs((int)ball) - 1) { v0 = v1; v1 = v2; v2 = v3; v3 = v4; v4 = v5; v5 = v6; v6 = v7; v7 = inf; }
                    const double mv = __hiloint2double((int)mh, (int)ml);
                    if (r == index) valueAtIndex = mv;
                    if (r == index + 1) valueNext = mv;
                }
            }
            if (P == 1) limit = valueAtIndex;
            else {
                double rest = __dsub_rn(indexDouble, (double)index);
                if (index == (int)P - 1 || rest == 0.0) limit = valueAtIndex;
                else limit = __dadd_rn(valueAtIndex, __dmul_rn(rest, __dsub_rn(valueNext, valueAtIndex)));
            }
        }
        bool has_dup = false;
        #pragma unroll 1
        for (uint32_t i = lane; i < P; i += 32) {
            W.keep[i] = 1;
            if (i > 0 && W.mz[W.ord[i]] == W.mz[W.ord[i - 1]]) has_dup = true;
        }
        has_dup = __any_sync(0xffffffffu, has_dup);
        __syncwarp();

        uint32_t nsurv = 0, flags = 0;
        if (METHOD == 1) {
            // ---- HyperscoreMatch.SpectrumPreProcessing (HyperscoreMatch.java:47-59) ----
            isotope_pass(W, P, 0.95, lane);
            const double tol_mz = __ddiv_rn(__dmul_rn(1.0, itol), (double)z);
            #pragma unroll 1
            for (uint32_t i = lane; i < P; i += 32)
                if (W.keep[i] && fabs(__dsub_rn(W.mz[W.ord[i]], prec_mz)) <= tol_mz) W.keep[i] = 0;          // removeParentPeak
            __syncwarp();
            if (count_kept(W.keep, P, lane) == 0) keep_all(W.keep, P, lane);
            #pragma unroll 1
            for (uint32_t i = lane; i < P; i += 32)
                if (W.keep[i] && W.mz[W.ord[i]] <= 150.0) W.keep[i] = 0;                                      // removeLowMasses
            __syncwarp();
            if (count_kept(W.keep, P, lane) == 0) {
                // getMaxIntensity() of an empty list is 0, then getPeaks() reloads the raw peaks (JSpectrum.java:71-76)
                keep_all(W.keep, P, lane);
            } else {
                double mx = max_kept(W, P, lane);
                double lim = __dmul_rn(__dmul_rn(1.0, 0.05), mx);
                #pragma unroll 1
                for (uint32_t i = lane; i < P; i += 32)
                    if (W.keep[i] && W.in[W.ord[i]] < lim) W.keep[i] = 0;                                     // removeLowIntensityPeak
                __syncwarp();
            }
            isotope_pass(W, P, 1.5, lane);                                                                    // cleanIsotopes
            const bool cut = count_kept(W.keep, P, lane) > 50;
            if (cut) keep_top(W, P, 50, lane);                                                                // filterByPeakCount
            const double mx = max_kept(W, P, lane);
            if (!has_dup) {
                // survivor id = position among the kept peaks; normalizePeaks(100)
                uint32_t run = 0;
                #pragma unroll 1
                for (uint32_t c0 = 0; c0 < P; c0 += 32) {
                    uint32_t i = c0 + lane;
                    bool k = i < P && W.keep[i];
                    uint32_t bal = __ballot_sync(0xffffffffu, k);
                    if (i < P) {
                        uint32_t sid = run + __popc(bal & ((1u << lane) - 1u));
                        W.aux[i] = k ? (uint16_t)sid : (uint16_t)kNoSurvivor;
                        if (k && sid < (uint32_t)kValSlots) val[sid] = __ddiv_rn(__dmul_rn(__dmul_rn(1.0, 100.0), W.in[W.ord[i]]), mx);
                    }
                    run += __popc(bal);
                }
                nsurv = run;
            } else {
                // Surviving duplicates of one m/z (possible below 200 Th) share a slot whose value is the LAST HashMap.put of
                // calcHyperScore's final m/z-sorted walk (:71-78): that sort is stable, so equal m/z keep the order of the
                // previous sort — by intensity when the top-50 cut ran, else file order.
                if (lane == 0) {
                    uint32_t sid = 0; bool have = false; double last = 0.0, last_in = 0.0;
                    #pragma unroll 1
                    for (uint32_t i = 0; i < P; ++i) {
                        W.aux[i] = (uint16_t)kNoSurvivor;
                        if (!W.keep[i]) continue;
                        double m = W.mz[W.ord[i]], v = W.in[W.ord[i]];
                        bool same = have && m == last;
                        if (!same) { if (have) ++sid; have = true; last = m; }
                        W.aux[i] = (uint16_t)sid;
                        if (!same || !cut || v >= last_in) { if (sid < (uint32_t)kValSlots) val[sid] = __ddiv_rn(__dmul_rn(__dmul_rn(1.0, 100.0), v), mx); last_in = v; }
                    }
                    uint32_t ns = have ? sid + 1 : 0;
                    // a peak that was filtered out but shares its m/z with a survivor still hits the survivor map
                    #pragma unroll 1
                    for (uint32_t i = 0; i < P; ++i) {
                        if (W.aux[i] != kNoSurvivor) continue;
                        double m = W.mz[W.ord[i]];
                        uint32_t found = kNoSurvivor;
                        #pragma unroll 1
                        for (uint32_t j = i; j-- > 0 && W.mz[W.ord[j]] == m;) if (W.aux[j] != kNoSurvivor) { found = W.aux[j]; break; }
                        if (found == kNoSurvivor)
                            #pragma unroll 1
                            for (uint32_t j = i + 1; j < P && W.mz[W.ord[j]] == m; ++j) if (W.aux[j] != kNoSurvivor) { found = W.aux[j]; break; }
                        W.aux[i] = (uint16_t)found;
                    }
                    s_u[warp][0] = ns;
                }
                __syncwarp();
                nsurv = s_u[warp][0];
            }
            __syncwarp();
        } else {
            // ---- MVHMatch.SpectrumPreProcessing + classifyPeakIntensities (MVHMatch.java:46-89) ----
            #pragma unroll 1
            for (uint32_t i = lane; i < P; i += 32) { double m = W.mz[W.ord[i]]; if (m < 200.0 || m > 2000.0) W.keep[i] = 0; }
            __syncwarp();
            if (count_kept(W.keep, P, lane) == 0) keep_all(W.keep, P, lane);
            const double tol_mz = __ddiv_rn(__dmul_rn(1.0, itol), (double)z);
            #pragma unroll 1
            for (uint32_t i = lane; i < P; i += 32)
                if (W.keep[i] && fabs(__dsub_rn(W.mz[W.ord[i]], prec_mz)) <= tol_mz) W.keep[i] = 0;
            __syncwarp();
            if (count_kept(W.keep, P, lane) == 0) keep_all(W.keep, P, lane);
            // filterByTIC (JMatch.java:341-364): aux[] = sorted position of each file index
            #pragma unroll 1
            for (uint32_t i = lane; i < P; i += 32) W.aux[W.ord[i]] = (uint16_t)i;
            __syncwarp();
            if (count_kept(W.keep, P, lane) > 7) {
                double tic = 0.0;
                if (lane == 0) for (uint32_t f = 0; f < P; ++f) if (W.keep[W.aux[f]]) tic = __dadd_rn(tic, W.in[f]);   // list order = file order
                tic = __shfl_sync(0xffffffffu, tic, 0);
                // the quotients I / TIC by every lane (the ~30-instruction divide was the bulk of the serial walk); lane 0 then adds
                // them in the reference's order (descending intensity), which is what decides the cut
                #pragma unroll 1
                for (uint32_t k = lane; k < P; k += 32) { const uint32_t f = W.ordI[k]; W.db[k] = W.keep[W.aux[f]] ? __ddiv_rn(W.in[f], tic) : 0.0; }
                __syncwarp();
                if (lane == 0) {
                    double rel = 0.0;
                    for (uint32_t k = P; k-- > 0;) {
                        uint32_t f = W.ordI[k]; uint32_t pos = W.aux[f];
                        if (!W.keep[pos]) continue;
                        rel = __dadd_rn(rel, W.db[k]);
                        if (!(rel <= 0.95)) W.keep[pos] = 0;
                    }
                }
                __syncwarp();
            }
            if (count_kept(W.keep, P, lane) == 0) keep_all(W.keep, P, lane);
            isotope_pass(W, P, 0.95, lane);
            if (count_kept(W.keep, P, lane) > 300) keep_top(W, P, 300, lane);
            uint32_t N = count_kept(W.keep, P, lane);
            if (N == 0) { keep_all(W.keep, P, lane); N = P; }
            // classes by descending (intensity, m/z position): top n -> 0, next 2n -> 1, rest -> 2
            top_rank(W, P, W.eidx, lane);
            const uint32_t npeak = N / 7;
            if (!has_dup) {
                uint32_t run = 0, c0n = 0, c1n = 0, c2n = 0;
                #pragma unroll 1
                for (uint32_t c0 = 0; c0 < P; c0 += 32) {
                    uint32_t i = c0 + lane;
                    bool k = i < P && W.keep[i];
                    uint32_t bal = __ballot_sync(0xffffffffu, k);
                    uint32_t c = 3;
                    if (k) { uint32_t g = W.eidx[i]; c = g < npeak ? 0u : (g < 3 * npeak ? 1u : 2u); }
                    if (i < P) {
                        uint32_t sid = run + __popc(bal & ((1u << lane) - 1u));
                        W.aux[i] = k ? (uint16_t)((c << 12) | sid) : (uint16_t)kNoSurvivor;
                    }
                    c0n += __popc(__ballot_sync(0xffffffffu, c == 0)); c1n += __popc(__ballot_sync(0xffffffffu, c == 1)); c2n += __popc(__ballot_sync(0xffffffffu, c == 2));
                    run += __popc(bal);
                }
                if (lane == 0) { s_cls[warp][0] = (int32_t)c0n; s_cls[warp][1] = (int32_t)c1n; s_cls[warp][2] = (int32_t)c2n; s_cls[warp][3] = cfg.mvh_bins - (int32_t)N; }
            } else if (lane == 0) {
                int32_t cc[4] = {0, 0, 0, 0};
                uint32_t sid = 0; bool have = false; double last = 0.0;
                #pragma unroll 1
                for (uint32_t i = 0; i < P; ++i) {
                    W.aux[i] = (uint16_t)kNoSurvivor;
                    if (!W.keep[i]) continue;
                    uint32_t g = W.eidx[i];
                    uint32_t c = g < npeak ? 0u : (g < 3 * npeak ? 1u : 2u);
                    cc[c]++;
                    double m = W.mz[W.ord[i]];
                    if (have && m == last) { } else { if (have) ++sid; have = true; last = m; }
                    W.aux[i] = (uint16_t)((c << 12) | sid);
                }
                #pragma unroll 1
                for (uint32_t i = 0; i < P; ++i) {
                    if (W.aux[i] != kNoSurvivor) continue;
                    double m = W.mz[W.ord[i]];
                    uint32_t found = kNoSurvivor;
                    #pragma unroll 1
                    for (uint32_t j = i; j-- > 0 && W.mz[W.ord[j]] == m;) if (W.aux[j] != kNoSurvivor) { found = W.aux[j]; break; }
                    if (found == kNoSurvivor)
                        #pragma unroll 1
                        for (uint32_t j = i + 1; j < P && W.mz[W.ord[j]] == m; ++j) if (W.aux[j] != kNoSurvivor) { found = W.aux[j]; break; }
                    W.aux[i] = (uint16_t)found;
                }
                cc[3] = cfg.mvh_bins - (int32_t)N;
                #pragma unroll 1
                for (int c = 0; c < 4; ++c) s_cls[warp][c] = cc[c];
            }
            __syncwarp();
            nsurv = N;
            flags = N >= 7 ? 1u : 0u;
        }

        // ---- matchable peaks (intensity >= limit) in m/z order: compact m/z -> da, intensity -> db, claim code -> cx ----
        uint32_t E = 0;
        #pragma unroll 1
        for (uint32_t c0 = 0; c0 < P; c0 += 32) {
            uint32_t i = c0 + lane;
            bool el = i < P && W.in[W.ord[i]] >= limit;
            uint32_t bal = __ballot_sync(0xffffffffu, el);
            if (i < P) W.eidx[i] = el ? (uint16_t)(E + __popc(bal & ((1u << lane) - 1u))) : (uint16_t)0xFFFFu;
            E += __popc(bal);
        }
        __syncwarp();
        const BlobLayout BL = blob_layout(E);
        const int32_t bi = blob_index ? blob_index[s] : item;
        uint8_t* blob = blobs + blob_off[bi];
        double* g_ev = reinterpret_cast<double*>(blob + BL.ev);
        uint16_t* g_ans = reinterpret_cast<uint16_t*>(blob + BL.ans);
        double* g_pmz = reinterpret_cast<double*>(blob + BL.pmz);
        double* g_pin = reinterpret_cast<double*>(blob + BL.pin);
        uint16_t* g_pcode = reinterpret_cast<uint16_t*>(blob + BL.pcode);
        double* g_val = reinterpret_cast<double*>(blob + BL.val);
        uint16_t* g_cells = reinterpret_cast<uint16_t*>(blob + BL.cells);
        #pragma unroll 1
        for (uint32_t i = lane; i < P; i += 32) {
            uint32_t e = W.eidx[i];
            if (e == 0xFFFFu) continue;
            uint32_t f = W.ord[i];
            double m = W.mz[f], v = W.in[f];
            uint32_t code = W.aux[i];
            W.da[e] = m; W.db[e] = v; W.cx[e] = (uint16_t)code;
            g_pmz[e] = m; g_pin[e] = v; g_pcode[e] = (uint16_t)code;
        }
        if (lane == 0) { g_pmz[E] = inf; if (((E + 1u) & 1u) != 0u) g_pmz[E + 1] = inf; }
        __syncwarp();
        // ---- exact break points of the window predicate fabs(fl(mz - t)) <= itol: enter at lo, leave at rem = nextup(hi) ----
        // (the raw arrays are dead now: lo -> mz, rem -> in)
        double* lo_a = W.mz; double* rem_a = W.in;
        #pragma unroll 1
        for (uint32_t e = lane; e < E; e += 32) {
            const double m = W.da[e];
            double g = __dsub_rn(m, itol);
            double lo;
            if (!(g > 0.0)) lo = 0.0;
            else {
                #pragma unroll 1
                while (!(__dsub_rn(m, g) <= itol)) g = next_up(g);
                #pragma unroll 1
                for (;;) { double q = next_down(g); if (q > 0.0 && __dsub_rn(m, q) <= itol) g = q; else break; }
                lo = g;
            }
            double h = __dadd_rn(m, itol);
            #pragma unroll 1
            while (!(__dsub_rn(h, m) <= itol)) h = next_down(h);
            #pragma unroll 1
            for (;;) { double q = next_up(h); if (__dsub_rn(q, m) <= itol) h = q; else break; }
            lo_a[e] = lo; rem_a[e] = next_up(h);
        }
        __syncwarp();
        // merged, sorted event list (a leave event goes before an enter event of the same value) and its answers.
        // lo[] and rem[] are the m/z list shifted by -+itol, so the counts below are a short walk from the peak itself.
        const uint32_t NE = 2u * E + 2u;
        #pragma unroll 1
        for (uint32_t q = lane; q < 2u * E; q += 32) {
            const bool is_add = q < E;
            const uint32_t e = is_add ? q : q - E;
            const double v = is_add ? lo_a[e] : rem_a[e];
            // a = #{rem <= v}: peaks that have left; bend = #{lo <= v}: peaks that have entered
            uint32_t a, bend, pos;
            if (is_add) {
                a = e; while (a > 0 && rem_a[a - 1] > v) --a;                 // rem[e] > lo[e]: walk down from e
                bend = e + 1; while (bend < E && lo_a[bend] <= v) ++bend;     // equal lo above e
                pos = 1u + e + a;
            } else {
                a = e + 1; while (a < E && rem_a[a] <= v) ++a;               // equal rem above e
                bend = e + 1; while (bend < E && lo_a[bend] <= v) ++bend;     // lo[e] < rem[e]: walk up from e
                uint32_t lt = bend; while (lt > 0 && !(lo_a[lt - 1] < v)) --lt;   // #{lo < v}
                pos = 1u + e + lt;
            }
            // window for ev[pos] <= t < ev[pos+1] = peaks [a, bend): the most intense wins, equal intensities tie
            uint32_t code = kAnsNone;
            if (a < bend) {
                uint32_t best = a; double best_v = W.db[a]; bool tie = false;
                #pragma unroll 1
                for (uint32_t j = a + 1; j < bend; ++j) {
                    double r = W.db[j];
                    if (r > best_v) { best = j; best_v = r; tie = false; }
                    else if (r == best_v) tie = true;
                }
                if (tie) code = kAnsTie | a;
                else {
                    uint32_t cc = W.cx[best];
                    if ((cc & 0xFFFu) == kNoSurvivor) code = kAnsNone;
                    else if (METHOD == 2 && (W.da[best] < 200.0 || W.da[best] > 2000.0)) code = kAnsNone;     // MVHMatch.java:248-250
                    else code = cc;
                }
            }
            g_ev[pos] = v; g_ans[pos] = (uint16_t)code;
        }
        if (lane == 0) { g_ev[0] = -inf; g_ans[0] = (uint16_t)kAnsNone; }
        // ---- merge neighbouring intervals with the same answer (a peak entering or leaving under a stronger one, windows of
        // peaks that did not survive pre-processing): only the change points stay, about half of the 2E events.  In-place
        // left compaction, 32 events at a time (every lane reads its event before any lane writes).  The cell of every
        // kept event is noted for the cell table below (shared memory the dead work arrays leave behind).
        // (Deciding "kept" while the events are created — comparing the window just below each event with the window at it —
        // avoids this read-back but measured slower: 17.2 vs 14.5 ms for K0.)
        __syncwarp();
        // (every work array is dead once the events are in global memory: the cell table is assembled at the start of the warp's
        // region, the cell of every kept event — top bit: the event answers "nothing" — behind it)
        uint16_t* cells_s = reinterpret_cast<uint16_t*>(W.mz);                       // kCells entries, copied out at the end
        uint16_t* ecell = cells_s + kCells + 8;                                      // [kept]
        if (lane == 0) ecell[0] = (uint16_t)kCellClean;                              // the -inf sentinel: below every cell, answers "nothing"
        uint32_t kept = 1;                                                           // slot 0 = the -inf sentinel
        {
            uint32_t prev_code = kAnsNone;
            #pragma unroll 1
            for (uint32_t c0 = 1; c0 < NE - 1; c0 += 32) {
                const uint32_t p = c0 + lane;
                const bool in = p < NE - 1;
                const double v = in ? __ldcg(g_ev + p) : 0.0;
                const uint32_t code = in ? (uint32_t)__ldcg(g_ans + p) : 0u;
                uint32_t left = __shfl_up_sync(0xffffffffu, code, 1);
                if (lane == 0) left = prev_code;
                const bool keep = in && code != left;
                const uint32_t bal = __ballot_sync(0xffffffffu, keep);
                const uint32_t n_in = min(32u, NE - 1 - c0);
                prev_code = __shfl_sync(0xffffffffu, code, (int)n_in - 1);
                __syncwarp();
                if (keep) {
                    const uint32_t idx = kept + __popc(bal & ((1u << lane) - 1u));
                    g_ev[idx] = v; g_ans[idx] = (uint16_t)code;
                    ecell[idx] = (uint16_t)((uint32_t)cell_of(v) | (code == kAnsNone ? kCellClean : 0u));
                }
                kept += __popc(bal);
                __syncwarp();
            }
        }
        const uint32_t NE2 = kept + 1;                                               // + the +inf sentinel
        if (lane == 0) { g_ev[kept] = inf; g_ans[kept] = (uint16_t)kAnsNone; }
        #pragma unroll 1
        for (uint32_t k = NE2 + lane; k < ((NE2 + 7u) & ~7u); k += 32) g_ans[k] = (uint16_t)kAnsNone;
        // ---- cell table: cells[c] = number of real events below the lower edge of cell c (= index of the last of them), i.e.
        // kept event k owns the cells above its own up to and including the cell of event k + 1 (the last event: up to the end of
        // the table).  kCellClean: no event inside the cell and nothing to claim at its lower edge, i.e. every t of the cell
        // answers "nothing" — all owned cells but the one event k + 1 lies in, when event k answers "nothing": K1 skips the event
        // scan for such cells (most of the m/z axis).  One lane per event; the ranges tile the table, nothing needs clearing.
        __syncwarp();
        #pragma unroll 1
        for (uint32_t k0 = 0; k0 < kept; k0 += 32) {
            const uint32_t k = k0 + lane;
            uint32_t first_c = 1, end_c = 0, vb = 0, ve = 0, last = 0;
            if (k < kept) {
                const uint32_t ec = ecell[k];
                last = k + 1 >= kept ? 1u : 0u;
                first_c = k == 0 ? 0u : (ec & 0x7FFFu) + 1u;
                end_c = last ? (uint32_t)kCells - 1u : (uint32_t)(ecell[k + 1] & 0x7FFFu);          // inclusive
                vb = k | (ec & kCellClean); ve = k;
            }
            const bool wide = end_c + 1u > first_c + 64u;              // a long stretch of the axis without events: the whole warp fills it
            if (!wide && first_c <= end_c) {
                // [first_c, eb) holds vb, two cells per store once aligned; cell end_c holds ve unless the stretch ends the table
                const uint32_t eb = last ? end_c + 1u : end_c;
                uint32_t c = first_c;
                if ((c & 1u) && c < eb) cells_s[c++] = (uint16_t)vb;
                const uint32_t vv = vb | (vb << 16);
                #pragma unroll 1
                for (; c + 2 <= eb; c += 2) *reinterpret_cast<uint32_t*>(cells_s + c) = vv;
                if (c < eb) cells_s[c] = (uint16_t)vb;
                if (!last) cells_s[end_c] = (uint16_t)ve;
            }
            uint32_t todo = __ballot_sync(0xffffffffu, wide);
            #pragma unroll 1
            while (todo) {
                const int src = __ffs((int)todo) - 1;
                todo &= todo - 1;
                const uint32_t f = __shfl_sync(0xffffffffu, first_c, src), e = __shfl_sync(0xffffffffu, end_c, src);
                const uint32_t b = __shfl_sync(0xffffffffu, vb, src), v = __shfl_sync(0xffffffffu, ve, src), l = __shfl_sync(0xffffffffu, last, src);
                // [f, eb) holds b, two cells per lane and trip; cell e holds v unless the stretch ends the table
                const uint32_t eb = l ? e + 1u : e, c0 = f + (f & 1u);
                const uint32_t bb = b | (b << 16);
                for (uint32_t c = c0 + 2u * lane; c + 2u <= eb; c += 64u) *reinterpret_cast<uint32_t*>(cells_s + c) = bb;
                if (lane == 0) {
                    if ((f & 1u) && f < eb) cells_s[f] = (uint16_t)b;
                    if (eb > c0 && ((eb - c0) & 1u)) cells_s[eb - 1u] = (uint16_t)b;
                    if (!l) cells_s[e] = (uint16_t)v;
                }
            }
        }
        __syncwarp();
        // cell_filter() sends every m/z outside the table to the last cell: it may only stay clean if cell 0 (all events below
        // 128 Th) is clean too
        if (lane == 0) {
            const uint16_t first = cells_s[0], last = cells_s[kCells - 1];
            if (!(first & kCellClean) && (last & kCellClean)) cells_s[kCells - 1] = (uint16_t)(last & ~kCellClean);
        }
        __syncwarp();
        // copy-out, 16 bytes per lane and trip (coalesced), and the clean flags once more as a bit map (pq_common.h: cbits)
        {
            static_assert(kCells % (32 * 8) == 0, "cell table must split into uint4 chunks per lane");
            uint32_t* g_bits = reinterpret_cast<uint32_t*>(blob + BL.cbits);
            #pragma unroll 2
            for (uint32_t k = 0; k < (uint32_t)kCells / 256; ++k) {
                const uint32_t q = lane + 32 * k;
                const uint4 v = reinterpret_cast<const uint4*>(cells_s)[q];
                reinterpret_cast<uint4*>(g_cells)[q] = v;
                const uint32_t b8 = ((v.x >> 15) & 1u) | ((v.x >> 30) & 2u) | (((v.y >> 15) & 1u) << 2) | (((v.y >> 30) & 2u) << 2) |
                                    (((v.z >> 15) & 1u) << 4) | (((v.z >> 30) & 2u) << 4) | (((v.w >> 15) & 1u) << 6) | (((v.w >> 30) & 2u) << 6);
                const uint32_t w = b8 | (__shfl_down_sync(0xffffffffu, b8, 1) << 8) | (__shfl_down_sync(0xffffffffu, b8, 2) << 16) | (__shfl_down_sync(0xffffffffu, b8, 3) << 24);
                if ((lane & 3) == 0) g_bits[q >> 2] = w;
            }
        }
        if (lane == 1) *reinterpret_cast<uint4*>(g_cells + kCells) = make_uint4(0x80008000u, 0x80008000u, 0x80008000u, 0x80008000u);      // the sentinel: always clean
        if (METHOD == 1) for (uint32_t k = lane; k < (uint32_t)kValSlots; k += 32) g_val[k] = k < nsurv ? val[k] : 0.0;
        else for (uint32_t k = lane; k < (uint32_t)kValSlots; k += 32) g_val[k] = 0.0;
        if (lane == 0) {
            BlobHeader h;
            h.n_peaks = E; h.n_events = NE2; h.n_survivors = nsurv; h.flags = flags;
            #pragma unroll 1
            for (int c = 0; c < 4; ++c) h.cls_count[c] = METHOD == 2 ? s_cls[warp][c] : 0;
            h.total_bytes = BL.total; h.raw_peaks = P; h.pad[0] = h.pad[1] = 0;
            *reinterpret_cast<BlobHeader*>(blob) = h;
            blob_len[bi] = BL.total;
        }
        __syncwarp();
    }
}

}  // namespace

int launch_prepare(const SpectraView& sp, const int32_t* list, int32_t n_list, const PrepConfig& cfg, uint8_t* blobs,
                   const uint64_t* blob_off, uint32_t* blob_len, uint32_t max_peaks, cudaStream_t st, uint32_t min_peaks, const int32_t* blob_index,
                   int leave_blocks_per_sm, uint32_t* work_counters /* one zeroed word per kernel this call may launch (4), or null */) {
    if (n_list <= 0) return 0;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int launched = 0;
    auto run = [&](auto kernel, int warps, uint32_t cap, uint32_t p_lo, uint32_t p_hi) {
        size_t per_warp = (warp_smem_bytes(cap) + 15) & ~(size_t)15;
        size_t smem = per_warp * (size_t)warps;
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per_sm = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * warps, smem);
        per_sm -= leave_blocks_per_sm;               // room for a kernel of another stream (the candidate-first gather) beside this persistent grid
        if (per_sm < 1) per_sm = 1;
        int grid = sms * per_sm; int need = (n_list + warps - 1) / warps; if (grid > need) grid = need;
        kernel<<<grid, 32 * warps, smem, st>>>(sp, list, n_list, cfg, blobs, blob_off, blob_len, cap, p_lo, p_hi, blob_index,
                                               work_counters ? work_counters + launched : nullptr);
        ++launched;
    };
    // size class 1: up to 256 peaks (the common case), four spectra per CTA; size class 2: the rest, capacity = next power of two
    // size classes by peak count: K0 is bound by shared-memory capacity (46 bytes per peak slot and warp), so spectra of up
    // to 128 peaks run with half the footprint and twice the resident warps
    const bool mvh = cfg.method == 2;
    const bool split = getenv("PQ_K0_ONE_CLASS") == nullptr;
    if (split) {
        if (mvh) run(k0_prepare_kernel<4, 2>, 4, 128, min_peaks, 128); else run(k0_prepare_kernel<4, 1>, 4, 128, min_peaks, 128);
        if (max_peaks > 128) { if (mvh) run(k0_prepare_kernel<4, 2>, 4, 256, 129, 256); else run(k0_prepare_kernel<4, 1>, 4, 256, 129, 256); }
    } else {
        if (mvh) run(k0_prepare_kernel<4, 2>, 4, 256, min_peaks, 256); else run(k0_prepare_kernel<4, 1>, 4, 256, min_peaks, 256);
    }
    if (max_peaks > 256) {
        uint32_t cap = 512; while (cap < max_peaks) cap <<= 1;
        if (cap <= 1024) { if (mvh) run(k0_prepare_kernel<4, 2>, 4, cap, 257, 0xFFFFFFFFu); else run(k0_prepare_kernel<4, 1>, 4, cap, 257, 0xFFFFFFFFu); }
        else { if (mvh) run(k0_prepare_kernel<1, 2>, 1, cap, 257, 0xFFFFFFFFu); else run(k0_prepare_kernel<1, 1>, 1, cap, 257, 0xFFFFFFFFu); }
    }
    return launched;
}

}  // namespace pq

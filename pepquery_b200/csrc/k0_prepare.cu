// k0_prepare.cu — K0: peptide-independent spectrum pre-processing, hoisted out of the pair loop.
//
// The reference redoes this for every (peptide, spectrum) pair inside scorePeptide2Spectrum
// (pg/PeptideSearchMT.java:1144-1251): HyperscoreMatch.SpectrumPreProcessing (PSMMatch/HyperscoreMatch.java:47-59),
// MVHMatch.SpectrumPreProcessing + classifyPeakIntensities (PSMMatch/MVHMatch.java:46-89) over the JMatch filters
// (PSMMatch/JMatch.java:101-292,341-383), plus the annotator's intensity threshold (JMatch.java:427-430).
// It depends only on the spectrum, so here it runs once per spectrum that has at least one candidate and writes one
// "prepared spectrum" blob (pq_common.h) that K1 stages into shared memory.
//
// One CTA per spectrum.  Peaks live in shared memory; order-defining steps use a bitonic sort over index arrays
// (stable through the file index), the two chain-dependent isotope passes and the TIC walk are sequential on one
// thread because each decision depends on the previous one.
#include "pq_internal.h"

namespace pq {

namespace {

constexpr int kPrepBlock = 128;

struct PrepSmem {
    double* r_mz; double* r_in; double* w_d;
    uint32_t* ord; uint32_t* ordI; uint32_t* aux;
    uint8_t* keep; uint16_t* cells;
};
__host__ __device__ inline size_t prep_smem_bytes(uint32_t npad) {
    return (size_t)npad * (8 * 3 + 4 * 3 + 1) + 16 + kCells * 2 + 64;
}

__device__ __forceinline__ bool key_less(const double* key, uint32_t a, uint32_t b) {
    double ka = key[a], kb = key[b];
    return ka < kb || (ka == kb && a < b);
}

__device__ void bitonic_sort_idx(uint32_t* idx, const double* key, uint32_t n) {
    for (uint32_t k = 2; k <= n; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                uint32_t ixj = i ^ j;
                if (ixj > i) {
                    uint32_t a = idx[i], b = idx[ixj];
                    bool asc = (i & k) == 0;
                    if (key_less(key, b, a) == asc) { idx[i] = b; idx[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// JMatch.removeIsotopes / cleanIsotopes (PSMMatch/JMatch.java:223-280) over the kept entries, in m/z order.
// Runs on one thread: each step depends on the current chain head.
__device__ void isotope_pass(const double* r_mz, const double* r_in, const uint32_t* ord, uint8_t* keep, uint32_t P, double gap) {
    uint32_t cnt = 0;
    for (uint32_t i = 0; i < P; ++i) cnt += keep[i];
    if (cnt <= 2) return;
    int cur = -1; double ref = 0.0;
    for (uint32_t i = 0; i < P; ++i) {
        if (!keep[i]) continue;
        keep[i] = 0;
        double m = r_mz[ord[i]];
        if (cur < 0) { cur = (int)i; ref = m; continue; }
        if (__dsub_rn(m, ref) >= gap || m < 200.0) { keep[cur] = 1; cur = (int)i; ref = m; }
        else if (r_in[ord[i]] > r_in[ord[cur]]) { cur = (int)i; ref = m; }
    }
    if (cur >= 0) keep[cur] = 1;
}

__device__ uint32_t block_count(const uint8_t* keep, uint32_t P, uint32_t* s_tmp) {
    uint32_t c = 0;
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) c += keep[i];
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = c;
    __syncthreads();
    uint32_t t = 0;
    for (uint32_t w = 0; w < blockDim.x / 32; ++w) t += s_tmp[w];
    __syncthreads();
    return t;
}
__device__ double block_max_kept(const double* r_in, const uint32_t* ord, const uint8_t* keep, uint32_t P, double* s_tmp) {
    double m = 0.0;                       // JSpectrum.getMaxIntensity starts at 0 (PSMMatch/JSpectrum.java:108-116)
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) if (keep[i]) { double v = r_in[ord[i]]; if (m < v) m = v; }
    for (int o = 16; o > 0; o >>= 1) { double x = __shfl_xor_sync(0xffffffffu, m, o); if (m < x) m = x; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_tmp[threadIdx.x >> 5] = m;
    __syncthreads();
    double t = 0.0;
    for (uint32_t w = 0; w < blockDim.x / 32; ++w) if (t < s_tmp[w]) t = s_tmp[w];
    __syncthreads();
    return t;
}
// keep only the `top` largest kept entries by (intensity, m/z position): JMatch.filterByPeakCount (:148-174)
__device__ void keep_top(const double* r_in, const uint32_t* ord, uint8_t* keep, uint32_t* aux, uint32_t P, uint32_t top) {
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
        uint32_t g = 0;
        if (keep[i]) {
            double vi = r_in[ord[i]];
            for (uint32_t j = 0; j < P; ++j) {
                if (!keep[j]) continue;
                double vj = r_in[ord[j]];
                g += (vj > vi || (vj == vi && j > i)) ? 1u : 0u;
            }
        }
        aux[i] = g;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) if (keep[i] && aux[i] >= top) keep[i] = 0;
    __syncthreads();
}

__global__ void __launch_bounds__(kPrepBlock)
k0_prepare_kernel(SpectraView sp, const int32_t* __restrict__ list, int32_t n_list, PrepConfig cfg,
                  uint8_t* __restrict__ blobs, const uint64_t* __restrict__ blob_off, uint32_t npad_max) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ uint32_t s_u[8];
    __shared__ double s_d[8];
    __shared__ double s_limit, s_tic;
    __shared__ uint32_t s_E, s_nsurv;
    __shared__ int32_t s_cls[4];
    __shared__ uint32_t s_flags;
    __shared__ double s_val[kValSlots];

    double* r_mz = reinterpret_cast<double*>(smem_raw);
    double* r_in = r_mz + npad_max;
    double* w_d = r_in + npad_max;
    uint32_t* ord = reinterpret_cast<uint32_t*>(w_d + npad_max);
    uint32_t* ordI = ord + npad_max;
    uint32_t* aux = ordI + npad_max;
    uint8_t* keep = reinterpret_cast<uint8_t*>(aux + npad_max);
    uint16_t* cells = reinterpret_cast<uint16_t*>(smem_raw + (((size_t)npad_max * 37 + 15) & ~(size_t)15));

    const double inf = __longlong_as_double(0x7ff0000000000000LL);

    for (int32_t item = blockIdx.x; item < n_list; item += gridDim.x) {
        const int32_t s = list[item];
        const uint64_t base = sp.off[s];
        const uint32_t P = (uint32_t)(sp.off[s + 1] - base);
        uint32_t n = 2; while (n < P) n <<= 1;
        const double prec_mz = sp.prec_mz[s];
        int z = sp.charge[s]; if (z == 0) z = 1;

        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            r_mz[i] = i < P ? sp.mz[base + i] : inf;
            r_in[i] = i < P ? sp.inten[base + i] : inf;
            ord[i] = i; ordI[i] = i;
        }
        __syncthreads();
        bitonic_sort_idx(ord, r_mz, n);       // stable m/z order (JSpectrum.sortPeaksByMZ)
        bitonic_sort_idx(ordI, r_in, n);      // stable intensity order (JSpectrum.sortPeaksByIntensity)
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) w_d[i] = r_in[ordI[i]];
        __syncthreads();

        // annotator intensity limit: Spectrum.getIntensityLimit(percentile, 0.05) (JMatch.java:427-430; A2)
        if (threadIdx.x == 0) {
            double lim;
            if (P == 1) lim = w_d[0];
            else {
                double indexDouble = __dmul_rn(0.05, (double)(P - 1));
                int index = (int)indexDouble;
                double valueAtIndex = w_d[index];
                double rest = __dsub_rn(indexDouble, (double)index);
                if (index == (int)P - 1 || rest == 0.0) lim = valueAtIndex;
                else lim = __dadd_rn(valueAtIndex, __dmul_rn(rest, __dsub_rn(w_d[index + 1], valueAtIndex)));
            }
            s_limit = lim;
            s_flags = 0;
        }
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) keep[i] = 1;
        __syncthreads();
        const double limit = s_limit;

        uint32_t nsurv = 0;
        if (cfg.method == 1) {
            // ---- HyperscoreMatch.SpectrumPreProcessing (HyperscoreMatch.java:47-59) ----
            if (threadIdx.x == 0) isotope_pass(r_mz, r_in, ord, keep, P, 0.95);
            __syncthreads();
            const double tol_mz = __ddiv_rn(__dmul_rn(1.0, cfg.itol), (double)z);
            for (uint32_t i = threadIdx.x; i < P; i += blockDim.x)
                if (keep[i] && fabs(__dsub_rn(r_mz[ord[i]], prec_mz)) <= tol_mz) keep[i] = 0;       // removeParentPeak
            __syncthreads();
            if (block_count(keep, P, s_u) == 0) { for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) keep[i] = 1; __syncthreads(); }
            for (uint32_t i = threadIdx.x; i < P; i += blockDim.x)
                if (keep[i] && r_mz[ord[i]] <= 150.0) keep[i] = 0;                                   // removeLowMasses
            __syncthreads();
            if (block_count(keep, P, s_u) == 0) {
                // getMaxIntensity() of an empty list is 0, then getPeaks() reloads the raw peaks (JSpectrum.java:71-76)
                for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) keep[i] = 1;
                __syncthreads();
            } else {
                double mx = block_max_kept(r_in, ord, keep, P, s_d);
                double lim = __dmul_rn(__dmul_rn(1.0, 0.05), mx);
                for (uint32_t i = threadIdx.x; i < P; i += blockDim.x)
                    if (keep[i] && r_in[ord[i]] < lim) keep[i] = 0;                                  // removeLowIntensityPeak
                __syncthreads();
            }
            if (threadIdx.x == 0) isotope_pass(r_mz, r_in, ord, keep, P, 1.5);                       // cleanIsotopes
            __syncthreads();
            const bool cut = block_count(keep, P, s_u) > 50;
            if (cut) keep_top(r_in, ord, keep, aux, P, 50);                                          // filterByPeakCount
            double mx = block_max_kept(r_in, ord, keep, P, s_d);
            // survivor ids by distinct m/z.  Surviving duplicates of one m/z (possible below 200 Th) share a slot whose
            // value is the LAST HashMap.put of calcHyperScore's final m/z-sorted walk (:71-78): that sort is stable, so
            // equal m/z keep the order of the previous sort — by intensity when the top-50 cut ran, else file order.
            if (threadIdx.x == 0) {
                uint32_t sid = 0; bool have = false; double last = 0.0, last_in = 0.0;
                double* val = s_val;
                for (uint32_t i = 0; i < P; ++i) {
                    aux[i] = kNoSurvivor;
                    if (!keep[i]) continue;
                    double m = r_mz[ord[i]], v = r_in[ord[i]];
                    bool same = have && m == last;
                    if (!same) { if (have) ++sid; have = true; last = m; }
                    aux[i] = sid;
                    if (!same || !cut || v >= last_in) { val[sid] = __ddiv_rn(__dmul_rn(__dmul_rn(1.0, 100.0), v), mx); last_in = v; }   // normalizePeaks(100)
                }
                uint32_t ns = have ? sid + 1 : 0;
                // a peak that was filtered out but shares its m/z with a survivor still hits the survivor map
                for (uint32_t i = 0; i < P; ++i) {
                    if (aux[i] != kNoSurvivor) continue;
                    double m = r_mz[ord[i]];
                    uint32_t found = kNoSurvivor;
                    for (uint32_t j = i; j-- > 0 && r_mz[ord[j]] == m;) if (aux[j] != kNoSurvivor) { found = aux[j]; break; }
                    if (found == kNoSurvivor)
                        for (uint32_t j = i + 1; j < P && r_mz[ord[j]] == m; ++j) if (aux[j] != kNoSurvivor) { found = aux[j]; break; }
                    aux[i] = found;
                }
                s_nsurv = ns;
            }
            __syncthreads();
            nsurv = s_nsurv;
        } else {
            // ---- MVHMatch.SpectrumPreProcessing + classifyPeakIntensities (MVHMatch.java:46-89) ----
            for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) { double m = r_mz[ord[i]]; if (m < 200.0 || m > 2000.0) keep[i] = 0; }
            __syncthreads();
            if (block_count(keep, P, s_u) == 0) { for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) keep[i] = 1; __syncthreads(); }
            const double tol_mz = __ddiv_rn(__dmul_rn(1.0, cfg.itol), (double)z);
            for (uint32_t i = threadIdx.x; i < P; i += blockDim.x)
                if (keep[i] && fabs(__dsub_rn(r_mz[ord[i]], prec_mz)) <= tol_mz) keep[i] = 0;
            __syncthreads();
            if (block_count(keep, P, s_u) == 0) { for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) keep[i] = 1; __syncthreads(); }
            // filterByTIC (JMatch.java:341-364): aux[] = sorted position of each file index
            for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) aux[ord[i]] = i;
            __syncthreads();
            if (block_count(keep, P, s_u) > 7) {
                if (threadIdx.x == 0) {
                    double tic = 0.0;
                    for (uint32_t f = 0; f < P; ++f) if (keep[aux[f]]) tic = __dadd_rn(tic, r_in[f]);   // list order = file order
                    double rel = 0.0;
                    for (uint32_t k = P; k-- > 0;) {
                        uint32_t f = ordI[k]; uint32_t pos = aux[f];
                        if (!keep[pos]) continue;
                        rel = __dadd_rn(rel, __ddiv_rn(r_in[f], tic));
                        if (!(rel <= 0.95)) keep[pos] = 0;
                    }
                }
                __syncthreads();
            }
            if (block_count(keep, P, s_u) == 0) { for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) keep[i] = 1; __syncthreads(); }
            if (threadIdx.x == 0) isotope_pass(r_mz, r_in, ord, keep, P, 0.95);
            __syncthreads();
            if (block_count(keep, P, s_u) > 300) keep_top(r_in, ord, keep, aux, P, 300);
            uint32_t N = block_count(keep, P, s_u);
            if (N == 0) { for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) keep[i] = 1; __syncthreads(); N = P; }
            // classes by descending (intensity, m/z position): top n -> 0, next 2n -> 1, rest -> 2
            for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
                uint32_t g = 0;
                if (keep[i]) {
                    double vi = r_in[ord[i]];
                    for (uint32_t j = 0; j < P; ++j) { if (!keep[j]) continue; double vj = r_in[ord[j]]; g += (vj > vi || (vj == vi && j > i)) ? 1u : 0u; }
                }
                ordI[i] = g;      // ordI is free from here on: rank from the top
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t npeak = N / 7;
                int32_t cc[4] = {0, 0, 0, 0};
                uint32_t sid = 0; bool have = false; double last = 0.0;
                for (uint32_t i = 0; i < P; ++i) {
                    aux[i] = kNoSurvivor;
                    if (!keep[i]) continue;
                    uint32_t g = ordI[i];
                    uint32_t c = g < npeak ? 0u : (g < 3 * npeak ? 1u : 2u);
                    cc[c]++;
                    double m = r_mz[ord[i]];
                    if (have && m == last) { } else { if (have) ++sid; have = true; last = m; }
                    aux[i] = (c << 12) | sid;
                }
                for (uint32_t i = 0; i < P; ++i) {
                    if (aux[i] != kNoSurvivor) continue;
                    double m = r_mz[ord[i]];
                    uint32_t found = kNoSurvivor;
                    for (uint32_t j = i; j-- > 0 && r_mz[ord[j]] == m;) if (aux[j] != kNoSurvivor) { found = aux[j]; break; }
                    if (found == kNoSurvivor)
                        for (uint32_t j = i + 1; j < P && r_mz[ord[j]] == m; ++j) if (aux[j] != kNoSurvivor) { found = aux[j]; break; }
                    aux[i] = found;
                }
                cc[3] = cfg.mvh_bins - (int32_t)N;
                for (int c = 0; c < 4; ++c) s_cls[c] = cc[c];
                s_nsurv = N;
                s_flags = N >= 7 ? 1u : 0u;
            }
            __syncthreads();
            nsurv = s_nsurv;
        }

        // ---- blob: eligible peaks (intensity >= limit) in m/z order ----
        uint8_t* blob = blobs + blob_off[item];
        // compaction index by a serial prefix over 128-wide chunks (P is small)
        if (threadIdx.x == 0) {
            uint32_t e = 0;
            for (uint32_t i = 0; i < P; ++i) { bool el = r_in[ord[i]] >= limit; ord[i] |= el ? 0x80000000u : 0u; ordI[i] = e; e += el ? 1u : 0u; }
            s_E = e;
        }
        __syncthreads();
        const uint32_t E = s_E;
        const uint32_t Epad = (E + 3u) & ~3u;
        double* pmz = reinterpret_cast<double*>(blob + sizeof(BlobHeader));
        uint32_t* pinfo = reinterpret_cast<uint32_t*>(blob + sizeof(BlobHeader) + (size_t)Epad * 8);
        double* bval = reinterpret_cast<double*>(blob + sizeof(BlobHeader) + (size_t)Epad * 12);
        uint16_t* bcells = reinterpret_cast<uint16_t*>(blob + sizeof(BlobHeader) + (size_t)Epad * 12 + kValSlots * 8);
        for (uint32_t c = threadIdx.x; c < (uint32_t)kCells; c += blockDim.x) cells[c] = (uint16_t)E;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
            uint32_t o = ord[i];
            if (!(o & 0x80000000u)) continue;
            uint32_t f = o & 0x7fffffffu;
            uint32_t e = ordI[i];
            double m = r_mz[f];
            // dense rank of the intensity among all raw peaks (+1): equal intensities share a rank
            double v = r_in[f];
            uint32_t lo = 0, hi = P;
            while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (w_d[mid] < v) lo = mid + 1; else hi = mid; }
            pmz[e] = m;
            pinfo[e] = ((lo + 1u) << 16) | (aux[i] & 0xFFFFu);
            // first-index table over the monotone cell map
            int ce = cell_of(m);
            int cp = -1;
            // previous eligible peak
            for (uint32_t j = i; j-- > 0;) if (ord[j] & 0x80000000u) { cp = cell_of(r_mz[ord[j] & 0x7fffffffu]); break; }
            for (int c = cp + 1; c <= ce; ++c) cells[c] = (uint16_t)e;
        }
        for (uint32_t e = E + threadIdx.x; e < Epad; e += blockDim.x) { pmz[e] = inf; pinfo[e] = kNoSurvivor; }
        if (cfg.method == 1) for (uint32_t k = threadIdx.x; k < (uint32_t)kValSlots; k += blockDim.x) bval[k] = k < nsurv ? s_val[k] : 0.0;
        else for (uint32_t k = threadIdx.x; k < (uint32_t)kValSlots; k += blockDim.x) bval[k] = 0.0;
        __syncthreads();
        for (uint32_t c = threadIdx.x; c < (uint32_t)kCells / 2; c += blockDim.x)
            reinterpret_cast<uint32_t*>(bcells)[c] = reinterpret_cast<const uint32_t*>(cells)[c];
        if (threadIdx.x == 0) {
            BlobHeader h;
            h.n_peaks = E; h.n_pad = Epad; h.n_survivors = nsurv; h.flags = s_flags;
            for (int c = 0; c < 4; ++c) h.cls_count[c] = cfg.method == 2 ? s_cls[c] : 0;
            h.total_bytes = blob_bytes(Epad); h.raw_peaks = P; h.pad[0] = h.pad[1] = 0;
            *reinterpret_cast<BlobHeader*>(blob) = h;
        }
        __syncthreads();
    }
}

}  // namespace

int launch_prepare(const SpectraView& sp, const int32_t* list, int32_t n_list, const PrepConfig& cfg, uint8_t* blobs,
                   const uint64_t* blob_off, uint32_t max_peaks, cudaStream_t st) {
    if (n_list <= 0) return 0;
    uint32_t npad = 32; while (npad < max_peaks) npad <<= 1;
    size_t smem = prep_smem_bytes(npad);
    cudaFuncSetAttribute(k0_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int per_sm = (int)(200 * 1024 / (smem + 1024)); if (per_sm < 1) per_sm = 1; if (per_sm > 12) per_sm = 12;
    int grid = sms * per_sm; if (grid > n_list) grid = n_list;
    k0_prepare_kernel<<<grid, kPrepBlock, smem, st>>>(sp, list, n_list, cfg, blobs, blob_off, npad);
    return 1;
}

}  // namespace pq

// k3_window.cu — spectrum packing helpers and Step-2 precursor-window enumeration on device.
//
// Packing replaces the per-spectrum task submission of SpectraInput.readMSMSfast (pg/SpectraInput.java:91-279):
// spectra are re-ordered by precursor neutral mass into CSR arrays (north_star item 1).
// Enumeration replaces the 0.1-Da HashMap probe of FindCandidateSpectraAndScoreWorker.run
// (pg/FindCandidateSpectraAndScoreWorker.java:33-58): a conservative run of mass-sorted target rows per
// (spectrum, isotope error); K1 applies the exact bucket + tolerance predicate to every row of the run.
#include <cstdlib>

#include <algorithm>

#include "pq_internal.h"

namespace pq {

namespace {

// Precursor.getMass(charge) = mz*z - z*proton (e.g. DBSearchWorker.java:28).
// "Virtual" spectra: caller spectrum i is virtual spectrum i; a spectrum loaded without a charge (0) is searched as z = 2 and,
// only if that saved no PSM, as z = 3 (FindCandidateSpectraAndScoreWorker.java:96-228): it appears twice, as virtual i (z = 2)
// and as virtual n + j (z = 3, j = its rank among the charge-less spectra, zero_idx[j] = i).  Both share the caller's peaks.
__global__ void k_virtual_meta(int64_t nv, int64_t n, const double* __restrict__ prec_mz, const int32_t* __restrict__ charge,
                               const int32_t* __restrict__ zero_idx, double* __restrict__ mass, int32_t* __restrict__ vz,
                               uint64_t* __restrict__ key, int32_t* __restrict__ idx) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const int64_t src = v < n ? v : (int64_t)zero_idx[v - n];
    int32_t zi = charge[src];
    if (zi == 0) zi = v < n ? 2 : 3;
    double z = (double)zi;
    double m = __dsub_rn(__dmul_rn(prec_mz[src], z), __dmul_rn(z, kProton));
    mass[v] = m; vz[v] = zi;
    // order-preserving key for positive doubles; negative/NaN masses sort first
    uint64_t u = (uint64_t)__double_as_longlong(m);
    key[v] = (m > 0.0) ? u : 0ull;
    idx[v] = (int32_t)v;
}

// sorted position i <- virtual spectrum order[i]
__global__ void k_gather_meta(int64_t nv, int64_t n, const int32_t* __restrict__ order, const double* __restrict__ prec_mz_in,
                              const int32_t* __restrict__ charge_in, const int32_t* __restrict__ zero_idx, const int32_t* __restrict__ vz,
                              const double* __restrict__ mass_in, const uint64_t* __restrict__ off_in, double* __restrict__ prec_mz,
                              int32_t* __restrict__ charge, double* __restrict__ mass, int32_t* __restrict__ caller, uint32_t* __restrict__ npeaks,
                              uint8_t* __restrict__ alias) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nv) return;
    const int32_t v = order[i];
    const int64_t src = v < n ? (int64_t)v : (int64_t)zero_idx[v - n];
    prec_mz[i] = prec_mz_in[src]; charge[i] = vz[v]; mass[i] = mass_in[v]; caller[i] = (int32_t)src;
    npeaks[i] = (uint32_t)(off_in[src + 1] - off_in[src]);
    alias[i] = v >= n ? 2 : (charge_in[src] == 0 ? 1 : 0);      // 0: charge given; 1: charge-less searched as z = 2; 2: as z = 3
}

// flag of every virtual spectrum (caller order) from the flags in mass order
__global__ void k_flags_by_virtual(int64_t nv, const int32_t* __restrict__ inv, const int32_t* __restrict__ flag_sorted, uint8_t* __restrict__ flag_virtual) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nv) flag_virtual[v] = flag_sorted[inv[v]] ? 1 : 0;
}

// storage size of every sorted spectrum: its peak count when its peaks are (to be) resident, else 0
__global__ void k_resident_counts(int64_t nv, const uint32_t* __restrict__ npeaks, const int32_t* __restrict__ resident /* null = all */,
                                  uint64_t* __restrict__ count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nv) return;
    count[i] = (i < nv && (!resident || resident[i])) ? (uint64_t)npeaks[i] : 0ull;
}

// Candidate-first upload: one warp per needed spectrum reads its peaks straight from the caller's pinned (mapped) host arrays
// over PCIe and writes them to their place in the mass-sorted CSR.  list = sorted indices IN CALLER ORDER: neighbouring warps
// read neighbouring host addresses, so the 128-byte line two adjacent spectra share crosses once (peak lists start at
// arbitrary 8-byte offsets; read in mass order every spectrum paid for a partial line at both ends: 41 instead of ~50 GB/s).
// Every lane issues all its loads before the first store (up to 8 x 8 bytes in flight per lane).
__global__ void __launch_bounds__(128)
k_gather_peaks_from_host(const int32_t* __restrict__ list, int32_t n_list, const int32_t* __restrict__ caller, const uint64_t* __restrict__ off_in,
                         const uint64_t* __restrict__ off_out, const double* __restrict__ h_mz, const double* __restrict__ h_in,
                         double* __restrict__ mz_out, double* __restrict__ in_out) {
    const int lane = threadIdx.x & 31;
    const int32_t warps = (int32_t)((gridDim.x * blockDim.x) >> 5);
    for (int32_t w = (int32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < n_list; w += warps) {
        const int32_t s = list[w];
        const int64_t src = caller[s];
        const uint64_t a = off_in[src], cnt = off_in[src + 1] - a, b = off_out[s];
        for (uint64_t k0 = 0; k0 < cnt; k0 += 128) {
            double m[4], v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t k = k0 + (uint64_t)(32 * j + lane);
                if (k < cnt) { m[j] = __ldcs(h_mz + a + k); v[j] = __ldcs(h_in + a + k); }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint64_t k = k0 + (uint64_t)(32 * j + lane);
                if (k < cnt) { mz_out[b + k] = m[j]; in_out[b + k] = v[j]; }
            }
        }
    }
}

// The same gather as ONE launch over all chunks: its few CTAs stay resident for the whole upload and count every finished chunk
// in chunk_done[c] (release), so the upload never queues behind the persistent scoring kernels of the parts that are searched
// meanwhile (seen in the timeline: every per-chunk gather launch waited up to 2 ms for a CTA slot; upload 35 ms instead of 22).
// Consumers order themselves behind chunk c with k_wait_chunk on their own stream.
__global__ void __launch_bounds__(128)
k_gather_peaks_chunked(const int32_t* __restrict__ list, int32_t n_list, int32_t chunk, const int32_t* __restrict__ caller, const uint64_t* __restrict__ off_in,
                       const uint64_t* __restrict__ off_out, const double* __restrict__ h_mz, const double* __restrict__ h_in,
                       double* __restrict__ mz_out, double* __restrict__ in_out, uint32_t* __restrict__ chunk_done) {
    const int lane = threadIdx.x & 31;
    const int32_t warps = (int32_t)((gridDim.x * blockDim.x) >> 5);
    const int32_t w0 = (int32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    int32_t c = 0;
    for (int32_t first = 0; first < n_list; first += chunk, ++c) {
        const int32_t last = min(first + chunk, n_list);
        for (int32_t w = first + w0; w < last; w += warps) {
            const int32_t s = list[w];
            const int64_t src = caller[s];
            const uint64_t a = off_in[src], cnt = off_in[src + 1] - a, b = off_out[s];
            for (uint64_t k0 = 0; k0 < cnt; k0 += 128) {
                double m[4], v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t k = k0 + (uint64_t)(32 * j + lane);
                    if (k < cnt) { m[j] = __ldcs(h_mz + a + k); v[j] = __ldcs(h_in + a + k); }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint64_t k = k0 + (uint64_t)(32 * j + lane);
                    if (k < cnt) { mz_out[b + k] = m[j]; in_out[b + k] = v[j]; }
                }
            }
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(chunk_done + c, 1u);
    }
}
__global__ void k_wait_chunk(const uint32_t* flag, uint32_t expected) {
    uint32_t v;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= expected) break;
        __nanosleep(500);
    }
}

// Candidate-first from PAGEABLE caller arrays: host threads pack the needed spectra (gather order) into a page-locked ring, the chunk
// goes up with one copy, k_scatter_staged puts every spectrum in its place of the mass-sorted CSR and k_set_flag releases the chunk.
__global__ void k_gather_sources(const int32_t* __restrict__ list, int32_t n, const int32_t* __restrict__ caller, const uint32_t* __restrict__ npeaks,
                                 int32_t* __restrict__ src, uint64_t* __restrict__ cnt) {
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { cnt[i] = 0; return; }
    const int32_t s = list[i];
    src[i] = caller[s]; cnt[i] = npeaks[s];
}
__global__ void __launch_bounds__(256)
k_scatter_staged(const int32_t* __restrict__ list, int32_t n, const uint64_t* __restrict__ g_off /* stream offsets of these entries (+ one) */, uint64_t base,
                 const uint64_t* __restrict__ off_out, const double* __restrict__ st_mz, const double* __restrict__ st_in, double* __restrict__ mz_out, double* __restrict__ in_out) {
    const int lane = threadIdx.x & 31;
    const int32_t warps = (int32_t)((gridDim.x * blockDim.x) >> 5);
    for (int32_t w = (int32_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); w < n; w += warps) {
        const uint64_t a = g_off[w] - base, cnt = g_off[w + 1] - g_off[w], b = off_out[list[w]];
        for (uint64_t k = lane; k < cnt; k += 32) { mz_out[b + k] = st_mz[a + k]; in_out[b + k] = st_in[a + k]; }
    }
}
__global__ void k_set_flag(uint32_t* flag, uint32_t v) { __threadfence(); atomicExch(flag, v); }

// full-upload path: peaks arrive in caller order chunk by chunk, so the copy is driven by the virtual id
// (inv[v] = mass-sorted position of virtual spectrum v; src = the caller spectrum whose peaks it has)
__global__ void k_invert_order(int64_t n, const int32_t* __restrict__ order, int32_t* __restrict__ inv) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[order[i]] = (int32_t)i;
}
__global__ void k_gather_peaks_by_virtual(int64_t first, int64_t count, int64_t n, const int32_t* __restrict__ zero_idx, const int32_t* __restrict__ inv,
                                          const uint64_t* __restrict__ off_in, const uint64_t* __restrict__ off_out, const double* __restrict__ mz_in,
                                          const double* __restrict__ in_in, double* __restrict__ mz_out, double* __restrict__ in_out) {
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= count) return;
    int64_t v = first + w;
    int64_t src = v < n ? v : (int64_t)zero_idx[v - n];
    uint64_t a = off_in[src], cnt = off_in[src + 1] - a, b = off_out[inv[v]];
    for (uint64_t k = lane; k < cnt; k += 32) { mz_out[b + k] = mz_in[a + k]; in_out[b + k] = in_in[a + k]; }
}

__device__ __forceinline__ uint32_t lower_bound_d(const double* a, uint32_t n, double v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void k_enumerate_targets(SpectraView sp, const double* __restrict__ table_mass, uint32_t n_rows, WindowConfig cfg,
                                    Job* __restrict__ jobs, uint32_t* __restrict__ n_jobs, int32_t* __restrict__ need) {
    // n_jobs[0] = job count, n_jobs[2..3] = 64-bit total of candidate rows (sizes the Step-2 hit buffer)
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t s = t / cfg.n_isotope; int ii = (int)(t % cfg.n_isotope);
    if (s >= sp.n) return;
    uint64_t P = sp.npeaks[s];
    if ((int64_t)P <= (int64_t)cfg.min_peaks) return;                  // SpectraInput.java:183
    double m = sp.mass[s];
    bool plain = cfg.n_isotope == 1 && cfg.isotope[0] == 0;           // CParameter.java:423-448
    if (!plain) m = __dsub_rn(m, __dmul_rn((double)cfg.isotope[ii], kIsotope));
    double w = cfg.tol_ppm ? cfg.tol * 1e-6 * (fabs(m) + 1.0) * 1.01 + 1e-6 : cfg.tol * 1.01 + 1e-6;
    if (w > 0.2 && !cfg.no_bucket) w = 0.2;                            // the +-1 bucket rule never reaches further
    uint32_t lo = lower_bound_d(table_mass, n_rows, m - w);
    uint32_t hi = lower_bound_d(table_mass, n_rows, m + w);
    if (hi <= lo) return;
    uint32_t slot = atomicAdd(n_jobs, 1u);
    Job J;
    J.blob = 0; J.cand_begin = lo; J.cand_count = hi - lo; J.out = slot; J.mass = m; J.ref_score = 0.0;
    J.charge = sp.charge[s]; J.iso = plain ? 0 : cfg.isotope[ii]; J.spectrum = (int32_t)s; J.aux = 0;
    jobs[slot] = J;
    need[s] = 1;
    atomicAdd(reinterpret_cast<unsigned long long*>(n_jobs + 2), (unsigned long long)(hi - lo));
}

// candidate-first residency check: spectra the present targets need (need[s]) whose peaks were not brought over (resident[s] == 0)
__global__ void k_count_missing(int64_t n, const int32_t* __restrict__ need, const int32_t* __restrict__ resident, uint32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool miss = i < n && need[i] && !resident[i];
    const uint32_t b = __ballot_sync(0xffffffffu, miss);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(out, (uint32_t)__popc(b));
}

// Spectra loaded without a charge are two virtual spectra (z = 2, z = 3).  FindCandidateSpectraAndScoreWorker.java:96-228 tries
// z = 3 only when z = 2 saved no PSM: mark the virtual spectra with a Step-2 hit, then drop the hits and the pair counts of
// every z = 3 reading whose z = 2 sibling has one.
__global__ void k_mark_hits(const Hit* __restrict__ hits, uint32_t nh, uint8_t* __restrict__ has_hit) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nh) has_hit[hits[i].spectrum] = 1;
}
__device__ __forceinline__ bool z3_superseded(int32_t s, const uint8_t* alias, int64_t n_caller, const int32_t* zero_idx, const int32_t* vid, const int32_t* inv,
                                              const uint8_t* has_hit) {
    if (alias[s] != 2) return false;
    const int32_t caller = zero_idx[vid[s] - n_caller];               // its z = 2 reading is virtual spectrum `caller`
    return has_hit[inv[caller]] != 0;
}
__global__ void k_chargeless_flags(const Hit* __restrict__ hits, uint32_t nh, const Job* __restrict__ jobs, uint32_t n_jobs, const uint8_t* __restrict__ alias,
                                   int64_t n_caller, const int32_t* __restrict__ zero_idx, const int32_t* __restrict__ vid, const int32_t* __restrict__ inv,
                                   const uint8_t* __restrict__ has_hit, uint8_t* __restrict__ skip_job, uint8_t* __restrict__ keep_hit) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nh) keep_hit[i] = z3_superseded(hits[i].spectrum, alias, n_caller, zero_idx, vid, inv, has_hit) ? 0 : 1;
    if (i < n_jobs) skip_job[i] = z3_superseded(jobs[i].spectrum, alias, n_caller, zero_idx, vid, inv, has_hit) ? 1 : 0;
}

__global__ void k_blob_sizes(SpectraView sp, const int32_t* __restrict__ list, int32_t n_list, uint64_t* __restrict__ sizes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_list) return;
    if (i == n_list) { sizes[i] = 0; return; }
    int32_t s = list[i];
    uint32_t P = sp.npeaks[s];
    sizes[i] = blob_bytes(P);
}

__global__ void k_fix_job_blobs(Job* __restrict__ jobs, uint32_t n_jobs, const int32_t* __restrict__ prep_id) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_jobs) return;
    jobs[i].blob = (uint32_t)prep_id[jobs[i].spectrum];
}

// Step-2 jobs leave k_enumerate_targets in atomic-slot order.  The scatter scorer gives every lane its own job, so a warp
// runs as long as its slowest lane and serialises lanes that differ in fragment-charge count: order the jobs by
// (precursor charge, first target row, row count) -> neighbouring lanes score peptides of the same length and charge.
__global__ void k_job_keys(const Job* __restrict__ jobs, uint32_t n, uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Job& J = jobs[i];
    uint64_t z = (uint64_t)min(max(J.charge, 0), 15);
    key[i] = (z << 40) | ((uint64_t)J.cand_begin << 8) | (uint64_t)min(J.cand_count, 255u);
    idx[i] = i;
}
__global__ void k_gather_jobs(const Job* __restrict__ in, const uint32_t* __restrict__ idx, uint32_t n, Job* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[idx[i]];
}

}  // namespace

int launch_job_keys(const Job* jobs, uint32_t n, uint64_t* key, uint32_t* idx, cudaStream_t st) {
    if (!n) return 0;
    k_job_keys<<<(n + 255) / 256, 256, 0, st>>>(jobs, n, key, idx);
    return 1;
}
int launch_gather_jobs(const Job* in, const uint32_t* idx, uint32_t n, Job* out, cudaStream_t st) {
    if (!n) return 0;
    k_gather_jobs<<<(n + 255) / 256, 256, 0, st>>>(in, idx, n, out);
    return 1;
}

int launch_virtual_meta(int64_t nv, int64_t n, const double* prec_mz, const int32_t* charge, const int32_t* zero_idx, double* mass, int32_t* vz,
                        uint64_t* key, int32_t* idx, cudaStream_t st) {
    if (nv <= 0) return 0;
    k_virtual_meta<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(nv, n, prec_mz, charge, zero_idx, mass, vz, key, idx);
    return 1;
}
int launch_gather_meta(int64_t nv, int64_t n, const int32_t* order, const double* prec_mz_in, const int32_t* charge_in, const int32_t* zero_idx,
                       const int32_t* vz, const double* mass_in, const uint64_t* off_in, double* prec_mz, int32_t* charge, double* mass,
                       int32_t* caller, uint32_t* npeaks, uint8_t* alias, cudaStream_t st) {
    if (nv <= 0) return 0;
    k_gather_meta<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(nv, n, order, prec_mz_in, charge_in, zero_idx, vz, mass_in, off_in, prec_mz, charge, mass,
                                                               caller, npeaks, alias);
    return 1;
}
int launch_flags_by_virtual(int64_t nv, const int32_t* inv, const int32_t* flag_sorted, uint8_t* flag_virtual, cudaStream_t st) {
    if (nv <= 0) return 0;
    k_flags_by_virtual<<<(unsigned)((nv + 255) / 256), 256, 0, st>>>(nv, inv, flag_sorted, flag_virtual);
    return 1;
}
int launch_resident_counts(int64_t nv, const uint32_t* npeaks, const int32_t* resident, uint64_t* count, cudaStream_t st) {
    k_resident_counts<<<(unsigned)((nv + 1 + 255) / 256), 256, 0, st>>>(nv, npeaks, resident, count);
    return 1;
}
int launch_gather_peaks_from_host(const int32_t* list, int32_t n_list, const int32_t* caller, const uint64_t* off_in, const uint64_t* off_out,
                                  const double* h_mz, const double* h_in, double* mz_out, double* in_out, int sms, cudaStream_t st) {
    if (n_list <= 0) return 0;
    // A FEW CTAs only: reads of host memory stay ~2 us in an SM's miss queues, and K0 running on the same SM was seen to slow down
    // 3-4x; PCIe needs only some hundred KB in flight, so the gather occupies a fraction of the SMs and leaves the rest to K0.
    int grid = 32; (void)sms;
    if (const char* e = getenv("PQ_GATHER_CTAS")) { int v = atoi(e); if (v > 0) grid = v; }
    int need = (n_list + 3) / 4; if (grid > need) grid = need;
    k_gather_peaks_from_host<<<grid, 128, 0, st>>>(list, n_list, caller, off_in, off_out, h_mz, h_in, mz_out, in_out);
    return 1;
}
int gather_grid(int32_t n_list) {
    int grid = 32;
    if (const char* e = getenv("PQ_GATHER_CTAS")) { int v = atoi(e); if (v > 0) grid = v; }
    int need = (n_list + 3) / 4; if (grid > need) grid = need;
    return grid < 1 ? 1 : grid;
}
int launch_gather_peaks_chunked(const int32_t* list, int32_t n_list, int32_t chunk, const int32_t* caller, const uint64_t* off_in, const uint64_t* off_out,
                                const double* h_mz, const double* h_in, double* mz_out, double* in_out, uint32_t* chunk_done /* zeroed */, uint32_t* n_ctas, cudaStream_t st) {
    if (n_list <= 0) { *n_ctas = 0; return 0; }
    const int grid = gather_grid(n_list);
    *n_ctas = (uint32_t)grid;
    k_gather_peaks_chunked<<<grid, 128, 0, st>>>(list, n_list, chunk, caller, off_in, off_out, h_mz, h_in, mz_out, in_out, chunk_done);
    return 1;
}
int launch_gather_sources(const int32_t* list, int32_t n, const int32_t* caller, const uint32_t* npeaks, int32_t* src, uint64_t* cnt, cudaStream_t st) {
    k_gather_sources<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(list, n, caller, npeaks, src, cnt);
    return 1;
}
int launch_scatter_staged(const int32_t* list, int32_t n, const uint64_t* g_off, uint64_t base, const uint64_t* off_out, const double* st_mz, const double* st_in,
                          double* mz_out, double* in_out, uint32_t* flag, uint32_t flag_value, cudaStream_t st) {
    if (n > 0) k_scatter_staged<<<std::min((n + 7) / 8, 148 * 8), 256, 0, st>>>(list, n, g_off, base, off_out, st_mz, st_in, mz_out, in_out);
    k_set_flag<<<1, 1, 0, st>>>(flag, flag_value);
    return 2;
}
int launch_wait_chunk(const uint32_t* flag, uint32_t expected, cudaStream_t st) {
    k_wait_chunk<<<1, 1, 0, st>>>(flag, expected);
    return 1;
}
int launch_invert_order(int64_t n, const int32_t* order, int32_t* inv, cudaStream_t st) {
    if (n <= 0) return 0;
    k_invert_order<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, order, inv);
    return 1;
}
int launch_gather_peaks_by_virtual(int64_t first, int64_t count, int64_t n, const int32_t* zero_idx, const int32_t* inv, const uint64_t* off_in,
                                   const uint64_t* off_out, const double* mz_in, const double* in_in, double* mz_out, double* in_out, cudaStream_t st) {
    if (count <= 0) return 0;
    int64_t threads = count * 32;
    k_gather_peaks_by_virtual<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(first, count, n, zero_idx, inv, off_in, off_out, mz_in, in_in, mz_out, in_out);
    return 1;
}
int launch_enumerate_targets(const SpectraView& sp, const double* table_mass, uint32_t n_rows, const WindowConfig& cfg,
                             Job* jobs, uint32_t* d_n_jobs, int32_t* need_prepare, cudaStream_t st) {
    int64_t t = sp.n * cfg.n_isotope;
    if (t <= 0 || n_rows == 0) return 0;
    k_enumerate_targets<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(sp, table_mass, n_rows, cfg, jobs, d_n_jobs, need_prepare);
    return 1;
}
int launch_count_missing(int64_t n, const int32_t* need, const int32_t* resident, uint32_t* out, cudaStream_t st) {
    if (n <= 0) return 0;
    k_count_missing<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, need, resident, out);
    return 1;
}
int launch_chargeless_filter(const Hit* hits, uint32_t nh, const Job* jobs, uint32_t n_jobs, const SpectraView& sp, int64_t n_caller, const int32_t* zero_idx,
                             const int32_t* vid, const int32_t* inv, uint8_t* has_hit /* zeroed */, uint8_t* skip_job, uint8_t* keep_hit, cudaStream_t st) {
    int k = 0;
    if (nh) { k_mark_hits<<<(nh + 255) / 256, 256, 0, st>>>(hits, nh, has_hit); ++k; }
    const uint32_t m = nh > n_jobs ? nh : n_jobs;
    if (m) { k_chargeless_flags<<<(m + 255) / 256, 256, 0, st>>>(hits, nh, jobs, n_jobs, sp.alias, n_caller, zero_idx, vid, inv, has_hit, skip_job, keep_hit); ++k; }
    return k;
}
int launch_blob_sizes(const SpectraView& sp, const int32_t* list, int32_t n_list, uint64_t* sizes, cudaStream_t st) {
    k_blob_sizes<<<(unsigned)((n_list + 1 + 255) / 256), 256, 0, st>>>(sp, list, n_list, sizes);
    return 1;
}
int launch_fix_job_blobs(Job* jobs, uint32_t n_jobs, const int32_t* prep_id, cudaStream_t st) {
    if (n_jobs == 0) return 0;
    k_fix_job_blobs<<<(n_jobs + 255) / 256, 256, 0, st>>>(jobs, n_jobs, prep_id);
    return 1;
}

}  // namespace pq

// k3_window.cu — spectrum packing helpers and Step-2 precursor-window enumeration on device.
//
// Packing replaces the per-spectrum task submission of SpectraInput.readMSMSfast (pg/SpectraInput.java:91-279):
// spectra are re-ordered by precursor neutral mass into CSR arrays (north_star item 1).
// Enumeration replaces the 0.1-Da HashMap probe of FindCandidateSpectraAndScoreWorker.run
// (pg/FindCandidateSpectraAndScoreWorker.java:33-58): a conservative run of mass-sorted target rows per
// (spectrum, isotope error); K1 applies the exact bucket + tolerance predicate to every row of the run.
#include "pq_internal.h"

namespace pq {

namespace {

// Precursor.getMass(charge) = mz*z - z*proton (e.g. DBSearchWorker.java:28)
__global__ void k_precursor_mass(int64_t n, const double* __restrict__ prec_mz, const int32_t* __restrict__ charge,
                                 double* __restrict__ mass, uint64_t* __restrict__ key, int32_t* __restrict__ idx) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double z = (double)charge[i];
    double m = __dsub_rn(__dmul_rn(prec_mz[i], z), __dmul_rn(z, kProton));
    mass[i] = m;
    // order-preserving key for positive doubles; negative/NaN masses sort first
    uint64_t u = (uint64_t)__double_as_longlong(m);
    key[i] = (m > 0.0) ? u : 0ull;
    idx[i] = (int32_t)i;
}

__global__ void k_gather_meta(int64_t n, const int32_t* __restrict__ order, const double* __restrict__ prec_mz_in,
                              const int32_t* __restrict__ charge_in, const double* __restrict__ mass_in,
                              const uint64_t* __restrict__ off_in, double* __restrict__ prec_mz, int32_t* __restrict__ charge,
                              double* __restrict__ mass, uint64_t* __restrict__ count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t s = order[i];
    prec_mz[i] = prec_mz_in[s]; charge[i] = charge_in[s]; mass[i] = mass_in[s];
    count[i] = off_in[s + 1] - off_in[s];
}

// one warp per spectrum copies its peaks to the sorted position (coalesced 8-byte lanes)
__global__ void k_gather_peaks(int64_t n, const int32_t* __restrict__ order, const uint64_t* __restrict__ off_in,
                               const uint64_t* __restrict__ off_out, const double* __restrict__ mz_in,
                               const double* __restrict__ in_in, double* __restrict__ mz_out, double* __restrict__ in_out) {
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= n) return;
    int32_t s = order[w];
    uint64_t a = off_in[s], cnt = off_in[s + 1] - a, b = off_out[w];
    for (uint64_t k = lane; k < cnt; k += 32) { mz_out[b + k] = mz_in[a + k]; in_out[b + k] = in_in[a + k]; }
}

// prepare-at-load path: peaks arrive in caller order chunk by chunk, so the copy is driven by the caller index
// (inv[i] = mass-sorted position of caller spectrum i)
__global__ void k_invert_order(int64_t n, const int32_t* __restrict__ order, int32_t* __restrict__ inv) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[order[i]] = (int32_t)i;
}
__global__ void k_gather_peaks_by_caller(int64_t first, int64_t count, const int32_t* __restrict__ inv, const uint64_t* __restrict__ off_in,
                                         const uint64_t* __restrict__ off_out, const double* __restrict__ mz_in,
                                         const double* __restrict__ in_in, double* __restrict__ mz_out, double* __restrict__ in_out) {
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= count) return;
    int64_t s = first + w;
    uint64_t a = off_in[s], cnt = off_in[s + 1] - a, b = off_out[inv[s]];
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
    uint64_t P = sp.off[s + 1] - sp.off[s];
    if ((int64_t)P <= (int64_t)cfg.min_peaks) return;                  // SpectraInput.java:183
    double m = sp.mass[s];
    bool plain = cfg.n_isotope == 1 && cfg.isotope[0] == 0;           // CParameter.java:423-448
    if (!plain) m = __dsub_rn(m, __dmul_rn((double)cfg.isotope[ii], kIsotope));
    double w = cfg.tol_ppm ? cfg.tol * 1e-6 * (fabs(m) + 1.0) * 1.01 + 1e-6 : cfg.tol * 1.01 + 1e-6;
    if (w > 0.2) w = 0.2;                                              // the +-1 bucket rule never reaches further
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

__global__ void k_blob_sizes(SpectraView sp, const int32_t* __restrict__ list, int32_t n_list, uint64_t* __restrict__ sizes) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_list) return;
    if (i == n_list) { sizes[i] = 0; return; }
    int32_t s = list[i];
    uint32_t P = (uint32_t)(sp.off[s + 1] - sp.off[s]);
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

int launch_precursor_mass(int64_t n, const double* prec_mz, const int32_t* charge, double* mass, uint64_t* key, int32_t* idx, cudaStream_t st) {
    if (n <= 0) return 0;
    k_precursor_mass<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, prec_mz, charge, mass, key, idx);
    return 1;
}
int launch_gather_meta(int64_t n, const int32_t* order, const double* prec_mz_in, const int32_t* charge_in, const double* mass_in,
                       const uint64_t* off_in, double* prec_mz, int32_t* charge, double* mass, uint64_t* count, cudaStream_t st) {
    if (n <= 0) return 0;
    k_gather_meta<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, order, prec_mz_in, charge_in, mass_in, off_in, prec_mz, charge, mass, count);
    return 1;
}
int launch_gather_peaks(int64_t n, const int32_t* order, const uint64_t* off_in, const uint64_t* off_out, const double* mz_in,
                        const double* in_in, double* mz_out, double* in_out, cudaStream_t st) {
    if (n <= 0) return 0;
    int64_t threads = n * 32;
    k_gather_peaks<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(n, order, off_in, off_out, mz_in, in_in, mz_out, in_out);
    return 1;
}
int launch_invert_order(int64_t n, const int32_t* order, int32_t* inv, cudaStream_t st) {
    if (n <= 0) return 0;
    k_invert_order<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, order, inv);
    return 1;
}
int launch_gather_peaks_by_caller(int64_t first, int64_t count, const int32_t* inv, const uint64_t* off_in, const uint64_t* off_out,
                                  const double* mz_in, const double* in_in, double* mz_out, double* in_out, cudaStream_t st) {
    if (count <= 0) return 0;
    int64_t threads = count * 32;
    k_gather_peaks_by_caller<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(first, count, inv, off_in, off_out, mz_in, in_in, mz_out, in_out);
    return 1;
}
int launch_enumerate_targets(const SpectraView& sp, const double* table_mass, uint32_t n_rows, const WindowConfig& cfg,
                             Job* jobs, uint32_t* d_n_jobs, int32_t* need_prepare, cudaStream_t st) {
    int64_t t = sp.n * cfg.n_isotope;
    if (t <= 0 || n_rows == 0) return 0;
    k_enumerate_targets<<<(unsigned)((t + 255) / 256), 256, 0, st>>>(sp, table_mass, n_rows, cfg, jobs, d_n_jobs, need_prepare);
    return 1;
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

// k7_steps.cu — K7: the step drivers on the device.  The PSM table is built, joined with the Step-3/4 results, ranked and
// validated without leaving HBM; the host sees only counts until pq_get_psms exports the records.
//
// Replaces the bookkeeping the reference does on the Java heap around its worker pools:
//   Step 2  JPeptide.saveMatchedSpectrum lists (PSMMatch/JPeptide.java:39-50)          -> psm_table_from_hits
//   Step 3  DatabaseInput.get_spectra2score "max" (pg/DatabaseInput.java:637-669), the per-spectrum DBSearchWorker
//           submission (:695-700) and the n_db/total_db/valid write-back (:740-753)      -> k7_step3_jobs / k7_step3_aggregate
//           detail.txt rows (pg/DBSearchWorker.java:71-76, 93-98)                       -> k7_comp3_*
//   Step 4  StatisticScoreCalc.run (pg/StatisticScoreCalc.java:19-63): which sequences get shuffles, one
//           StatisticScoreWorker per still-valid PSM, p = (nRand+1)/(N+1) (pg/StatisticScoreWorker.java:56-58);
//           group order and round count of generateRandomPeptidesFast (pg/PeptideInput.java:387-424)
//                                                                                        -> k7_flag_valid / k7_shuffle_plans /
//                                                                                           k7_layout_seqs / k7_layout_forms / k7_step4_*
//   rank    RankPSM.RankPSMCompare (pg/RankPSM.java:107-149), CParameter.passed_psm_validation (pg/CParameter.java:913-940)
//                                                                                        -> k7_rank
#include <cub/cub.cuh>

#include "pq_internal.h"
#include "k_cells.cuh"

namespace pq {

namespace {

#define RB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return false; } } while (0)

__device__ __forceinline__ uint32_t psm_subkey(int32_t form, int32_t iso) { return ((uint32_t)form << 3) | ((uint32_t)(iso + 4) & 7u); }

__global__ void k7_hit_keys(const Hit* __restrict__ hits, uint32_t n, const uint32_t* __restrict__ form_of_row, uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Hit& h = hits[i];
    key[i] = ((uint64_t)(uint32_t)h.spectrum << 32) | psm_subkey((int32_t)form_of_row[h.cand], h.iso);
    idx[i] = i;
}

__global__ void k7_build_psms(const Hit* __restrict__ hits, const uint32_t* __restrict__ idx, uint32_t n, SpectraView sp, const double* __restrict__ t_mass,
                              const uint8_t* __restrict__ t_rows, TargetAux A, DevPsm* __restrict__ psms, uint32_t* __restrict__ head) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Hit h = hits[idx[i]];
    DevPsm p;
    p.spectrum_sorted = h.spectrum; p.spectrum = sp.caller_index[h.spectrum]; p.form = (int32_t)A.form_of_row[h.cand]; p.row = h.cand; p.blob = h.blob;
    p.charge = sp.charge[h.spectrum]; p.iso = h.iso; p.len = (int32_t)t_rows[(size_t)h.cand * kRowBytes];
    p.score = h.score; p.exp_mass = h.mass; p.pep_mass = t_mass[h.cand]; p.p_value = 100.0;
    p.ca = h.count_a; p.cb = h.count_b; p.n_db = 0; p.total_db = 0; p.n_random = -1; p.total_random = -1; p.valid = 1; p.rank = 0; p.n_ptm = -1; p.confident = 0;
    p.seg = 0; p.pad = 0;
    psms[i] = p;
    head[i] = (i == 0 || hits[idx[i - 1]].spectrum != h.spectrum) ? 1u : 0u;
}

__global__ void k7_set_segments(const uint32_t* __restrict__ head, const uint32_t* __restrict__ seg_inc, uint32_t n, DevPsm* __restrict__ psms,
                                uint32_t* __restrict__ seg_start, uint32_t* __restrict__ n_seg) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t s = seg_inc[i] - 1u;
    psms[i].seg = s;
    if (head[i]) seg_start[s] = i;
    if (i == n - 1) { seg_start[s + 1] = n; *n_seg = s + 1; }
}

__global__ void k7_out_keys(const DevPsm* __restrict__ psms, uint32_t n, uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = ((uint64_t)(uint32_t)psms[i].spectrum << 32) | psm_subkey(psms[i].form, psms[i].iso);
    idx[i] = i;
}

__global__ void k7_job_counters(const Job* __restrict__ jobs, const JobResult* __restrict__ res, uint32_t n, const uint32_t* __restrict__ sp_npeaks, int which,
                                DevCounters* __restrict__ c, const uint8_t* __restrict__ skip_job) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long pairs = 0, bytes = 0, bounded = 0;
    if (i < n && !(skip_job && skip_job[i])) {
        int32_t s = jobs[i].spectrum;
        unsigned long long P = sp_npeaks[s];
        pairs = res[i].n_in_window; bounded = res[i].pad;
        bytes = pairs * (16ull * P + 96ull);
    }
    for (int o = 16; o > 0; o >>= 1) { pairs += __shfl_xor_sync(0xffffffffu, pairs, o); bytes += __shfl_xor_sync(0xffffffffu, bytes, o); bounded += __shfl_xor_sync(0xffffffffu, bounded, o); }
    if ((threadIdx.x & 31) == 0 && pairs) { atomicAdd(&c->pairs[which], pairs); atomicAdd(&c->bytes[which], bytes); if (bounded) atomicAdd(&c->bounded, bounded); }
}

__device__ __forceinline__ uint32_t lower_bound_d(const double* a, uint32_t n, double v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

// One job per (matched spectrum, isotope error): the conservative run of reference rows around the precursor mass;
// K1 applies the exact bucket + tolerance predicate of DBSearchWorker.java:38-52 to every row of the run.
__global__ void k7_step3_jobs(const DevPsm* __restrict__ psms, const uint32_t* __restrict__ seg_start, uint32_t n_seg, const double* __restrict__ r_mass,
                              uint32_t n_rows, WindowConfig W, const double* __restrict__ sp_mass, Job* __restrict__ jobs, SegAgg* __restrict__ agg,
                              unsigned long long* __restrict__ total_cand) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = 0;
    if (t < n_seg * (uint32_t)W.n_isotope) {
        uint32_t seg = t / (uint32_t)W.n_isotope; int ii = (int)(t % (uint32_t)W.n_isotope);
        uint32_t a = seg_start[seg], b = seg_start[seg + 1];
        double best = psms[a].score;                                  // get_spectra2score(..., "max")
        for (uint32_t i = a + 1; i < b; ++i) if (best < psms[i].score) best = psms[i].score;
        const DevPsm& p0 = psms[a];
        double m = sp_mass[p0.spectrum_sorted];
        bool plain = W.n_isotope == 1 && W.isotope[0] == 0;
        if (!plain) m = __dsub_rn(m, __dmul_rn((double)W.isotope[ii], kIsotope));
        double w = W.tol_ppm ? W.tol * 1e-6 * (fabs(m) + 1.0) * 1.01 + 1e-6 : W.tol * 1.01 + 1e-6;
        if (w > 0.2) w = 0.2;
        uint32_t lo = lower_bound_d(r_mass, n_rows, m - w), hi = lower_bound_d(r_mass, n_rows, m + w);
        Job J;
        J.blob = p0.blob; J.cand_begin = lo; J.cand_count = hi > lo ? hi - lo : 0u; J.out = t; J.mass = m; J.ref_score = best;
        J.charge = p0.charge; J.iso = plain ? 0 : W.isotope[ii]; J.spectrum = p0.spectrum_sorted; J.aux = 0;
        jobs[t] = J;
        mine = J.cand_count;
        if (ii == 0) { SegAgg g; g.matched = 0; g.better = 0; g.best = -1.0; g.ref_best = best; g.best_row = 0xffffffffu; g.pad = 0; agg[seg] = g; }
    }
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total_cand, mine);
}

__global__ void k7_step3_aggregate(const JobResult* __restrict__ res, uint32_t n_seg, int n_iso, const uint32_t* __restrict__ seg_start,
                                   DevPsm* __restrict__ psms, SegAgg* __restrict__ agg) {
    uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= n_seg) return;
    SegAgg g = agg[seg];
    g.matched = 0; g.better = 0; g.best = -1.0; g.best_row = 0xffffffffu;
    for (int ii = 0; ii < n_iso; ++ii) {                                      // isotope order, strict '<' (DBSearchWorker.java:63-65)
        const JobResult r = res[(size_t)seg * n_iso + ii];
        g.matched += (int32_t)r.n_in_window; g.better += (int32_t)r.n_better;
        if (r.best_score > g.best) { g.best = r.best_score; g.best_row = r.best_cand; }
    }
    agg[seg] = g;
    for (uint32_t i = seg_start[seg]; i < seg_start[seg + 1]; ++i) {
        psms[i].total_db = g.matched; psms[i].n_db = g.better;
        if (g.better >= 1) psms[i].valid = 0;                                  // DatabaseInput.java:745-750
    }
}

__global__ void k7_comp3_hits(const Hit* __restrict__ hits, uint32_t nh, const int32_t* __restrict__ caller_index, const double* __restrict__ r_mass,
                              pq_competitor* __restrict__ rows, uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nh) return;
    const Hit& h = hits[i];
    pq_competitor c; c.spectrum = caller_index[h.spectrum]; c.ref_form = (int32_t)h.cand; c.ptm = -1; c.ptm_site = -1; c.score = h.score; c.pep_mass = r_mass[h.cand];
    rows[i] = c; keys[i] = ((uint64_t)(uint32_t)c.spectrum << 32) | h.cand; idx[i] = i;
}
// the best competitor of a spectrum is also a detail row when its score > 0 (DBSearchWorker.java:93-98), unless it is one already
__global__ void k7_comp3_best(const SegAgg* __restrict__ agg, uint32_t n_seg, const uint32_t* __restrict__ seg_start, const DevPsm* __restrict__ psms,
                              const double* __restrict__ r_mass, uint32_t nh, pq_competitor* __restrict__ rows, uint64_t* __restrict__ keys,
                              uint32_t* __restrict__ idx, uint32_t* __restrict__ count) {
    uint32_t seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg >= n_seg) return;
    const SegAgg g = agg[seg];
    if (!(g.best > 0) || g.best_row == 0xffffffffu) return;
    if (g.ref_best <= g.best && g.best >= 20.0) return;
    uint32_t slot = nh + atomicAdd(count, 1u);
    pq_competitor c; c.spectrum = psms[seg_start[seg]].spectrum; c.ref_form = (int32_t)g.best_row; c.ptm = -1; c.ptm_site = -1; c.score = g.best; c.pep_mass = r_mass[g.best_row];
    rows[slot] = c; keys[slot] = ((uint64_t)(uint32_t)c.spectrum << 32) | g.best_row; idx[slot] = slot;
}
__global__ void k7_gather_comp(const pq_competitor* __restrict__ in, const uint32_t* __restrict__ idx, uint32_t n, pq_competitor* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[idx[i]];
}

__global__ void k7_flag_valid(const DevPsm* __restrict__ psms, uint32_t n, const int32_t* __restrict__ seq_of_form, uint8_t* __restrict__ seq_flag,
                              uint32_t* __restrict__ form_flag, uint8_t* __restrict__ valid_flag, uint32_t* __restrict__ n_valid) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = i < n ? psms[i].valid : 0;
    if (i < n) {
        valid_flag[i] = v ? 1 : 0;
        if (v) {
            int32_t f = psms[i].form;
            const int z = psms[i].charge;                                       // fragment charges 1 .. z-1 (JMatch.java:398-407)
            atomicMax(&form_flag[f], (uint32_t)min(max(z - 1, 1), 255));
            seq_flag[seq_of_form[f]] = 1;
        }
    }
    const uint32_t b = __ballot_sync(0xffffffffu, v != 0);             // still-valid PSMs, one atomic per warp
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_valid, (uint32_t)__popc(b));
}

__global__ void k7_valid_flags(const DevPsm* __restrict__ psms, uint32_t n, uint8_t* __restrict__ valid_flag, uint32_t* __restrict__ n_valid) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = i < n ? psms[i].valid : 0;
    if (i < n) valid_flag[i] = v ? 1 : 0;
    const uint32_t b = __ballot_sync(0xffffffffu, v != 0);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_valid, (uint32_t)__popc(b));
}

__device__ unsigned long long choose_u64(int n, int k) {
    if (k < 0 || k > n) return 0ull;
    if (k > n - k) k = n - k;
    unsigned __int128 r = 1;
    for (int i = 1; i <= k; ++i) r = r * (unsigned)(n - k + i) / (unsigned)i;
    return (unsigned long long)r;
}

// Group order + round count of PeptideInput.generateRandomPeptidesFast (pg/PeptideInput.java:387-424), one thread per
// target sequence that needs shuffles.  java.util.HashMap<String,Integer> over single-character keys: hash = char code,
// table 16 slots doubling while size > 0.75 * capacity, iteration = slot order then insertion order; then a stable
// sort by count (pg/HashMapValueCompare.java:14-25).
__global__ void k7_shuffle_plans(const uint8_t* __restrict__ seq_flag, TargetAux A, int32_t n_random, ShuffleSeq* __restrict__ plans) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= A.n_seq || !seq_flag[s]) return;
    const uint8_t* q = A.seq_letters + A.seq_off[s];
    const int L = (int)(A.seq_off[s + 1] - A.seq_off[s]);
    const int len = L - 1;
    ShuffleSeq S;
    S.len = (uint32_t)L; S.out_base = 0;
    for (int i = 0; i < kMaxLen + 4; ++i) S.seq[i] = (i < L) ? q[i] : 0;
    int count[26], seen[26], n_keys = 0;
    for (int c = 0; c < 26; ++c) { count[c] = 0; seen[c] = -1; }
    for (int i = 0; i < len; ++i) { int c = q[i]; if (count[c]++ == 0) seen[c] = n_keys++; }
    int cap = 16;
    while (n_keys * 4 > cap * 3) cap *= 2;
    uint32_t key[26]; int ne = 0;
    for (int c = 0; c < 26; ++c) if (count[c]) key[ne++] = ((uint32_t)count[c] << 24) | ((uint32_t)((c + 'A') & (cap - 1)) << 16) | ((uint32_t)seen[c] << 8) | (uint32_t)c;
    for (int i = 1; i < ne; ++i) { uint32_t k = key[i]; int j = i - 1; while (j >= 0 && key[j] > k) { key[j + 1] = key[j]; --j; } key[j + 1] = k; }
    for (int g = 0; g < 32; ++g) { S.group_code[g] = 0; S.group_count[g] = 0; }
    S.n_groups = (uint32_t)ne;
    long long nComb = 1; int rn = len;
    for (int g = 0; g < ne; ++g) {
        int c = (int)(key[g] & 0xFF), cnt = (int)(key[g] >> 24);
        S.group_code[g] = (uint8_t)c; S.group_count[g] = (uint8_t)cnt;
        unsigned long long ch = choose_u64(rn, cnt);
        if (nComb >= 10000000) nComb = 10000000; else nComb = (long long)((unsigned long long)nComb * ch);     // :402-418
        rn -= cnt;
    }
    int num = (long long)n_random > nComb ? (int)nComb : n_random;                                              // :424
    S.num = num < 0 ? 0u : (uint32_t)num;
    plans[s] = S;
}

// Compacts the flagged sequences / forms and lays out their shuffle rows (exclusive scans).  Every CTA owns kLayoutThreads entries:
// it first sums what the entries before its own contribute (a strided read of flags and counts: a few thousand words), then scans
// its own — no CTA waits for another (as ONE CTA walking all chunks this took 0.7 ms at 10 k sequences / 15 k forms).
constexpr int kLayoutThreads = 1024;
struct Tri { uint32_t n, rows; unsigned long long units; };
struct AddTri { __device__ Tri operator()(const Tri& a, const Tri& b) const { return Tri{a.n + b.n, a.rows + b.rows, a.units + b.units}; } };
__global__ void __launch_bounds__(kLayoutThreads)
k7_layout_seqs(const uint8_t* __restrict__ seq_flag, const ShuffleSeq* __restrict__ plans, uint32_t n_seq, uint32_t* __restrict__ seq_slot,
               ShuffleSeq* __restrict__ seqs, ShuffleTotals* __restrict__ totals) {
    typedef cub::BlockScan<uint2, kLayoutThreads> Scan;
    typedef cub::BlockReduce<uint2, kLayoutThreads> Reduce;
    __shared__ union { typename Scan::TempStorage scan; typename Reduce::TempStorage red; } tmp;
    __shared__ uint2 carry;
    struct Add { __device__ uint2 operator()(const uint2& a, const uint2& b) const { return make_uint2(a.x + b.x, a.y + b.y); } };
    const uint32_t base = blockIdx.x * kLayoutThreads;
    uint2 before = make_uint2(0u, 0u);
    for (uint32_t s = threadIdx.x; s < base; s += kLayoutThreads) if (seq_flag[s]) { before.x += 1u; before.y += plans[s].num; }
    before = Reduce(tmp.red).Reduce(before, Add());
    if (threadIdx.x == 0) carry = before;
    __syncthreads();
    const uint2 c = carry;
    const uint32_t s = base + threadIdx.x;
    const bool f = s < n_seq && seq_flag[s];
    uint2 v = make_uint2(f ? 1u : 0u, f ? plans[s].num : 0u), ex, tot;
    Scan(tmp.scan).ExclusiveScan(v, ex, make_uint2(0u, 0u), Add(), tot);
    if (s < n_seq) seq_slot[s] = c.x + ex.x;
    if (f) { ShuffleSeq S = plans[s]; S.out_base = c.y + ex.y; seqs[c.x + ex.x] = S; }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { totals->n_seqs = c.x + tot.x; totals->raw_rows = c.y + tot.y; }
}
// forms: slot, first shuffle row, first cell-list unit (k_cells.cuh)
__global__ void __launch_bounds__(kLayoutThreads)
k7_layout_forms(const uint32_t* __restrict__ form_flag, const ShuffleSeq* __restrict__ plans, TargetAux A, const uint32_t* __restrict__ seq_slot,
                uint32_t* __restrict__ form_slot, uint32_t* __restrict__ row_base, ShuffleForm* __restrict__ forms, ShuffleTotals* __restrict__ totals) {
    typedef cub::BlockScan<Tri, kLayoutThreads> ScanTri;
    typedef cub::BlockReduce<Tri, kLayoutThreads> ReduceTri;
    __shared__ union { typename ScanTri::TempStorage scan; typename ReduceTri::TempStorage red; } tmp;
    __shared__ Tri carry;
    auto entry = [&](uint32_t fi, uint32_t& nch_out, int32_t& sq_out, uint32_t& len_out, uint32_t& chunks_out) -> Tri {
        const uint32_t nch = fi < A.n_forms ? form_flag[fi] : 0u;
        const bool f = nch != 0u;
        const int32_t sq = f ? A.seq_of_form[fi] : 0;
        const uint32_t num = f ? plans[sq].num : 0u, len = f ? plans[sq].len : 0u;
        const uint32_t lch = min(nch, (uint32_t)kListMaxCharge), chunks = f ? list_chunks_side((int)len, (int)lch) : 0u;
        nch_out = lch; sq_out = sq; len_out = len; chunks_out = chunks;
        return Tri{f ? 1u : 0u, num, (unsigned long long)num * (2ull * chunks)};
    };
    const uint32_t base = blockIdx.x * kLayoutThreads;
    Tri before{0u, 0u, 0ull};
    for (uint32_t fi = threadIdx.x; fi < base; fi += kLayoutThreads) { uint32_t a, l, ch; int32_t q; const Tri t = entry(fi, a, q, l, ch); before = AddTri()(before, t); }
    before = ReduceTri(tmp.red).Reduce(before, AddTri());
    if (threadIdx.x == 0) carry = before;
    __syncthreads();
    const Tri c = carry;
    const uint32_t fi = base + threadIdx.x;
    uint32_t lch, len, chunks; int32_t sq;
    const Tri v = entry(fi, lch, sq, len, chunks);
    Tri ex, tot;
    ScanTri(tmp.scan).ExclusiveScan(v, ex, Tri{0u, 0u, 0ull}, AddTri(), tot);
    if (fi < A.n_forms) { form_slot[fi] = c.n + ex.n; row_base[fi] = c.rows + ex.rows; }
    if (v.n) {
        ShuffleForm F;
        F.seq_slot = seq_slot[sq]; F.row_base = c.rows + ex.rows; F.n_mods = A.form_mods[fi].n_mods;
        for (int i = 0; i < PQ_MAX_FORM_MODS; ++i) F.mod[i] = A.form_mods[fi].mod[i];
        F.len = len; F.list_nch = lch; F.list_chunks = chunks; F.list_rows = v.rows; F.list_base = c.units + ex.units;
        forms[c.n + ex.n] = F;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { totals->n_forms = c.n + tot.n; totals->out_rows = c.rows + tot.rows; totals->list_units = c.units + tot.units; }       // totals->n_valid: counted by k7_flag_valid
}

__global__ void k7_psm_form_keys(const DevPsm* __restrict__ psms, const uint32_t* __restrict__ list, uint32_t n, uint32_t* __restrict__ keys) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = (uint32_t)psms[list[i]].form;
}

// one job per still-valid PSM (StatisticScoreWorker), ordered by form so consecutive CTAs share shuffle rows in L2
__global__ void k7_step4_jobs(const DevPsm* __restrict__ psms, const uint32_t* __restrict__ order, uint32_t n, const uint32_t* __restrict__ form_slot,
                              const ShuffleForm* __restrict__ forms, const uint32_t* __restrict__ kept_count, Job* __restrict__ jobs) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const DevPsm& p = psms[order[k]];
    const ShuffleForm& F = forms[form_slot[p.form]];
    Job J;
    J.blob = p.blob; J.cand_begin = F.row_base; J.cand_count = kept_count[F.seq_slot]; J.out = k; J.mass = p.exp_mass; J.ref_score = p.score;
    J.charge = p.charge; J.iso = p.iso; J.spectrum = p.spectrum_sorted; J.aux = p.form;
    jobs[k] = J;
}

__global__ void k7_step4_apply(const Job* __restrict__ jobs, const JobResult* __restrict__ res, const uint32_t* __restrict__ order, uint32_t n,
                               DevPsm* __restrict__ psms) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int32_t N = (int32_t)jobs[k].cand_count;
    if (N < 1) return;           // the reference exits with "There is no random peptide!" (StatisticScoreWorker.java:59-62)
    DevPsm& p = psms[order[k]];
    p.n_random = (int32_t)res[k].n_better; p.total_random = N;
    p.p_value = __ddiv_rn(__dmul_rn(1.0, (double)(p.n_random + 1)), (double)(N + 1));                             // :56-58
}

__device__ __forceinline__ double tol_ppm_of(const DevPsm& p) { return __ddiv_rn(__dmul_rn(1.0e6, __dsub_rn(p.pep_mass, p.exp_mass)), p.pep_mass); }

// Order of the texts "peptide|modification" of two target forms (RankPSMCompare's last key, pg/RankPSM.java:139-147; the
// modification text is ModificationDB.getModificationString, "id@site[%.4f]" joined by ';', "-" without modifications).
// Compared without building the strings: the sequences letter by letter — where one is a proper prefix of the other its '|'
// (0x7C) meets a letter, so the SHORTER one is greater —, then the modification lists piece by piece.  Every piece ends in
// the only ']' it contains, so no piece is a prefix of another and the first differing piece decides by its own text order
// (piece_rank, from the host); a list that ends first is a proper prefix and sorts first; "-" sorts before every piece.
__device__ __forceinline__ int form_text_cmp(const TargetAux& A, int32_t fa, int32_t fb) {
    if (fa == fb) return 0;
    const int32_t qa = A.seq_of_form[fa], qb = A.seq_of_form[fb];
    if (qa != qb) {
        const uint8_t* x = A.seq_letters + A.seq_off[qa]; const uint8_t* y = A.seq_letters + A.seq_off[qb];
        const uint32_t lx = A.seq_off[qa + 1] - A.seq_off[qa], ly = A.seq_off[qb + 1] - A.seq_off[qb], m = min(lx, ly);
        for (uint32_t i = 0; i < m; ++i) if (x[i] != y[i]) return x[i] < y[i] ? -1 : 1;
        if (lx != ly) return lx > ly ? -1 : 1;
    }
    const FormMods& ma = A.form_mods[fa]; const FormMods& mb = A.form_mods[fb];
    const uint32_t m = min(ma.n_mods, mb.n_mods);
    for (uint32_t i = 0; i < m; ++i) {
        const uint16_t ra = A.piece_rank[(uint32_t)ma.mod[i] * 64u + ma.site[i]], rb = A.piece_rank[(uint32_t)mb.mod[i] * 64u + mb.site[i]];
        if (ra != rb) return ra < rb ? -1 : 1;
    }
    if (ma.n_mods != mb.n_mods) return ma.n_mods < mb.n_mods ? -1 : 1;
    return 0;
}

// a sorts before b (RankPSMCompare: score desc, p-value asc, |tol_ppm| asc, "peptide|modification" asc)
__device__ __forceinline__ int rank_cmp(const DevPsm& a, const DevPsm& b, const TargetAux& A) {
    if (a.score != b.score) return a.score > b.score ? -1 : 1;
    if (a.p_value != b.p_value) return a.p_value < b.p_value ? -1 : 1;
    double ta = fabs(tol_ppm_of(a)), tb = fabs(tol_ppm_of(b));
    if (ta != tb) return ta < tb ? -1 : 1;
    return form_text_cmp(A, a.form, b.form);
}

__global__ void k7_rank(DevPsm* __restrict__ psms, uint32_t n, const uint32_t* __restrict__ seg_start, TargetAux A) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevPsm me = psms[i];
    const uint32_t a = seg_start[me.seg], b = seg_start[me.seg + 1];
    int32_t rank = 1;
    for (uint32_t j = a; j < b; ++j) {
        if (j == i) continue;
        int c = rank_cmp(psms[j], me, A);
        if (c < 0 || (c == 0 && j < i)) ++rank;          // stable: equal keys keep table order
    }
    bool pv = (me.len <= 8 && me.p_value <= 0.05) || (me.len > 8 && me.p_value <= 0.01);       // CParameter.java:913-940
    bool ok = pv && me.n_db == 0 && rank == 1;
    if (me.n_ptm >= 0) ok = ok && me.n_ptm == 0;
    psms[i].rank = rank;                                  // rank/confident are not read by rank_cmp: no race with other threads
    psms[i].confident = ok ? 1 : 0;
}

__global__ void k7_export(const DevPsm* __restrict__ psms, const uint32_t* __restrict__ out_perm, uint32_t n, pq_psm* __restrict__ out) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const DevPsm& p = psms[out_perm[j]];
    pq_psm q;
    q.spectrum = p.spectrum; q.form = p.form; q.score = p.score; q.exp_mass = p.exp_mass; q.pep_mass = p.pep_mass; q.tol_ppm = tol_ppm_of(p);
    q.isotope_error = p.iso; q.count_b = p.ca; q.count_y = p.cb; q.n_db = p.n_db; q.total_db = p.total_db; q.n_random = p.n_random;
    if (p.rank >= 2) { q.n_db = -1; q.total_db = -1; }       // psm_rank.txt: RankPSM.rankPSMs rewrites both for rank >= 2 (pg/RankPSM.java:89-92)
    q.total_random = p.total_random; q.valid = p.valid; q.p_value = p.p_value; q.rank = p.rank; q.n_ptm = p.n_ptm; q.confident = p.confident; q.charge = p.charge;
    out[j] = q;
}

// PeptideSearchMT.add_delta_score (pg/PeptideSearchMT.java:1365-1475) in export order: score minus the best detail.txt row
// of the spectrum (= the best reference competitor when its score > 0, DBSearchWorker.java:93-98), score minus the best
// ptm.txt row of the spectrum (sorted caller-spectrum keys, binary search); the plain score where there is no row.
__global__ void k7_delta_scores(const DevPsm* __restrict__ psms, const uint32_t* __restrict__ out_perm, uint32_t n, const SegAgg* __restrict__ agg,
                                const int32_t* __restrict__ mod_key, const double* __restrict__ mod_best, uint32_t n_mod,
                                double* __restrict__ ref_delta, double* __restrict__ mod_delta) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const DevPsm& p = psms[out_perm[j]];
    double r = p.score, m = p.score;
    if (agg) { const double b = agg[p.seg].best; if (b > 0.0) r = __dsub_rn(p.score, b); }
    uint32_t lo = 0, hi = n_mod;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (mod_key[mid] < p.spectrum) lo = mid + 1; else hi = mid; }
    if (lo < n_mod && mod_key[lo] == p.spectrum) m = __dsub_rn(p.score, mod_best[lo]);
    ref_delta[j] = r; mod_delta[j] = m;
}

// PeptideSearchMT.export_targeted_psm_status (pg/PeptideSearchMT.java:1484-1589): per targeted pair the best-ranked PSM row of
// (peptide sequence, spectrum), then TargetedPSM.get_psm_status (pg/TargetedPSM.java:29-46) on that row
__global__ void k7_targeted_best(const DevPsm* __restrict__ psms, uint32_t n, TargetedView V, unsigned long long* __restrict__ best) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DevPsm& p = psms[i];
    const uint32_t seq = V.canon[V.seq_of_form[p.form]];
    if (!V.seq_listed[seq]) return;
    const uint64_t key = ((uint64_t)seq << 32) | (uint32_t)p.spectrum;
    uint32_t lo = 0, hi = V.n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (V.keys[mid] < key) lo = mid + 1; else hi = mid; }
    if (lo < V.n && V.keys[lo] == key) atomicMin(&best[lo], ((unsigned long long)(uint32_t)p.rank << 32) | i);
}
__global__ void k7_targeted_type(const DevPsm* __restrict__ psms, TargetedView V, const unsigned long long* __restrict__ best, int32_t* __restrict__ type) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= V.n) return;
    const unsigned long long b = best[k];
    if (b == ~0ull) { type[k] = V.seen[k] ? PQ_PSM_LOW_SCORE : PQ_PSM_NO_SPECTRUM_MATCH; return; }
    const DevPsm& p = psms[(uint32_t)b];
    const bool pv = (p.len <= 8 && p.p_value <= 0.05) || (p.len > 8 && p.p_value <= 0.01);       // CParameter.passed_p_value_validation
    type[k] = p.rank >= 2 ? PQ_PSM_NOT_TOP_RANK : (p.n_db >= 1 ? PQ_PSM_BETTER_REF_PEP : (!pv ? PQ_PSM_HIGH_P_VALUE : (p.n_ptm >= 1 ? PQ_PSM_BETTER_REF_PEP_WITH_MOD : PQ_PSM_CONFIDENT)));
}

__global__ void k7_set_nptm(DevPsm* __restrict__ psms, uint32_t n, const int32_t* __restrict__ n_ptm) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) psms[i].n_ptm = n_ptm[i];
}

inline unsigned blocks(uint32_t n, int per = 256) { return (n + per - 1) / per; }

// ---- merging the parts of a search that ran while the spectra were still arriving (pq_load_spectra, search_at_load) ----
__global__ void k7_merge_psms(const DevPsm* __restrict__ in, uint32_t n, uint32_t seg_off, DevPsm* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    DevPsm p = in[i]; p.seg += seg_off; out[i] = p;
}
__global__ void k7_merge_segs(const uint32_t* __restrict__ seg_start_in, const SegAgg* __restrict__ agg_in, uint32_t n_seg, uint32_t psm_off, uint32_t row_off,
                              uint32_t* __restrict__ seg_start_out, SegAgg* __restrict__ agg_out, uint32_t end_value, int last) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_seg) return;
    seg_start_out[i] = seg_start_in[i] + psm_off;
    if (agg_in) { SegAgg g = agg_in[i]; if (g.best_row != 0xffffffffu) g.best_row -= row_off; agg_out[i] = g; }
    if (last && i == n_seg - 1) seg_start_out[n_seg] = end_value;
}
__global__ void k7_merge_hits(const Hit* __restrict__ in, uint32_t n, uint32_t row_off, Hit* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Hit h = in[i]; h.cand -= row_off; out[i] = h;
}
// Step-2 jobs of a search at load: key = (part of the job's spectrum, charge, first target row, rows); rank_in_upload[s] = position of
// sorted spectrum s in the order the upload brings the spectra over, part = (rank + shift) / part_spectra (shift > 0: a shorter first part)
__global__ void k7_part_job_keys(const Job* __restrict__ jobs, uint32_t n, const uint32_t* __restrict__ rank_in_upload, uint32_t part_spectra, uint32_t shift,
                                 uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Job& J = jobs[i];
    uint64_t part = (rank_in_upload[J.spectrum] + shift) / part_spectra, z = (uint64_t)min(max(J.charge, 0), 15);
    key[i] = (part << 52) | (z << 48) | ((uint64_t)(J.cand_begin & 0xFFFFFFFFFFull) << 8) | (uint64_t)min(J.cand_count, 255u);
    idx[i] = i;
}
__global__ void k7_part_bounds(const uint64_t* __restrict__ sorted_key, uint32_t n, uint32_t n_parts, uint32_t* __restrict__ first /* [n_parts + 1] */) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n_parts) return;
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if ((sorted_key[mid] >> 52) < (uint64_t)p) lo = mid + 1; else hi = mid; }
    first[p] = lo;
}
// candidate rows of the Step-2 jobs of every part (sizes the parts' hit buffers)
__global__ void k7_part_cand(const Job* __restrict__ jobs, uint32_t n, const uint64_t* __restrict__ sorted_key, unsigned long long* __restrict__ cand /* [n_parts] zeroed */) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&cand[(uint32_t)(sorted_key[i] >> 52)], (unsigned long long)jobs[i].cand_count);
}
// rows of a mass-sorted table inside the open window (lo, hi): [a, b)
__global__ void k7_window_rows(const double* __restrict__ mass, uint32_t n, double lo, double hi, uint32_t* __restrict__ ab) {
    if (blockIdx.x || threadIdx.x) return;
    uint32_t l = 0, h = n;
    while (l < h) { uint32_t m = (l + h) >> 1; if (mass[m] <= lo) l = m + 1; else h = m; }
    ab[0] = l;
    l = 0; h = n;
    while (l < h) { uint32_t m = (l + h) >> 1; if (mass[m] < hi) l = m + 1; else h = m; }
    ab[1] = l;
}
__global__ void k7_scatter_rank(const int32_t* __restrict__ list, uint32_t n, uint32_t* __restrict__ rank) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) rank[list[i]] = i;
}
__global__ void k7_fill_u8(uint8_t* p, uint32_t n, uint8_t v) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }
__global__ void k7_fill_u32(uint32_t* p, uint32_t n, uint32_t v) { uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

}  // namespace

int launch_merge_psms(const DevPsm* in, uint32_t n, uint32_t seg_off, DevPsm* out, cudaStream_t st) {
    if (!n) return 0;
    k7_merge_psms<<<blocks(n), 256, 0, st>>>(in, n, seg_off, out);
    return 1;
}
int launch_merge_segs(const uint32_t* seg_start_in, const SegAgg* agg_in, uint32_t n_seg, uint32_t psm_off, uint32_t row_off, uint32_t* seg_start_out, SegAgg* agg_out,
                      uint32_t end_value, int last, cudaStream_t st) {
    if (!n_seg) return 0;
    k7_merge_segs<<<blocks(n_seg), 256, 0, st>>>(seg_start_in, agg_in, n_seg, psm_off, row_off, seg_start_out, agg_out, end_value, last);
    return 1;
}
int launch_merge_hits(const Hit* in, uint32_t n, uint32_t row_off, Hit* out, cudaStream_t st) {
    if (!n) return 0;
    k7_merge_hits<<<blocks(n), 256, 0, st>>>(in, n, row_off, out);
    return 1;
}
int launch_part_job_keys(const Job* jobs, uint32_t n, const uint32_t* rank_in_upload, uint32_t part_spectra, uint32_t shift, uint64_t* key, uint32_t* idx, cudaStream_t st) {
    if (!n) return 0;
    k7_part_job_keys<<<blocks(n), 256, 0, st>>>(jobs, n, rank_in_upload, part_spectra, shift, key, idx);
    return 1;
}
int launch_part_bounds(const uint64_t* sorted_key, uint32_t n, uint32_t n_parts, uint32_t* first, cudaStream_t st) {
    k7_part_bounds<<<blocks(n_parts + 1), 256, 0, st>>>(sorted_key, n, n_parts, first);
    return 1;
}
int launch_part_cand(const Job* jobs, uint32_t n, const uint64_t* sorted_key, unsigned long long* cand, cudaStream_t st) {
    if (!n) return 0;
    k7_part_cand<<<blocks(n), 256, 0, st>>>(jobs, n, sorted_key, cand);
    return 1;
}
int launch_window_rows(const double* mass, uint32_t n, double lo, double hi, uint32_t* ab, cudaStream_t st) {
    k7_window_rows<<<1, 32, 0, st>>>(mass, n, lo, hi, ab);
    return 1;
}
int launch_scatter_rank(const int32_t* list, uint32_t n, uint32_t* rank, cudaStream_t st) {
    if (!n) return 0;
    k7_scatter_rank<<<blocks(n), 256, 0, st>>>(list, n, rank);
    return 1;
}
int launch_fill_u8(uint8_t* p, uint32_t n, uint8_t v, cudaStream_t st) { if (!n) return 0; k7_fill_u8<<<blocks(n), 256, 0, st>>>(p, n, v); return 1; }
int launch_fill_u32(uint32_t* p, uint32_t n, uint32_t v, cudaStream_t st) { if (!n) return 0; k7_fill_u32<<<blocks(n), 256, 0, st>>>(p, n, v); return 1; }
bool build_out_perm(const DevPsm* psms, uint32_t n, DevBuf& key_a, DevBuf& key_b, DevBuf& idx_a, DevBuf& tmp, uint32_t* out_perm, cudaStream_t st, std::string& err) {
    if (!n) return true;
    RB(key_a.reserve((size_t)n * 8)); RB(key_b.reserve((size_t)n * 8)); RB(idx_a.reserve((size_t)n * 4));
    size_t b1 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, key_a.as<uint64_t>(), key_b.as<uint64_t>(), idx_a.as<uint32_t>(), out_perm, (int)n, 0, 64, st);
    RB(tmp.reserve(b1 + 64));
    k7_out_keys<<<blocks(n), 256, 0, st>>>(psms, n, key_a.as<uint64_t>(), idx_a.as<uint32_t>());
    RB(cub::DeviceRadixSort::SortPairs(tmp.p, b1, key_a.as<uint64_t>(), key_b.as<uint64_t>(), idx_a.as<uint32_t>(), out_perm, (int)n, 0, 64, st));
    RB(cudaGetLastError());
    return true;
}

bool psm_table_from_hits(const Hit* hits, uint32_t nh, const SpectraView& sp, const double* t_mass, const uint8_t* t_rows, const TargetAux& A,
                         DevBuf& key_a, DevBuf& key_b, DevBuf& idx_a, DevBuf& idx_b, DevBuf& head, DevBuf& tmp,
                         DevPsm* psms, uint32_t* seg_start, uint32_t* d_nseg, uint32_t* out_perm, cudaStream_t st, int* launches, std::string& err) {
    if (nh == 0) return true;
    RB(key_a.reserve((size_t)nh * 8)); RB(key_b.reserve((size_t)nh * 8)); RB(idx_a.reserve((size_t)nh * 4)); RB(idx_b.reserve((size_t)nh * 4)); RB(head.reserve((size_t)nh * 4));
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, key_a.as<uint64_t>(), key_b.as<uint64_t>(), idx_a.as<uint32_t>(), idx_b.as<uint32_t>(), (int)nh, 0, 64, st);
    cub::DeviceScan::InclusiveSum(nullptr, b2, head.as<uint32_t>(), idx_a.as<uint32_t>(), (int)nh, st);
    RB(tmp.reserve(std::max(b1, b2) + 64));
    k7_hit_keys<<<blocks(nh), 256, 0, st>>>(hits, nh, A.form_of_row, key_a.as<uint64_t>(), idx_a.as<uint32_t>());
    RB(cub::DeviceRadixSort::SortPairs(tmp.p, b1, key_a.as<uint64_t>(), key_b.as<uint64_t>(), idx_a.as<uint32_t>(), idx_b.as<uint32_t>(), (int)nh, 0, 64, st));
    k7_build_psms<<<blocks(nh), 256, 0, st>>>(hits, idx_b.as<uint32_t>(), nh, sp, t_mass, t_rows, A, psms, head.as<uint32_t>());
    RB(cub::DeviceScan::InclusiveSum(tmp.p, b2, head.as<uint32_t>(), idx_a.as<uint32_t>(), (int)nh, st));
    k7_set_segments<<<blocks(nh), 256, 0, st>>>(head.as<uint32_t>(), idx_a.as<uint32_t>(), nh, psms, seg_start, d_nseg);
    if (out_perm) {
        k7_out_keys<<<blocks(nh), 256, 0, st>>>(psms, nh, key_a.as<uint64_t>(), idx_a.as<uint32_t>());
        RB(cub::DeviceRadixSort::SortPairs(tmp.p, b1, key_a.as<uint64_t>(), key_b.as<uint64_t>(), idx_a.as<uint32_t>(), out_perm, (int)nh, 0, 64, st));
    }
    RB(cudaGetLastError());
    *launches += 4 + 3 * 5 + 2;
    return true;
}

int launch_job_counters(const Job* jobs, const JobResult* res, uint32_t n_jobs, const uint32_t* sp_npeaks, int which, DevCounters* c, cudaStream_t st,
                        const uint8_t* skip_job) {
    if (!n_jobs) return 0;
    k7_job_counters<<<blocks(n_jobs), 256, 0, st>>>(jobs, res, n_jobs, sp_npeaks, which, c, skip_job);
    return 1;
}
int launch_step3_jobs(const DevPsm* psms, const uint32_t* seg_start, uint32_t n_seg, const double* r_mass, uint32_t n_rows, const WindowConfig& W,
                      const double* sp_mass, Job* jobs, SegAgg* agg, unsigned long long* total_cand, cudaStream_t st) {
    uint32_t n = n_seg * (uint32_t)W.n_isotope;
    if (!n) return 0;
    k7_step3_jobs<<<blocks(n), 256, 0, st>>>(psms, seg_start, n_seg, r_mass, n_rows, W, sp_mass, jobs, agg, total_cand);
    return 1;
}
int launch_step3_aggregate(const JobResult* res, uint32_t n_seg, int n_iso, const uint32_t* seg_start, DevPsm* psms, SegAgg* agg, cudaStream_t st) {
    if (!n_seg) return 0;
    k7_step3_aggregate<<<blocks(n_seg), 256, 0, st>>>(res, n_seg, n_iso, seg_start, psms, agg);
    return 1;
}
int launch_comp3_rows(const Hit* hits, uint32_t nh, const SegAgg* agg, uint32_t n_seg, const uint32_t* seg_start, const DevPsm* psms,
                      const int32_t* caller_index, const double* r_mass, pq_competitor* rows, uint64_t* keys, uint32_t* idx, uint32_t* count, cudaStream_t st) {
    int k = 0;
    if (nh) { k7_comp3_hits<<<blocks(nh), 256, 0, st>>>(hits, nh, caller_index, r_mass, rows, keys, idx); ++k; }
    if (n_seg) { k7_comp3_best<<<blocks(n_seg), 256, 0, st>>>(agg, n_seg, seg_start, psms, r_mass, nh, rows, keys, idx, count); ++k; }
    return k;
}
int launch_gather_comp(const pq_competitor* in, const uint32_t* idx, uint32_t n, pq_competitor* out, cudaStream_t st) {
    if (!n) return 0;
    k7_gather_comp<<<blocks(n), 256, 0, st>>>(in, idx, n, out);
    return 1;
}
int launch_flag_valid(const DevPsm* psms, uint32_t n, const int32_t* seq_of_form, uint8_t* seq_flag, uint32_t* form_flag, uint8_t* valid_flag,
                      ShuffleTotals* totals /* zeroed by the caller */, cudaStream_t st) {
    if (!n) return 0;
    k7_flag_valid<<<blocks(n), 256, 0, st>>>(psms, n, seq_of_form, seq_flag, form_flag, valid_flag, &totals->n_valid);
    return 1;
}
int launch_valid_flags(const DevPsm* psms, uint32_t n, uint8_t* valid_flag, uint32_t* n_valid /* zeroed by the caller */, cudaStream_t st) {
    if (!n) return 0;
    k7_valid_flags<<<blocks(n), 256, 0, st>>>(psms, n, valid_flag, n_valid);
    return 1;
}
int launch_shuffle_plans(const uint8_t* seq_flag, const TargetAux& A, int32_t n_random, ShuffleSeq* plans, cudaStream_t st) {
    if (!A.n_seq) return 0;
    k7_shuffle_plans<<<blocks(A.n_seq, 64), 64, 0, st>>>(seq_flag, A, n_random, plans);
    return 1;
}
int launch_shuffle_layout(const uint8_t* seq_flag, const uint32_t* form_flag, const uint8_t* valid_flag, uint32_t n_psm, const ShuffleSeq* plans, const TargetAux& A,
                          uint32_t* seq_slot, uint32_t* form_slot, uint32_t* row_base, ShuffleSeq* seqs, ShuffleForm* forms, ShuffleTotals* totals, cudaStream_t st) {
    (void)valid_flag; (void)n_psm;
    if (A.n_seq) k7_layout_seqs<<<(A.n_seq + kLayoutThreads - 1) / kLayoutThreads, kLayoutThreads, 0, st>>>(seq_flag, plans, A.n_seq, seq_slot, seqs, totals);
    if (A.n_forms) k7_layout_forms<<<(A.n_forms + kLayoutThreads - 1) / kLayoutThreads, kLayoutThreads, 0, st>>>(form_flag, plans, A, seq_slot, form_slot, row_base, forms, totals);
    return 1;
}
int launch_psm_form_keys(const DevPsm* psms, const uint32_t* list, uint32_t n, uint32_t* keys, cudaStream_t st) {
    if (!n) return 0;
    k7_psm_form_keys<<<blocks(n), 256, 0, st>>>(psms, list, n, keys);
    return 1;
}
int launch_step4_jobs(const DevPsm* psms, const uint32_t* order, uint32_t n_valid, const uint32_t* form_slot, const ShuffleForm* forms, const uint32_t* kept_count,
                      Job* jobs, cudaStream_t st) {
    if (!n_valid) return 0;
    k7_step4_jobs<<<blocks(n_valid), 256, 0, st>>>(psms, order, n_valid, form_slot, forms, kept_count, jobs);
    return 1;
}
int launch_step4_apply(const Job* jobs, const JobResult* res, const uint32_t* order, uint32_t n_valid, DevPsm* psms, cudaStream_t st) {
    if (!n_valid) return 0;
    k7_step4_apply<<<blocks(n_valid), 256, 0, st>>>(jobs, res, order, n_valid, psms);
    return 1;
}
int launch_rank(DevPsm* psms, uint32_t n, const uint32_t* seg_start, const TargetAux& A, cudaStream_t st) {
    if (!n) return 0;
    k7_rank<<<blocks(n), 256, 0, st>>>(psms, n, seg_start, A);
    return 1;
}
int launch_export_psms(const DevPsm* psms, const uint32_t* out_perm, uint32_t n, pq_psm* out, cudaStream_t st) {
    if (!n) return 0;
    k7_export<<<blocks(n), 256, 0, st>>>(psms, out_perm, n, out);
    return 1;
}
int launch_delta_scores(const DevPsm* psms, const uint32_t* out_perm, uint32_t n, const SegAgg* agg, const int32_t* mod_key, const double* mod_best,
                        uint32_t n_mod, double* ref_delta, double* mod_delta, cudaStream_t st) {
    if (!n) return 0;
    k7_delta_scores<<<blocks(n), 256, 0, st>>>(psms, out_perm, n, agg, mod_key, mod_best, n_mod, ref_delta, mod_delta);
    return 1;
}
int launch_targeted_status(const DevPsm* psms, uint32_t n_psm, const TargetedView& V, unsigned long long* best, int32_t* type_sorted, cudaStream_t st) {
    if (!V.n) return 0;
    cudaMemsetAsync(best, 0xFF, (size_t)V.n * 8, st);
    int k = 1;
    if (n_psm) { k7_targeted_best<<<blocks(n_psm), 256, 0, st>>>(psms, n_psm, V, best); ++k; }
    k7_targeted_type<<<blocks(V.n), 256, 0, st>>>(psms, V, best, type_sorted);
    return k;
}
int launch_set_nptm(DevPsm* psms, uint32_t n, const int32_t* n_ptm, cudaStream_t st) {
    if (!n) return 0;
    k7_set_nptm<<<blocks(n), 256, 0, st>>>(psms, n, n_ptm);
    return 1;
}

}  // namespace pq

// pq_api.cu — the C ABI of libpepquery_b200.so (include/pepquery_b200.h): context, device residency, step drivers.
// No CPU fallback: every scoring entry point launches CUDA kernels or fails.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <cub/cub.cuh>

#include "pq_host.h"
#include "k_cells.cuh"

using namespace pq;

namespace {

thread_local std::string g_create_error;

// PQ_DEBUG=1: host wall-clock marks inside the step drivers (stderr)
struct Marks {
    const char* name; bool on; std::chrono::steady_clock::time_point t0, last;
    explicit Marks(const char* n) : name(n), on(getenv("PQ_DEBUG") != nullptr) { t0 = last = std::chrono::steady_clock::now(); }
    void mark(const char* what) {
        if (!on) return;
        auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[pq] %s: %-28s %8.3f ms\n", name, what, std::chrono::duration<double, std::milli>(now - last).count());
        last = now;
    }
    ~Marks() { if (on) fprintf(stderr, "[pq] %s: total %8.3f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); }
};

struct PeptideTable {               // a mass-sorted candidate table resident on the device
    std::vector<std::string> seqs;  // parent sequences
    std::vector<FormRec> forms;     // mass ascending
    DevBuf rows, mass, pred, mask;
    bool have_mask = false;
    DevBuf lad, lad_len; uint32_t lad_stride = 0;   // precomputed fragment ladders of the rows (k1_ladder.cu), built before the first Step 3 over this table
    bool have_lad = false;
    std::vector<double> h_mass;     // host copy for range searches
    RefDeviceTable dev;             // device-built table (reference only)
    bool device_built = false, materialized = true;
    size_t n_rows = 0;
    size_t n() const { return device_built ? n_rows : forms.size(); }
};

struct TargetOrder { std::vector<uint32_t> row_of_form, form_of_row; };   // public (generation) order <-> device (mass) order

struct HostPsm {           // host mirror of a DevPsm (Step 5 only)
    int32_t spectrum_sorted, spectrum, form, len; uint32_t blob;
    double score; int32_t charge, n_db, rank, n_ptm; double p_value;
};

}  // namespace

struct pq_ctx {
    pq_params P;
    Codes codes;
    std::string err;
    cudaStream_t st = nullptr, st_copy = nullptr, st_aux = nullptr;
    std::vector<cudaEvent_t> chunk_ev;       // one per upload chunk of pq_load_spectra
    cudaEvent_t meta_ev = nullptr;           // precursor arrays of pq_load_spectra on the device
    cudaEvent_t targets_ev = nullptr;        // target table of this load on the device (the early shuffle generation waits for it)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, uev0 = nullptr, uev1 = nullptr;
    void* stage_host = nullptr; cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    pq_counters cnt;
    struct PendingTimer { cudaEvent_t a, b; double* acc; double* acc2; };
    std::vector<PendingTimer> timers; std::vector<cudaEvent_t> free_events;
    // constant tables
    DevBuf d_mods, d_l10, d_lnf, d_shufmods;
    std::vector<double> ln_fact; double log10_fact[130];    // [0..64] log10(n!), [65..129] log10(100 n) (upper-bound table of K1)
    // spectra
    int64_t nS = 0;                          // (virtual) spectra on the device: the caller's, plus one more per charge-less spectrum (its z = 3 reading)
    int64_t nCaller = 0, nZero = 0; uint32_t max_peaks = 0;
    DevBuf s_prec, s_charge, s_mass, s_off, s_mz, s_in, s_caller, s_npeaks, s_alias, s_vid, s_res, ld[12];
    std::vector<int32_t> h_caller; std::vector<double> h_mass; std::vector<int32_t> h_charge; std::vector<uint32_t> h_npeaks; std::vector<uint8_t> h_resident, h_alias;
    bool host_spectra_valid = false;
    // which spectra have their peaks in HBM: 0 = all (full upload); 1 = candidate-first (only those with a target in their precursor
    // window when they were loaded: s_res flags, prep_id = rank among them)
    int residency = 0; uint64_t targets_epoch = 0, loaded_epoch = 0; uint32_t n_resident = 0;
    SpectraView view() const {
        SpectraView v; v.n = nS; v.prec_mz = s_prec.as<double>(); v.charge = s_charge.as<int32_t>(); v.mass = s_mass.as<double>();
        v.off = s_off.as<uint64_t>(); v.mz = s_mz.as<double>(); v.inten = s_in.as<double>(); v.caller_index = s_caller.as<int32_t>();
        v.npeaks = s_npeaks.as<uint32_t>(); v.alias = s_alias.as<uint8_t>();
        return v;
    }
    // prepared spectra
    DevBuf blobs, blob_off, blob_len, prep_id, prep_list, need, scratch, scan_tmp, pp_blobs, pp_blob_off, pp_blob_len;
    uint32_t n_prepared = 0; uint32_t max_blob = 0;
    bool prepared_at_load = false;           // blobs were written by pq_load_spectra: blob id = virtual id (full upload) or rank among the resident spectra (candidate-first)
    DevBuf s_inv;                            // virtual id -> mass-sorted position
    DevBuf gather_list;                      // candidate-first: the resident spectra (sorted positions) in caller order
    DevBuf k0_counters;                      // work hand-out words of the per-chunk K0 launches
    DevBuf chunk_done;                       // candidate-first gather: CTAs that have finished upload chunk c
    // candidate-first from pageable caller arrays: page-locked packing ring (two chunks), its device twin, what to pack
    void* pack_host = nullptr; size_t pack_host_cap = 0; DevBuf pack_dev, pack_src, pack_cnt, pack_off; cudaEvent_t pack_ev[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> k0_ev;           // K0 of upload chunk c has finished
    DevBuf d_zero;                           // caller indices of the charge-less spectra, ascending
    // tables
    PeptideTable targets, reference;
    TargetOrder order;
    bool have_targets = false, have_reference = false, forms_expanded = false;
    bool ref_staged = false, ref_digested = false;      // pq_stage_reference: proteins on the device / digest done (valid for the current targets)
    int step_done = 0;
    // work buffers
    DevBuf jobs, jobs2, results, hits, hit_count, njobs, hits_f, z_hit, z_skip, z_keep;
    // device-resident PSM table (K7)
    DevBuf d_psm, d_seg_start, d_out_perm, d_scalars, d_counters, d_agg, hits3, d_export, d_delta, d_modkey, d_modbest;
    DevBuf k_key_a, k_key_b, k_idx_a, k_idx_b, k_head;
    uint32_t n_psm = 0, n_seg = 0, nh3 = 0, comp3_n = 0; bool comp3_ready = false;
    DevBuf c_rows, c_rows2, c_keys, c_keys2, c_idx, c_idx2;
    double psm_mass_lo = 0.0, psm_mass_hi = 0.0;
    // target-table side arrays on the device
    DevBuf t_form_of_row, t_seq_of_form, t_piece_rank, t_form_mods, t_seq_off, t_seq_letters;
    TargetDeviceTable tdev; bool target_table_ready = false; uint32_t n_target_forms = 0;
    std::vector<uint32_t> t_off; std::string t_flat;      // the target sequences back to back, as given
    TargetAux aux() const {
        TargetAux a; a.form_of_row = t_form_of_row.as<uint32_t>(); a.seq_of_form = t_seq_of_form.as<int32_t>(); a.piece_rank = t_piece_rank.as<uint16_t>();
        a.form_mods = t_form_mods.as<FormMods>(); a.seq_off = t_seq_off.as<uint32_t>(); a.seq_letters = t_seq_letters.as<uint8_t>();
        a.n_seq = (uint32_t)targets.seqs.size(); a.n_forms = n_target_forms;
        return a;
    }
    // shuffles
    DevBuf sh_seqs, sh_forms, sh_raw, sh_kidx, sh_kcnt, sh_rows, sh_pred, sh_lists;
    bool have_lists = false;
    bool shuffles_all = false; uint64_t shuffles_epoch = 0;      // shuffles exist for EVERY target form of that target list (search at load)
    // search at load: the parts of the PSM table, the superset reference table, what pq_step2/3/4 may hand back without working
    struct PartBufs {
        DevBuf jobs, results, hits, hit_count, psm, seg_start, agg, hits3, valid_flag, list, order, keys, keys2, key_a, key_b, idx_a, idx_b, head, tmp, scal;
        uint32_t n_psm = 0, n_seg = 0, nh3 = 0; int32_t s_first = 0, s_last = 0;
    };
    std::vector<PartBufs> parts;
    PeptideTable ref_super; uint32_t super_rows = 0; bool super_ready = false;
    bool pending_step[5] = {false, false, false, false, false};   // [2..4]: results of that step are already in the PSM table
    bool pending_ref = false;                                      // pq_build_reference(NULL) only has to cut the table out of the superset
    cudaStream_t st_part = nullptr, st_part2 = nullptr;      // search at load: the parts alternate between two streams
    cudaStream_t st_shuf = nullptr; cudaEvent_t shuf_ev = nullptr; bool shuf_pending = false;      // the shuffles of every target form, generated beside the upload; Step 4 of a part waits for shuf_ev
    // targeted PSM list (SpectraInput.pep2targeted_spectra)
    DevBuf tg_keys, tg_listed, tg_canon, tg_seen, tg_best, tg_type; uint32_t n_targeted = 0; std::vector<uint32_t> tg_pos;     // tg_pos[i] = position of pair i in the sorted keys
    TargetedView targeted_view() const {
        TargetedView v; memset(&v, 0, sizeof v);
        v.n = n_targeted; if (!n_targeted) return v;
        v.keys = tg_keys.as<uint64_t>(); v.seq_listed = tg_listed.as<uint8_t>(); v.canon = tg_canon.as<uint32_t>(); v.seen = tg_seen.as<uint8_t>();
        v.form_of_row = t_form_of_row.as<uint32_t>(); v.seq_of_form = t_seq_of_form.as<int32_t>(); v.caller_index = s_caller.as<int32_t>(); v.alias = s_alias.as<uint8_t>();
        return v;
    }
    DevBuf up_rank, part_first, super_seq;
    MgfDeviceWork mgf; void* mgf_host = nullptr; size_t mgf_host_cap = 0;
    const char* mgf_cache_text = nullptr; uint64_t mgf_cache_bytes = 0; int64_t mgf_cache_ns = 0; uint64_t mgf_cache_np = 0;
    DevBuf sh_plans, sh_seq_flag, sh_form_flag, sh_valid_flag, sh_seq_slot, sh_form_slot, sh_row_base, sh_totals, sh_list, sh_order, sh_keys, sh_keys2;
    bool have_shuffles = false;
    // results
    std::vector<HostPsm> psms;          // Step 5 mirror
    std::vector<pq_competitor> comp5;
};

namespace {

int32_t fail(pq_ctx* c, int32_t code, const std::string& msg) { if (c) c->err = msg; else g_create_error = msg; return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(ctx, PQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

// CUDA-event stopwatch around a group of launches.  Recording never blocks the host; the elapsed times are read when
// the counters are asked for (flush_timers).
struct Timer {
    pq_ctx* c; double* acc; double* acc2; cudaEvent_t a = nullptr, b = nullptr;
    Timer(pq_ctx* c_, double* t, double* t2 = nullptr);
    ~Timer();
};
void flush_timers(pq_ctx* c);
int32_t finalize_targets(pq_ctx* ctx);
void ensure_forms(pq_ctx* ctx);
int32_t digest_staged_reference(pq_ctx* ctx, cudaStream_t st, bool count_time = true);
int32_t generate_shuffles(pq_ctx* ctx, cudaStream_t st, bool all, ShuffleTotals& tot, bool timed, pq_counters* sink = nullptr);
// What pq_load_spectra hands to search_at_load: the Step-2 jobs its window enumeration left in ctx->jobs, the order the spectra
// arrive in (gather_list, chunks of chunk_spectra, one K0-done event per chunk) and how many chunks make a part.
struct LoadPlan {
    uint32_t n_jobs; unsigned long long cand; uint32_t n_res; uint32_t chunk_spectra; size_t n_chunks; const std::vector<cudaEvent_t>* k0_done;
    // chunks whose K0-done event a second thread has recorded so far (pageable caller arrays), null = all recorded; and its failure flag
    const std::atomic<size_t>* issued = nullptr; const std::atomic<bool>* failed = nullptr;
};

int32_t search_at_load(pq_ctx* ctx, const LoadPlan& plan);

Timer::Timer(pq_ctx* c_, double* t, double* t2) : c(c_), acc(t), acc2(t2) {
    auto get = [&]() { cudaEvent_t e; if (!c->free_events.empty()) { e = c->free_events.back(); c->free_events.pop_back(); } else cudaEventCreate(&e); return e; };
    a = get(); b = get();
    cudaEventRecord(a, c->st);
}
Timer::~Timer() {
    cudaEventRecord(b, c->st);
    c->timers.push_back({a, b, acc, acc2});
    if (c->timers.size() > 512) flush_timers(c);
}
void flush_timers(pq_ctx* c) {
    for (auto& t : c->timers) {
        cudaEventSynchronize(t.b);
        float ms = 0; cudaEventElapsedTime(&ms, t.a, t.b);
        *t.acc += ms; if (t.acc2) *t.acc2 += ms; c->cnt.ms_total += ms;
        c->free_events.push_back(t.a); c->free_events.push_back(t.b);
    }
    c->timers.clear();
}

// spectra with n_peaks <= this are never scored: min_peaks (SpectraInput.java:183), one less for a library search (n_peaks >= min_peaks)
int32_t min_peaks_eff(const pq_ctx* c) { return c->P.min_peaks - (c->P.library_mode ? 1 : 0); }
WindowConfig window_config(const pq_ctx* c, bool step2) {
    WindowConfig W; W.tol = c->P.tol; W.tol_ppm = c->P.tol_ppm; W.n_isotope = c->P.n_isotope; W.min_peaks = min_peaks_eff(c);
    for (int i = 0; i < PQ_MAX_ISOTOPE; ++i) W.isotope[i] = c->P.isotope[i];
    W.no_bucket = step2 && c->P.library_mode ? 1 : 0;
    return W;
}
ScoreConfig score_config(const pq_ctx* c, int check_window) {
    ScoreConfig s; s.itol = c->P.itol; s.tol = c->P.tol; s.tol_ppm = c->P.tol_ppm; s.method = c->P.score_method;
    s.min_score = c->P.min_score; s.fast = c->P.fast; s.hc = c->P.hc;
    s.mvh_bins = (int32_t)jround((double)(int)(2000.0 - 200.0) / (c->P.itol * 2.0));
    s.check_window = check_window;
    s.score_all = c->P.score_all_shuffles || getenv("PQ_SCORE_ALL_SHUFFLES") != nullptr;
    s.no_bucket = 0;
    s.prefetch_next = getenv("PQ_NO_BLOB_PREFETCH") ? 0 : 1;
    return s;
}

// where one K1 launch reads its jobs and leaves its results
struct ScoreIO {
    cudaStream_t st; const Job* jobs; DevBuf* results; DevBuf* hits; DevBuf* hit_count; uint32_t* job_counter; bool timed;
};
// run K1 over a job list; fills results (+hits)
int32_t run_score_io(pq_ctx* ctx, const ScoreIO& io, JobKind kind, uint32_t n_jobs, const int32_t* row_pred, const uint8_t* rows, const double* row_mass,
                     uint32_t hit_capacity, int check_window, const PeptideTable* ladders = nullptr) {
    CU(io.results->reserve(std::max<size_t>((size_t)n_jobs * sizeof(JobResult), 64)));
    CU(io.hits->reserve(std::max<size_t>((size_t)hit_capacity * sizeof(Hit), 64)));
    CU(io.hit_count->reserve(64));
    CU(cudaMemsetAsync(io.hit_count->p, 0, 4, io.st));
    ScoreLaunch L; memset(&L, 0, sizeof L);
    L.kind = kind; L.jobs = io.jobs; L.n_jobs = n_jobs;
    L.blobs = ctx->blobs.as<uint8_t>(); L.blob_off = ctx->blob_off.as<uint64_t>(); L.blob_len = ctx->blob_len.as<uint32_t>();
    L.rows = rows; L.row_mass = row_mass; L.mods = ctx->d_mods.as<ModTable>(); L.n_codes = (uint32_t)ctx->codes.n_codes;
    L.log10_fact = ctx->d_l10.as<double>(); L.ln_fact = ctx->d_lnf.as<double>();
    L.row_pred = row_pred;
    L.results = io.results->as<JobResult>(); L.hits = io.hits->as<Hit>(); L.hit_count = io.hit_count->as<uint32_t>();
    L.hit_capacity = hit_capacity; L.max_blob_bytes = ctx->max_blob; L.job_counter = io.job_counter;
    if (kind == kJobTargets) L.tgt = ctx->targeted_view();
    if (kind == kJobCompetitors && ladders && ladders->have_lad) { L.lad = ladders->lad.as<uint4>(); L.lad_stride = ladders->lad_stride; L.lad_len = ladders->lad_len.as<uint8_t>(); }
    if (kind == kJobShuffles && ctx->have_lists) { L.sh_forms = ctx->sh_forms.as<ShuffleForm>(); L.form_slot = ctx->sh_form_slot.as<uint32_t>(); L.lists = ctx->sh_lists.as<uint4>(); }
    ScoreConfig cfg = score_config(ctx, check_window);
    if (kind == kJobTargets && ctx->P.library_mode) cfg.no_bucket = 1;
    const int si = kind == kJobTargets ? 0 : kind == kJobCompetitors ? 1 : kind == kJobShuffles ? 2 : 3;
    if (io.timed) {
        Timer t(ctx, &ctx->cnt.ms_score, &ctx->cnt.ms_score_step[si]);
        ctx->cnt.score_kernel_launches += (uint64_t)launch_score(L, cfg, io.st);
    } else {
        ctx->cnt.score_kernel_launches += (uint64_t)launch_score(L, cfg, io.st);
    }
    CU(cudaGetLastError());
    return PQ_OK;
}
int32_t run_score(pq_ctx* ctx, JobKind kind, uint32_t n_jobs, const int32_t* row_pred, const uint8_t* rows, const double* row_mass,
                  uint32_t hit_capacity, int check_window, const PeptideTable* ladders = nullptr) {
    ScoreIO io{ctx->st, ctx->jobs.as<Job>(), &ctx->results, &ctx->hits, &ctx->hit_count, ctx->d_scalars.as<uint32_t>() + 12, true};
    return run_score_io(ctx, io, kind, n_jobs, row_pred, rows, row_mass, hit_capacity, check_window, ladders);
}
// Step 3 reads the fragment ladders of the reference rows from a table written once per reference build (k1_ladder.cu)
int32_t ensure_ladders(pq_ctx* ctx, PeptideTable& T, const uint8_t* rows, uint32_t n_rows, cudaStream_t st, bool timed) {
    if (T.have_lad || n_rows == 0 || getenv("PQ_STEP3_WALK")) return PQ_OK;
    std::string err; uint64_t launches = 0;
    bool ok;
    if (T.lad.reserve(16) != cudaSuccess) return fail(ctx, PQ_ERR_NOMEM, "ladder tables");
    const int max_len = std::max(ctx->P.max_len, 1);      // reference rows come from the digest: min_len .. max_len residues
    if (timed) { Timer t(ctx, &ctx->cnt.ms_ladder); ok = build_row_ladders(st, rows, n_rows, max_len, ctx->d_mods.as<ModTable>(), T.lad, T.lad_len, &T.lad_stride, &launches, err); }
    else ok = build_row_ladders(st, rows, n_rows, max_len, ctx->d_mods.as<ModTable>(), T.lad, T.lad_len, &T.lad_stride, &launches, err);
    if (!ok) return fail(ctx, PQ_ERR_CUDA, "ladder tables: " + err);
    ctx->cnt.other_kernel_launches += launches;
    T.have_lad = true;
    return PQ_OK;
}

template <class T> int32_t scan_exclusive(pq_ctx* ctx, const T* in, T* out, int n) {
    size_t tmp = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, ctx->st);
    CU(ctx->scan_tmp.reserve(tmp + 64));
    CU(cub::DeviceScan::ExclusiveSum(ctx->scan_tmp.p, tmp, in, out, n, ctx->st));
    return PQ_OK;
}

// pair / algorithmic-byte counters are accumulated on the device (k7_job_counters); fold them into ctx->cnt and zero them
int32_t pull_device_counters(pq_ctx* ctx) {
    flush_timers(ctx);
    if (!ctx->d_counters.p) return PQ_OK;
    DevCounters h;
    CU(cudaMemcpyAsync(&h, ctx->d_counters.p, sizeof h, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemsetAsync(ctx->d_counters.p, 0, sizeof h, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->cnt.pairs_step2 += h.pairs[0]; ctx->cnt.pairs_step3 += h.pairs[1]; ctx->cnt.pairs_step4 += h.pairs[2]; ctx->cnt.pairs_step5 += h.pairs[3];
    for (int i = 0; i < 4; ++i) { ctx->cnt.bytes_score_step[i] += h.bytes[i]; ctx->cnt.score_algorithmic_bytes += h.bytes[i]; }
    ctx->cnt.pairs_step4_bounded += h.bounded;
    return PQ_OK;
}

int32_t ensure_counters(pq_ctx* ctx) {
    if (ctx->d_counters.p) return PQ_OK;
    CU(ctx->d_counters.reserve(sizeof(DevCounters)));
    CU(cudaMemsetAsync(ctx->d_counters.p, 0, sizeof(DevCounters), ctx->st));
    CU(ctx->d_scalars.reserve(256));
    return PQ_OK;
}

template <class K, class V> int32_t sort_pairs(pq_ctx* ctx, const K* kin, K* kout, const V* vin, V* vout, int n, int end_bit) {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, n, 0, end_bit, ctx->st);
    CU(ctx->scan_tmp.reserve(tmp + 64));
    CU(cub::DeviceRadixSort::SortPairs(ctx->scan_tmp.p, tmp, kin, kout, vin, vout, n, 0, end_bit, ctx->st));
    return PQ_OK;
}

// Device -> caller buffer.  A pinned (page-locked) caller buffer takes the DMA directly; pageable memory goes through the
// context's pinned staging buffer in two alternating halves so the DMA of one chunk overlaps the memcpy of the previous one.
int32_t copy_to_caller(pq_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!bytes) return PQ_OK;
    cudaPointerAttributes at; memset(&at, 0, sizeof at);
    bool pinned = cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        return PQ_OK;
    }
    const size_t kChunk = 4u << 20;
    if (!ctx->stage_host) { CU(cudaHostAlloc(&ctx->stage_host, 2 * kChunk, cudaHostAllocDefault)); CU(cudaEventCreateWithFlags(&ctx->stage_ev[0], cudaEventDisableTiming)); CU(cudaEventCreateWithFlags(&ctx->stage_ev[1], cudaEventDisableTiming)); }
    size_t n_chunks = (bytes + kChunk - 1) / kChunk;
    for (size_t c = 0; c <= n_chunks; ++c) {
        if (c < n_chunks) {
            size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
            CU(cudaMemcpyAsync(static_cast<uint8_t*>(ctx->stage_host) + (c & 1) * kChunk, static_cast<const uint8_t*>(src) + off, len, cudaMemcpyDeviceToHost, ctx->st));
            CU(cudaEventRecord(ctx->stage_ev[c & 1], ctx->st));
        }
        if (c > 0) {
            size_t off = (c - 1) * kChunk, len = std::min(kChunk, bytes - off);
            CU(cudaEventSynchronize(ctx->stage_ev[(c - 1) & 1]));
            memcpy(static_cast<uint8_t*>(dst) + off, static_cast<uint8_t*>(ctx->stage_host) + ((c - 1) & 1) * kChunk, len);
        }
    }
    return PQ_OK;
}

// Host mirrors of the sorted spectrum metadata, only needed by the host-driven entry points (Step 5, pq_score_pairs)
int32_t ensure_host_spectra(pq_ctx* ctx) {
    if (ctx->host_spectra_valid) return PQ_OK;
    const int64_t n = ctx->nS;
    ctx->h_caller.resize(n); ctx->h_mass.resize(n); ctx->h_charge.resize(n); ctx->h_npeaks.resize(n); ctx->h_resident.resize(n); ctx->h_alias.resize(n);
    if (n) {
        std::vector<uint64_t> off((size_t)n + 1);
        CU(cudaMemcpyAsync(ctx->h_caller.data(), ctx->s_caller.p, n * 4, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(ctx->h_mass.data(), ctx->s_mass.p, n * 8, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(ctx->h_charge.data(), ctx->s_charge.p, n * 4, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(ctx->h_npeaks.data(), ctx->s_npeaks.p, n * 4, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(ctx->h_alias.data(), ctx->s_alias.p, n, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(off.data(), ctx->s_off.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        ctx->cnt.d2h_bytes += (uint64_t)n * 29 + 8;
        for (int64_t i = 0; i < n; ++i) ctx->h_resident[i] = (ctx->h_npeaks[i] == 0 || off[i + 1] > off[i]) ? 1 : 0;     // candidate-first: only some spectra have their peaks here
    }
    ctx->host_spectra_valid = true;
    return PQ_OK;
}

void reset_psm_state(pq_ctx* ctx) { for (bool& b : ctx->pending_step) b = false; ctx->pending_ref = false; ctx->n_psm = 0; ctx->n_seg = 0; ctx->nh3 = 0; ctx->comp3_ready = false; ctx->comp3_n = 0; ctx->have_shuffles = false; ctx->psms.clear(); ctx->comp5.clear(); }

}  // namespace

extern "C" {

int32_t pq_version(void) { return PQ_VERSION_MAJOR * 100 + PQ_VERSION_MINOR; }
const char* pq_last_error(const pq_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void pq_default_params(pq_params* p) {
    memset(p, 0, sizeof *p);
    p->tol = 10.0; p->tol_ppm = 1; p->score_method = 1; p->itol = 0.6; p->min_score = 12.0; p->min_peaks = 10;
    p->n_isotope = 1; p->isotope[0] = 0; p->max_var_mods = 3; p->max_mods_per_aa = 1; p->n_random = 1000;
    p->enzyme = 1; p->max_missed = 2; p->min_len = 7; p->max_len = 45; p->min_mass = 500.0; p->max_mass = 10000.0;
    p->shuffle_seed = 12457; p->mod_seed = 2022; p->prepare_at_load = 1; p->candidate_first = 1; p->search_at_load = 1;
    p->n_fixed = 1; p->fixed[0].id = 1; p->fixed[0].position = PQ_MOD_ANYWHERE; strcpy(p->fixed[0].residues, "C"); p->fixed[0].delta = 57.02146372057;
    p->n_var = 1; p->var[0].id = 2; p->var[0].position = PQ_MOD_ANYWHERE; strcpy(p->var[0].residues, "M"); p->var[0].delta = 15.99491461956;
    p->var[0].neutral_loss = 63.99828574784;
}

int32_t pq_create(const pq_params* params, pq_ctx** out) {
    pq_ctx* ctx = nullptr;
    if (!params || !out) return fail(nullptr, PQ_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) return fail(nullptr, PQ_ERR_NO_DEVICE, std::string("no CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    if (params->device < 0 || params->device >= ndev) return fail(nullptr, PQ_ERR_BAD_ARG, "device ordinal out of range");
    if (params->score_method != 1 && params->score_method != 2) return fail(nullptr, PQ_ERR_BAD_ARG, "score_method must be 1 (hyperscore) or 2 (MVH)");
    if (params->n_isotope < 1 || params->n_isotope > PQ_MAX_ISOTOPE) return fail(nullptr, PQ_ERR_BAD_ARG, "n_isotope out of range");
    { bool has0 = false; for (int i = 0; i < params->n_isotope; ++i) has0 |= params->isotope[i] == 0; if (!has0) return fail(nullptr, PQ_ERR_BAD_ARG, "isotope error list must contain 0 (CParameter.java:436-446)"); }
    if (!(params->itol > 0) || !(params->tol >= 0)) return fail(nullptr, PQ_ERR_BAD_ARG, "tolerances must be positive");
    { EnzymeRule er; if (!enzyme_rule(params->enzyme, er)) return fail(nullptr, PQ_ERR_UNSUPPORTED, "enzyme must be 0..17 (pg/DatabaseInput.java:145-267)"); }
    if (params->max_len > kMaxLen) return fail(nullptr, PQ_ERR_UNSUPPORTED, "max_len > 60");
    ctx = new pq_ctx();
    ctx->P = *params;
    memset(&ctx->cnt, 0, sizeof ctx->cnt);
    std::string err;
    if (!ctx->codes.init(*params, err)) { delete ctx; return fail(nullptr, PQ_ERR_UNSUPPORTED, err); }
    // Search at load: the spectrum preparation (st), the gather (st_copy) and the reference table (st_aux) gate every part, so their
    // CTAs are dispatched ahead of the waiting CTAs of the parts' own kernels (st_part / st_part2); the shuffle generation (st_shuf),
    // which only Step 4 of the first part waits for, fills what is left.
    int prio_lo = 0, prio_hi = 0;
    if (cudaSetDevice(params->device) == cudaSuccess && !getenv("PQ_NO_PRIORITY")) cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const int prio_mid = prio_hi < prio_lo ? std::min(prio_hi + 1, prio_lo - 1 > prio_hi ? prio_lo - 1 : prio_lo) : prio_lo;
    if (cudaSetDevice(params->device) != cudaSuccess || cudaStreamCreateWithPriority(&ctx->st, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->st_copy, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->st_aux, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->st_part, cudaStreamNonBlocking, prio_mid) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->st_part2, cudaStreamNonBlocking, prio_mid) != cudaSuccess ||
        cudaStreamCreateWithPriority(&ctx->st_shuf, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->shuf_ev, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx; return fail(nullptr, PQ_ERR_CUDA, "cannot initialise device");
    }
    cudaEventCreate(&ctx->ev0); cudaEventCreate(&ctx->ev1); cudaEventCreate(&ctx->uev0); cudaEventCreate(&ctx->uev1);
    int bins = (int)jround((double)(int)(2000.0 - 200.0) / (params->itol * 2.0));
    factorial_tables(ctx->log10_fact, ctx->ln_fact, std::max(bins + 512, 4096));
    ctx->log10_fact[65] = 0.0;
    for (int i = 1; i <= 64; ++i) ctx->log10_fact[65 + i] = std::log10(100.0 * i);
    bool ok = ctx->d_mods.reserve(sizeof(ModTable)) == cudaSuccess && ctx->d_l10.reserve(130 * 8) == cudaSuccess &&
              ctx->d_lnf.reserve(ctx->ln_fact.size() * 8) == cudaSuccess && ctx->d_shufmods.reserve(sizeof(ShuffleMods)) == cudaSuccess;
    if (ok) {
        cudaMemcpy(ctx->d_mods.p, &ctx->codes.table, sizeof(ModTable), cudaMemcpyHostToDevice);
        cudaMemcpy(ctx->d_l10.p, ctx->log10_fact, 130 * 8, cudaMemcpyHostToDevice);
        cudaMemcpy(ctx->d_lnf.p, ctx->ln_fact.data(), ctx->ln_fact.size() * 8, cudaMemcpyHostToDevice);
        ok = cudaMemcpy(ctx->d_shufmods.p, &ctx->codes.shuffle, sizeof(ShuffleMods), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (ok) {
        // text order of the pieces "id@site[delta]" of ModificationDB.getModificationString (:1695-1708): the last key of RankPSMCompare
        const int nm = ctx->codes.n_var + ctx->codes.n_fixed;
        std::vector<std::string> piece((size_t)std::max(nm, 1) * 64);
        std::vector<uint32_t> ord(piece.size());
        for (int m = 0; m < nm; ++m) for (int site = 0; site < 64; ++site) {
            char buf[96]; snprintf(buf, sizeof buf, "%d@%d[%.4f]", ctx->codes.mods[(size_t)m].id, site, ctx->codes.mods[(size_t)m].delta);
            piece[(size_t)m * 64 + site] = buf;
        }
        std::iota(ord.begin(), ord.end(), 0u);
        std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return piece[a] < piece[b]; });
        std::vector<uint16_t> rank(piece.size(), 0); uint16_t r = 0;
        for (size_t k = 0; k < ord.size(); ++k) { if (k && piece[ord[k]] != piece[ord[k - 1]]) ++r; rank[ord[k]] = r; }
        ok = ctx->t_piece_rank.reserve(rank.size() * 2) == cudaSuccess && cudaMemcpy(ctx->t_piece_rank.p, rank.data(), rank.size() * 2, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) { delete ctx; return fail(nullptr, PQ_ERR_CUDA, "device allocation failed"); }
    *out = ctx;
    return PQ_OK;
}

int32_t pq_destroy(pq_ctx* ctx) {
    if (!ctx) return PQ_OK;
    cudaSetDevice(ctx->P.device);
    cudaStreamSynchronize(ctx->st);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    flush_timers(ctx);
    if (ctx->mgf_host) cudaFreeHost(ctx->mgf_host);
    for (cudaEvent_t e : ctx->mgf.ev) cudaEventDestroy(e);
    if (ctx->stage_host) { cudaFreeHost(ctx->stage_host); cudaEventDestroy(ctx->stage_ev[0]); cudaEventDestroy(ctx->stage_ev[1]); }
    if (ctx->pack_host) cudaFreeHost(ctx->pack_host);
    for (cudaEvent_t e : ctx->pack_ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->free_events) cudaEventDestroy(e);
    if (ctx->uev0) cudaEventDestroy(ctx->uev0);
    if (ctx->uev1) cudaEventDestroy(ctx->uev1);
    for (cudaEvent_t e : ctx->chunk_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->k0_ev) cudaEventDestroy(e);
    if (ctx->meta_ev) cudaEventDestroy(ctx->meta_ev);
    if (ctx->targets_ev) cudaEventDestroy(ctx->targets_ev);
    cudaStream_t st = ctx->st, st_copy = ctx->st_copy, st_aux = ctx->st_aux;
    if (ctx->st_part) { cudaStreamSynchronize(ctx->st_part); cudaStreamDestroy(ctx->st_part); }
    if (ctx->st_part2) { cudaStreamSynchronize(ctx->st_part2); cudaStreamDestroy(ctx->st_part2); }
    if (ctx->st_shuf) { cudaStreamSynchronize(ctx->st_shuf); cudaStreamDestroy(ctx->st_shuf); }
    if (ctx->shuf_ev) cudaEventDestroy(ctx->shuf_ev);
    delete ctx;
    if (st) cudaStreamDestroy(st);
    if (st_copy) cudaStreamDestroy(st_copy);
    if (st_aux) cudaStreamDestroy(st_aux);
    return PQ_OK;
}

int32_t pq_event_record(pq_ctx* ctx, int32_t which) {
    if (!ctx) return PQ_ERR_BAD_ARG;
    cudaSetDevice(ctx->P.device);
    CU(cudaEventRecord(which ? ctx->uev1 : ctx->uev0, ctx->st));
    return PQ_OK;
}
int32_t pq_event_elapsed_ms(pq_ctx* ctx, double* ms) {
    if (!ctx || !ms) return PQ_ERR_BAD_ARG;
    cudaSetDevice(ctx->P.device);
    CU(cudaEventSynchronize(ctx->uev1));
    float f = 0; CU(cudaEventElapsedTime(&f, ctx->uev0, ctx->uev1));
    *ms = f;
    return PQ_OK;
}
namespace pq { namespace {
__global__ void k_selftest_div3(uint64_t n, uint64_t seed, unsigned long long* bad) {
    unsigned long long local = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (i + 1);            // splitmix64
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        const uint64_t frac = z & 0xFFFFFFFFFFFFFull, ex = 1023 - 8 + (z >> 52) % 24;
        const double x = __longlong_as_double((long long)((ex << 52) | frac));
        for (int d = -2; d <= 2; ++d) {                                  // the value and its neighbours
            const double v = __longlong_as_double(__double_as_longlong(x) + d);
            if (__double_as_longlong(div3_rn(v)) != __double_as_longlong(__ddiv_rn(v, 3.0))) ++local;
        }
        const double y = __dmul_rn(3.0, x);                              // multiples of three (exact quotients) and what lies beside them
        for (int d = -1; d <= 1; ++d) {
            const double v = __longlong_as_double(__double_as_longlong(y) + d);
            if (__double_as_longlong(div3_rn(v)) != __double_as_longlong(__ddiv_rn(v, 3.0))) ++local;
        }
    }
    if (local) atomicAdd(bad, local);
}
} }
int32_t pq_selftest(pq_ctx* ctx, int32_t which, uint64_t n, uint64_t seed, uint64_t* mismatches) {
    if (!ctx || !mismatches || which != 1) return PQ_ERR_BAD_ARG;
    cudaSetDevice(ctx->P.device);
    CU(ctx->scratch.reserve(64));
    CU(cudaMemsetAsync(ctx->scratch.p, 0, 8, ctx->st));
    if (n) pq::k_selftest_div3<<<148 * 8, 256, 0, ctx->st>>>(n, seed, ctx->scratch.as<unsigned long long>());
    unsigned long long bad = 0;
    CU(cudaMemcpyAsync(&bad, ctx->scratch.p, 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    *mismatches = bad;
    return PQ_OK;
}
int32_t pq_get_counters(pq_ctx* ctx, pq_counters* out) {
    if (!ctx || !out) return PQ_ERR_BAD_ARG;
    cudaSetDevice(ctx->P.device);
    int32_t rc = pull_device_counters(ctx);
    if (rc) return rc;
    ensure_forms(ctx);                       // cnt.target_forms
    *out = ctx->cnt;
    return PQ_OK;
}
int32_t pq_reset_counters(pq_ctx* ctx) {
    if (!ctx) return PQ_ERR_BAD_ARG;
    cudaSetDevice(ctx->P.device);
    int32_t rc = pull_device_counters(ctx);
    if (rc) return rc;
    pq_counters k = ctx->cnt; memset(&ctx->cnt, 0, sizeof ctx->cnt);
    ctx->cnt.spectra_loaded = k.spectra_loaded; ctx->cnt.spectra_usable = k.spectra_usable; ctx->cnt.target_forms = k.target_forms;
    ctx->cnt.reference_sequences = k.reference_sequences; ctx->cnt.reference_forms = k.reference_forms;
    return PQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// pq_load_spectra.  Two ways for the peaks to reach HBM:
//   candidate-first (targets known, all five caller arrays in pinned, device-mapped memory, pq_params.candidate_first): the 28
//     bytes per spectrum of precursor data go up first, the spectra are ordered by mass and the precursor windows of the target
//     table are enumerated (K3); only spectra with a target in their window are ever read again — by Steps 2-5 here, by
//     SpectraInput.spectraMap in the reference (pg/SpectraInput.java:243-254) — so only their peaks cross PCIe: a gather kernel
//     reads them straight from the caller's arrays into the mass-sorted CSR, chunk by chunk, K0 following each chunk.
//   full upload (pageable buffers, or no targets yet): every peak crosses in 64 MB chunks on a copy stream, the compute stream
//     follows chunk by chunk (copy into the sorted CSR, K0 with prepare_at_load).
static int32_t load_spectra_impl(pq_ctx* ctx, int64_t n, const double* prec_mz, const int32_t* charge, const uint64_t* off,
                                 const double* mz, const double* inten, bool peaks_on_device);
int32_t pq_load_spectra(pq_ctx* ctx, int64_t n, const double* prec_mz, const int32_t* charge, const uint64_t* off,
                        const double* mz, const double* inten) {
    if (!ctx || n < 0 || (n > 0 && (!prec_mz || !charge || !off || !mz || !inten))) return fail(ctx, PQ_ERR_BAD_ARG, "null argument");
    return load_spectra_impl(ctx, n, prec_mz, charge, off, mz, inten, false);
}

// MGF text -> spectra on the device -> the loader above (replaces the compomics MgfFileIterator loop of
// SpectraInput.readMSMSfast, pg/SpectraInput.java:150-195).  The text is parsed on the GPU (pq_mgf.cu); the peaks never exist on the host.
int32_t pq_load_mgf(pq_ctx* ctx, const char* text, uint64_t n_bytes, int64_t* n_spectra) {
    if (!ctx || (!text && n_bytes) || !n_spectra) return fail(ctx, PQ_ERR_BAD_ARG, "null argument");
    cudaSetDevice(ctx->P.device);
    Marks mk("load_mgf");
    *n_spectra = 0;
    int64_t ns = 0; uint64_t np = 0; uint64_t launches = 0; std::string err;
    if (!parse_mgf_on_device(ctx->st, ctx->st_copy, text, n_bytes, ctx->mgf, ctx->ld[0], ctx->ld[1], ctx->ld[2], ctx->ld[3], ctx->ld[4], &ns, &np, &launches, err)) {
        cudaStreamSynchronize(ctx->st_copy); cudaStreamSynchronize(ctx->st); cudaGetLastError();
        return fail(ctx, err.find("malformed") != std::string::npos ? PQ_ERR_BAD_ARG : PQ_ERR_CUDA, "device MGF parser: " + err);
    }
    ctx->cnt.other_kernel_launches += launches; ctx->cnt.h2d_bytes += n_bytes;
    mk.mark("parsed on the device");
    // the 28 bytes per spectrum of precursor data come back to the host: the loader validates and cuts its chunks there
    const size_t need = (size_t)ns * 20 + 64;
    if (ctx->mgf_host_cap < need) {
        if (ctx->mgf_host) cudaFreeHost(ctx->mgf_host);
        ctx->mgf_host = nullptr; ctx->mgf_host_cap = 0;
        CU(cudaHostAlloc(&ctx->mgf_host, need + need / 8, cudaHostAllocDefault));
        ctx->mgf_host_cap = need + need / 8;
    }
    uint64_t* h_off = static_cast<uint64_t*>(ctx->mgf_host); double* h_prec = reinterpret_cast<double*>(h_off + ns + 1); int32_t* h_charge = reinterpret_cast<int32_t*>(h_prec + ns);
    if (ns) {
        CU(cudaMemcpyAsync(h_off, ctx->ld[2].p, ((size_t)ns + 1) * 8, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(h_prec, ctx->ld[0].p, (size_t)ns * 8, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(h_charge, ctx->ld[1].p, (size_t)ns * 4, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        ctx->cnt.d2h_bytes += (uint64_t)ns * 20 + 8;
    } else h_off[0] = 0;
    int32_t rc = load_spectra_impl(ctx, ns, h_prec, h_charge, h_off, ctx->ld[3].as<double>(), ctx->ld[4].as<double>(), true);
    if (rc) return rc;
    *n_spectra = ns;
    return PQ_OK;
}

// The device parser alone, results back on the host: pq_parse_mgf's contract (two calls: sizes, then fill), ~100x its speed.
int32_t pq_parse_mgf_device(pq_ctx* ctx, const char* text, uint64_t n_bytes, int64_t* n_spectra, uint64_t* n_peaks, double* precursor_mz, int32_t* charge,
                            uint64_t* peak_offset, double* mz, double* intensity) {
    if (!ctx || (!text && n_bytes) || !n_spectra || !n_peaks) return fail(ctx, PQ_ERR_BAD_ARG, "null argument");
    cudaSetDevice(ctx->P.device);
    const bool fill = precursor_mz && charge && peak_offset && mz && intensity;
    if (!(ctx->mgf_cache_text == text && ctx->mgf_cache_bytes == n_bytes && fill)) {        // the fill call re-uses what the size call parsed
        ctx->mgf_cache_text = nullptr;
        uint64_t launches = 0; std::string err;
        ctx->nS = 0; ctx->nCaller = 0; ctx->step_done = 0; reset_psm_state(ctx); ctx->n_prepared = 0; ctx->prepared_at_load = false; ctx->residency = 0;      // the loader's staging buffers are about to be overwritten
        if (!parse_mgf_on_device(ctx->st, ctx->st_copy, text, n_bytes, ctx->mgf, ctx->ld[0], ctx->ld[1], ctx->ld[2], ctx->ld[3], ctx->ld[4], &ctx->mgf_cache_ns, &ctx->mgf_cache_np,
                                 &launches, err)) {
            cudaStreamSynchronize(ctx->st_copy); cudaStreamSynchronize(ctx->st); cudaGetLastError();
            return fail(ctx, err.find("malformed") != std::string::npos ? PQ_ERR_BAD_ARG : PQ_ERR_CUDA, "device MGF parser: " + err);
        }
        ctx->cnt.other_kernel_launches += launches; ctx->cnt.h2d_bytes += n_bytes;
        ctx->mgf_cache_text = text; ctx->mgf_cache_bytes = n_bytes;
    }
    const int64_t ns = ctx->mgf_cache_ns; const uint64_t np = ctx->mgf_cache_np;
    if (fill) {
        if (*n_spectra < ns || *n_peaks < np) return fail(ctx, PQ_ERR_CAPACITY, "spectrum buffers too small");
        int32_t rc = PQ_OK;
        if (ns) {
            rc = copy_to_caller(ctx, precursor_mz, ctx->ld[0].p, (size_t)ns * 8); if (rc) return rc;
            rc = copy_to_caller(ctx, charge, ctx->ld[1].p, (size_t)ns * 4); if (rc) return rc;
            rc = copy_to_caller(ctx, peak_offset, ctx->ld[2].p, ((size_t)ns + 1) * 8); if (rc) return rc;
        } else peak_offset[0] = 0;
        if (np) { rc = copy_to_caller(ctx, mz, ctx->ld[3].p, (size_t)np * 8); if (rc) return rc; rc = copy_to_caller(ctx, intensity, ctx->ld[4].p, (size_t)np * 8); if (rc) return rc; }
        ctx->cnt.d2h_bytes += (uint64_t)ns * 20 + 8 + np * 16;
        ctx->mgf_cache_text = nullptr;
    }
    *n_spectra = ns; *n_peaks = np;
    return PQ_OK;
}

static int32_t load_spectra_impl(pq_ctx* ctx, int64_t n, const double* prec_mz, const int32_t* charge, const uint64_t* off,
                                 const double* mz, const double* inten, bool peaks_on_device) {
    cudaSetDevice(ctx->P.device);
    ctx->mgf_cache_text = nullptr;
    Marks mk("load_spectra");
    cudaStream_t st = ctx->st, sc = ctx->st_copy;
    // The spectrum-independent work of a search that will run inside this call (search_at_load) — the digest of the staged protein
    // database and the shuffles of every target form — is queued by a helper thread (both have host round trips for their sizes)
    // while this thread sets up the upload: by the time the first spectra have arrived and are prepared, it is done.
    struct EarlyWork { std::thread th; int32_t rc = PQ_OK; std::string err; uint64_t launches = 0, d2h = 0; bool digest = false, shuffles = false; } early;
    // cf_staged: a second thread packs, uploads and scatters the needed spectra chunk by chunk and queues each chunk's K0 behind it
    struct Stager { std::thread th; std::string err; uint64_t launches = 0; std::atomic<size_t> issued{0}; std::atomic<bool> failed{false}; } stager;
    // whatever goes wrong from here on, the context ends up without spectra and with nothing in flight
    auto fail_load = [&](int32_t code, const std::string& msg) -> int32_t {
        if (early.th.joinable()) early.th.join();
        if (stager.th.joinable()) stager.th.join();
        cudaStreamSynchronize(ctx->st_part); cudaStreamSynchronize(ctx->st_part2); cudaStreamSynchronize(ctx->st_shuf);
        cudaStreamSynchronize(sc); cudaStreamSynchronize(st); cudaStreamSynchronize(ctx->st_aux); cudaGetLastError();
        ctx->nS = 0; ctx->nCaller = 0; ctx->nZero = 0; ctx->max_peaks = 0; ctx->step_done = 0; reset_psm_state(ctx); ctx->n_prepared = 0; ctx->prepared_at_load = false;
        ctx->residency = 0; ctx->n_resident = 0; ctx->cnt.spectra_loaded = 0; ctx->cnt.spectra_usable = 0; ctx->host_spectra_valid = false;
        return fail(ctx, code, msg);
    };
#define CL(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail_load(PQ_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
    // spectra without a charge (0) are searched as z = 2 and, failing that, z = 3: each is two virtual spectra
    std::vector<int32_t> zeros;
    for (int64_t i = 0; i < n; ++i) { if (charge[i] < 0) return fail_load(PQ_ERR_BAD_ARG, "charge must be >= 0 (0 = no charge given)"); if (charge[i] == 0) zeros.push_back((int32_t)i); }
    if (ctx->P.library_mode && !zeros.empty()) return fail_load(PQ_ERR_BAD_ARG, "library_mode: a library spectrum needs a charge (MsLibrarySearchWorker.java:234-236 ignores the others)");
    const int64_t n0 = (int64_t)zeros.size(), nv = n + n0;
    if (nv > 0x7fffffffLL / PQ_MAX_ISOTOPE) return fail_load(PQ_ERR_UNSUPPORTED, "too many spectra for one context");
    ctx->nS = 0; ctx->step_done = 0; reset_psm_state(ctx); ctx->n_prepared = 0; ctx->prepared_at_load = false; ctx->residency = 0; ctx->host_spectra_valid = false;
    ctx->cnt.spectra_loaded = 0; ctx->cnt.spectra_usable = 0;
    if (n == 0) { ctx->nCaller = 0; ctx->nZero = 0; ctx->max_peaks = 0; return PQ_OK; }
    const uint64_t total = off[n];
    DevBuf& in_prec = ctx->ld[0]; DevBuf& in_charge = ctx->ld[1]; DevBuf& in_off = ctx->ld[2]; DevBuf& in_mz = ctx->ld[3]; DevBuf& in_in = ctx->ld[4];
    DevBuf& in_mass = ctx->ld[5]; DevBuf& key = ctx->ld[6]; DevBuf& key2 = ctx->ld[7]; DevBuf& idx = ctx->ld[8]; DevBuf& in_vz = ctx->ld[9];
    CL(in_prec.reserve(n * 8)); CL(in_charge.reserve(n * 4)); CL(in_off.reserve((n + 1) * 8)); CL(ctx->d_zero.reserve((size_t)n0 * 4 + 64));
    CL(in_mass.reserve(nv * 8)); CL(in_vz.reserve(nv * 4)); CL(key.reserve(nv * 8)); CL(key2.reserve(nv * 8)); CL(idx.reserve(nv * 4));
    CL(ctx->s_prec.reserve(nv * 8)); CL(ctx->s_charge.reserve(nv * 4)); CL(ctx->s_mass.reserve(nv * 8)); CL(ctx->s_off.reserve((nv + 1) * 8));
    CL(ctx->s_caller.reserve(nv * 4)); CL(ctx->s_vid.reserve(nv * 4)); CL(ctx->s_npeaks.reserve(nv * 4)); CL(ctx->s_alias.reserve(nv + 64)); CL(ctx->s_inv.reserve(nv * 4));
    CL(ctx->scratch.reserve((nv + 2) * 8));
    // pinned + device-mapped caller arrays?
    auto dev_ptr = [](const void* p) -> const void* {
        cudaPointerAttributes a; if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (a.type != cudaMemoryTypeHost) return nullptr;
        void* d = nullptr; if (cudaHostGetDevicePointer(&d, const_cast<void*>(p), 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return d;
    };
    const void* d_h_mz = peaks_on_device ? mz : dev_ptr(mz); const void* d_h_in = peaks_on_device ? inten : dev_ptr(inten);      // (the MGF parser left the peaks in ld[3] / ld[4])
    const bool pinned_all = d_h_mz && d_h_in;                     // only the peak arrays are read in place; the precursor arrays are copied
    const bool want_cf = ctx->P.candidate_first != 0 && getenv("PQ_LOAD_FULL") == nullptr;
    // candidate-first needs the targets; the peaks are then gathered by a kernel from page-locked caller arrays, or — pageable arrays —
    // packed by host threads into a page-locked ring and scattered on the device (cf_staged)
    const bool cf_staged = want_cf && ctx->have_targets && !ctx->targets.seqs.empty() && !pinned_all && !peaks_on_device && getenv("PQ_PAGEABLE_FULL") == nullptr;
    const bool cf = want_cf && ctx->have_targets && !ctx->targets.seqs.empty() && (pinned_all || cf_staged);
    const bool digest_here = ctx->ref_staged && !ctx->ref_digested && ctx->have_targets;
    if (digest_here && !cf) {   // full upload: the target keys of a staged reference digest go up before the copy engine fills with spectra
        std::string err;        // (candidate-first moves the peaks with a kernel: the keys go up on the digest's own stream, later)
        if (!upload_target_keys_for_reference(st, ctx->targets.seqs, ctx->reference.dev, err)) return fail_load(PQ_ERR_CUDA, "reference digest: " + err);
    }
    int k_launch = 0;
    const bool early_ok = cf && ctx->P.search_at_load != 0 && !getenv("PQ_NO_SEARCH_AT_LOAD") && !getenv("PQ_NO_EARLY_WORK") && ctx->ref_staged && ctx->n_targeted == 0 && n0 == 0 &&
                          ctx->P.n_random <= 20000 && ctx->P.prepare_at_load != 0 && getenv("PQ_PREPARE_AT_STEP2") == nullptr;
    if (early_ok) {
        int32_t rc = finalize_targets(ctx);                          // the shuffle generator reads the target table's side arrays
        if (rc) return fail_load(rc, ctx->err);
        if (!ctx->targets_ev) CL(cudaEventCreateWithFlags(&ctx->targets_ev, cudaEventDisableTiming));
        CL(cudaEventRecord(ctx->targets_ev, st));
        const bool need_digest = !ctx->ref_digested;
        const bool need_shuffles = !(ctx->shuffles_all && ctx->shuffles_epoch == ctx->targets_epoch && ctx->have_shuffles);
        if (need_digest || need_shuffles) early.th = std::thread([ctx, &early, need_digest, need_shuffles]() {
            cudaSetDevice(ctx->P.device);
            if (need_digest) {
                std::string err; uint64_t l = 0;
                if (!upload_target_keys_for_reference(ctx->st_aux, ctx->targets.seqs, ctx->reference.dev, err) ||
                    !digest_reference_on_device(ctx->st_aux, ctx->P, ctx->codes, ctx->targets.seqs, ctx->reference.dev, &l, err)) { early.rc = PQ_ERR_CUDA; early.err = "reference digest: " + err; return; }
                early.launches += l; early.digest = true;
            }
            if (need_shuffles) {
                if (cudaStreamWaitEvent(ctx->st_shuf, ctx->targets_ev, 0) != cudaSuccess) { early.rc = PQ_ERR_CUDA; early.err = "early shuffles: stream wait"; return; }
                pq_counters c; memset(&c, 0, sizeof c);
                ShuffleTotals tot; memset(&tot, 0, sizeof tot);
                const int32_t rc = generate_shuffles(ctx, ctx->st_shuf, true, tot, false, &c);      // a stream of its own: only Step 4 of a part waits for it (shuf_ev), Steps 2 and 3 do not
                early.launches += c.other_kernel_launches; early.d2h += c.d2h_bytes;
                if (rc) { early.rc = rc; return; }
                if (cudaEventRecord(ctx->shuf_ev, ctx->st_shuf) != cudaSuccess) { early.rc = PQ_ERR_CUDA; early.err = "early shuffles: event"; return; }
                early.shuffles = true;
            }
        });
    }
    // ---- precursor arrays, neutral masses, the mass order, the sorted precursor arrays ----
    CL(cudaMemcpyAsync(in_prec.p, prec_mz, n * 8, cudaMemcpyHostToDevice, st));
    CL(cudaMemcpyAsync(in_charge.p, charge, n * 4, cudaMemcpyHostToDevice, st));
    CL(cudaMemcpyAsync(in_off.p, off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n0) CL(cudaMemcpyAsync(ctx->d_zero.p, zeros.data(), (size_t)n0 * 4, cudaMemcpyHostToDevice, st));
    {
        Timer t(ctx, &ctx->cnt.ms_prepare);
        k_launch += launch_virtual_meta(nv, n, in_prec.as<double>(), in_charge.as<int32_t>(), ctx->d_zero.as<int32_t>(), in_mass.as<double>(), in_vz.as<int32_t>(),
                                        key.as<uint64_t>(), idx.as<int32_t>(), st);
        size_t tmp = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp, key.as<uint64_t>(), key2.as<uint64_t>(), idx.as<int32_t>(), ctx->s_vid.as<int32_t>(), (int)nv, 0, 64, st);
        CL(ctx->scan_tmp.reserve(tmp + 64));
        CL(cub::DeviceRadixSort::SortPairs(ctx->scan_tmp.p, tmp, key.as<uint64_t>(), key2.as<uint64_t>(), idx.as<int32_t>(), ctx->s_vid.as<int32_t>(), (int)nv, 0, 64, st));
        k_launch += launch_gather_meta(nv, n, ctx->s_vid.as<int32_t>(), in_prec.as<double>(), in_charge.as<int32_t>(), ctx->d_zero.as<int32_t>(), in_vz.as<int32_t>(),
                                       in_mass.as<double>(), in_off.as<uint64_t>(), ctx->s_prec.as<double>(), ctx->s_charge.as<int32_t>(), ctx->s_mass.as<double>(),
                                       ctx->s_caller.as<int32_t>(), ctx->s_npeaks.as<uint32_t>(), ctx->s_alias.as<uint8_t>(), st);
        k_launch += launch_invert_order(nv, ctx->s_vid.as<int32_t>(), ctx->s_inv.as<int32_t>(), st);
        k_launch += 4;
    }
    mk.mark("precursor arrays + sort queued");
    // ---- the offsets are validated on the host while that runs; the full upload cuts its chunks in the same walk ----
    uint64_t chunk_peaks = 4u << 20;                                 // 64 MB of m/z + intensity per chunk
    if (const char* e = getenv("PQ_LOAD_CHUNK_PEAKS")) { long long v = atoll(e); if (v > 0) chunk_peaks = (uint64_t)v; }   // tests: many small chunks
    std::vector<int64_t> cut; cut.push_back(0);
    uint32_t maxp = 0; uint64_t usable = 0, alias_peaks = 0;
    {
        uint64_t next = off[0] + chunk_peaks;
        for (int64_t i = 0; i < n; ++i) {
            if (off[i + 1] < off[i] || off[i + 1] > total) return fail_load(PQ_ERR_BAD_ARG, "peak offsets must be non-decreasing");
            const uint64_t c = off[i + 1] - off[i];
            if (c > 4096) return fail_load(PQ_ERR_UNSUPPORTED, "more than 4096 peaks in one spectrum");
            if (c > maxp) maxp = (uint32_t)c;
            if ((int64_t)c > (int64_t)min_peaks_eff(ctx)) ++usable;
            if (charge[i] == 0) alias_peaks += c;
            if (i + 1 < n && off[i + 1] >= next) { cut.push_back(i + 1); next = off[i + 1] + chunk_peaks; }
        }
        cut.push_back(n);
    }
    const size_t n_chunks = cut.size() - 1;
    ctx->nS = nv; ctx->nCaller = n; ctx->nZero = n0; ctx->max_peaks = maxp;
    ctx->cnt.spectra_loaded = (uint64_t)n; ctx->cnt.spectra_usable = usable;
    mk.mark("offset walk");
    PrepConfig pc; pc.itol = ctx->P.itol; pc.method = ctx->P.score_method;
    pc.mvh_bins = (int32_t)jround((double)(int)(2000.0 - 200.0) / (ctx->P.itol * 2.0));
    bool eager = ctx->P.prepare_at_load != 0 && getenv("PQ_PREPARE_AT_STEP2") == nullptr;
    bool stream_search = false, searched = false; size_t n_chunks_cf = 0; uint32_t chunk_spectra_cf = 0;
    SpectraView sp = ctx->view();
    if (cf) {
        // ---- candidate-first: target table -> precursor windows -> the spectra worth bringing over ----
        int32_t rc = finalize_targets(ctx);                          // device build of the target table (k6), queued on st
        if (rc) return fail_load(rc, ctx->err);
        rc = ensure_counters(ctx); if (rc) return fail_load(rc, ctx->err);
        PeptideTable& T = ctx->targets;
        CL(ctx->s_res.reserve((size_t)(nv + 1) * 4)); CL(ctx->prep_id.reserve((size_t)(nv + 1) * 4)); CL(ctx->prep_list.reserve((size_t)(nv + 2) * 8));
        CL(ctx->njobs.reserve(64)); CL(ctx->jobs.reserve((size_t)nv * ctx->P.n_isotope * sizeof(Job)));
        CL(cudaMemsetAsync(ctx->njobs.p, 0, 16, st)); CL(cudaMemsetAsync(ctx->s_res.p, 0, (size_t)(nv + 1) * 4, st));
        WindowConfig W = window_config(ctx, true);
        struct { uint32_t n_res; uint32_t pad; uint64_t res_peaks; uint64_t blob_bytes; } hn = {0, 0, 0, 0};
        {
            Timer t(ctx, &ctx->cnt.ms_window);
            if (T.n()) k_launch += launch_enumerate_targets(sp, T.mass.as<double>(), (uint32_t)T.n(), W, ctx->jobs.as<Job>(), ctx->njobs.as<uint32_t>(), ctx->s_res.as<int32_t>(), st);
            rc = scan_exclusive(ctx, ctx->s_res.as<int32_t>(), ctx->prep_id.as<int32_t>(), (int)nv + 1);                 // rank among the resident spectra = blob id
            if (rc) return fail_load(rc, ctx->err);
            size_t tmp = 0;
            cub::CountingInputIterator<int32_t> iota(0);
            uint32_t* d_nsel = ctx->njobs.as<uint32_t>() + 1;
            cub::DeviceSelect::Flagged(nullptr, tmp, iota, ctx->s_res.as<int32_t>(), ctx->prep_list.as<int32_t>(), d_nsel, (int)nv, st);
            CL(ctx->scan_tmp.reserve(tmp + 64));
            CL(cub::DeviceSelect::Flagged(ctx->scan_tmp.p, tmp, iota, ctx->s_res.as<int32_t>(), ctx->prep_list.as<int32_t>(), d_nsel, (int)nv, st));
            k_launch += launch_resident_counts(nv, ctx->s_npeaks.as<uint32_t>(), ctx->s_res.as<int32_t>(), ctx->scratch.as<uint64_t>(), st);
            rc = scan_exclusive(ctx, ctx->scratch.as<uint64_t>(), ctx->s_off.as<uint64_t>(), (int)nv + 1);
            if (rc) return fail_load(rc, ctx->err);
            // the same spectra in CALLER order (ascending host address): the order the gather reads them in
            CL(ctx->z_keep.reserve((size_t)nv + 64)); CL(ctx->gather_list.reserve((size_t)(nv + 2) * 4));
            k_launch += launch_flags_by_virtual(nv, ctx->s_inv.as<int32_t>(), ctx->s_res.as<int32_t>(), ctx->z_keep.as<uint8_t>(), st);
            cub::DeviceSelect::Flagged(nullptr, tmp, ctx->s_inv.as<int32_t>(), ctx->z_keep.as<uint8_t>(), ctx->gather_list.as<int32_t>(), d_nsel, (int)nv, st);
            CL(ctx->scan_tmp.reserve(tmp + 64));
            CL(cub::DeviceSelect::Flagged(ctx->scan_tmp.p, tmp, ctx->s_inv.as<int32_t>(), ctx->z_keep.as<uint8_t>(), ctx->gather_list.as<int32_t>(), d_nsel, (int)nv, st));
            k_launch += 6;
        }
        struct { uint32_t njobs, nsel; unsigned long long cand; } hj = {0, 0, 0};
        CL(cudaMemcpyAsync(&hn.n_res, ctx->prep_id.as<int32_t>() + nv, 4, cudaMemcpyDeviceToHost, st));
        CL(cudaMemcpyAsync(&hn.res_peaks, ctx->s_off.as<uint64_t>() + nv, 8, cudaMemcpyDeviceToHost, st));
        CL(cudaMemcpyAsync(&hj, ctx->njobs.p, 16, cudaMemcpyDeviceToHost, st));
        CL(cudaStreamSynchronize(st));
        ctx->cnt.d2h_bytes += 28;
        mk.mark("target table + windows (waits for the precursor sort)");
        const uint32_t n_res = hn.n_res;
        CL(ctx->s_mz.reserve(std::max<uint64_t>(hn.res_peaks, 1) * 8)); CL(ctx->s_in.reserve(std::max<uint64_t>(hn.res_peaks, 1) * 8));
        sp = ctx->view();
        ctx->residency = 1; ctx->n_resident = n_res; ctx->loaded_epoch = ctx->targets_epoch;
        ctx->cnt.h2d_bytes += (uint64_t)n * 28 + 8 + (uint64_t)n0 * 4 + (peaks_on_device ? 0 : hn.res_peaks * 16);
        if (n_res) {
            // blob offsets of the resident spectra (blob id = rank in prep_list)
            CL(ctx->blob_off.reserve((size_t)(n_res + 2) * 8)); CL(ctx->blob_len.reserve((size_t)(n_res + 2) * 4));
            uint64_t* sizes = ctx->scratch.as<uint64_t>();
            k_launch += launch_blob_sizes(sp, ctx->prep_list.as<int32_t>(), (int32_t)n_res, sizes, st);
            rc = scan_exclusive(ctx, sizes, ctx->blob_off.as<uint64_t>(), (int)n_res + 1);
            if (rc) return fail_load(rc, ctx->err);
            CL(cudaMemcpyAsync(&hn.blob_bytes, ctx->blob_off.as<uint64_t>() + n_res, 8, cudaMemcpyDeviceToHost, st));
            CL(cudaStreamSynchronize(st));
            if (eager && ctx->blobs.cap < hn.blob_bytes + 256) {
                size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b);
                if ((double)(hn.blob_bytes + 256 - ctx->blobs.cap) > 0.6 * (double)free_b) eager = false;
            }
            if (eager && ctx->blobs.reserve(hn.blob_bytes + 256) != cudaSuccess) { cudaGetLastError(); eager = false; }
            // the peaks of the resident spectra, chunk by chunk (mass order): gather from the caller's arrays on the copy stream, K0 behind it
            int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->P.device);
            uint32_t chunk_spectra = 32768;
            if (const char* e = getenv("PQ_LOAD_CHUNK_SPECTRA")) { long long v = atoll(e); if (v > 0) chunk_spectra = (uint32_t)v; }
            const size_t nck = ((size_t)n_res + chunk_spectra - 1) / chunk_spectra;
            while (ctx->chunk_ev.size() < nck + 1) { cudaEvent_t e; CL(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->chunk_ev.push_back(e); }
            if (!ctx->meta_ev) CL(cudaEventCreateWithFlags(&ctx->meta_ev, cudaEventDisableTiming));
            CL(ctx->k0_counters.reserve((nck + 1) * 16)); CL(cudaMemsetAsync(ctx->k0_counters.p, 0, (nck + 1) * 16, st));
            CL(ctx->chunk_done.reserve((nck + 1) * 4)); CL(cudaMemsetAsync(ctx->chunk_done.p, 0, (nck + 1) * 4, st));
            CL(cudaEventRecord(ctx->meta_ev, st));
            CL(cudaStreamWaitEvent(sc, ctx->meta_ev, 0));
            // search while loading?  everything a search needs must be known now, and K0 must run here
            stream_search = eager && ctx->n_targeted == 0 && ctx->P.search_at_load != 0 && !getenv("PQ_NO_SEARCH_AT_LOAD") && ctx->ref_staged && n0 == 0 && ctx->P.n_random <= 20000;
            if (stream_search) while (ctx->k0_ev.size() < nck + 1) { cudaEvent_t e; CL(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->k0_ev.push_back(e); }
            n_chunks_cf = nck; chunk_spectra_cf = chunk_spectra;
            Timer t(ctx, &ctx->cnt.ms_prepare);
            const bool trace_env = getenv("PQ_LOAD_TRACE") != nullptr;      // diagnostic: when did every chunk's gather / K0 finish?
            const bool trace = trace_env;
            std::vector<cudaEvent_t> tr_g, tr_k; cudaEvent_t tr0 = nullptr;
            if (trace) { cudaEventCreate(&tr0); cudaEventRecord(tr0, sc); }
            // one gather launch for the whole upload (its CTAs stay resident, chunk_done[c] counts the CTAs that have finished chunk c);
            // PQ_GATHER_PER_CHUNK: one launch per chunk, ordered by events (the earlier form, for A/B runs)
            const bool one_gather = cf_staged || (getenv("PQ_GATHER_PER_CHUNK") == nullptr && !trace_env);
            uint32_t gather_ctas = 0;
            if (cf_staged) {
                // what to pack: caller spectrum and peak count of every gather-list entry, and where each entry starts in the packed stream
                CL(ctx->pack_src.reserve((size_t)n_res * 4 + 64)); CL(ctx->pack_cnt.reserve(((size_t)n_res + 2) * 8)); CL(ctx->pack_off.reserve(((size_t)n_res + 2) * 8));
                k_launch += launch_gather_sources(ctx->gather_list.as<int32_t>(), (int32_t)n_res, ctx->s_caller.as<int32_t>(), ctx->s_npeaks.as<uint32_t>(),
                                                  ctx->pack_src.as<int32_t>(), ctx->pack_cnt.as<uint64_t>(), st);
                rc = scan_exclusive(ctx, ctx->pack_cnt.as<uint64_t>(), ctx->pack_off.as<uint64_t>(), (int)n_res + 1);
                if (rc) return fail_load(rc, ctx->err);
                std::vector<int32_t> h_src(n_res);
                CL(cudaMemcpyAsync(h_src.data(), ctx->pack_src.p, (size_t)n_res * 4, cudaMemcpyDeviceToHost, st));
                CL(cudaStreamSynchronize(st));
                ctx->cnt.d2h_bytes += (uint64_t)n_res * 4;
                // chunk sizes (peaks) -> ring of two chunks, page-locked on the host and its twin on the device
                std::vector<uint64_t> chunk_first(nck + 1, 0);       // stream offset (peaks) of the first entry of every chunk
                uint64_t max_chunk = 0;
                for (size_t c = 0; c < nck; ++c) {
                    const uint32_t a = (uint32_t)(c * chunk_spectra), cnt = std::min(chunk_spectra, n_res - a);
                    uint64_t sum = 0;
                    for (uint32_t i = a; i < a + cnt; ++i) sum += off[h_src[i] + 1] - off[h_src[i]];
                    chunk_first[c + 1] = chunk_first[c] + sum; max_chunk = std::max(max_chunk, sum);
                }
                const size_t slot_bytes = (size_t)std::max<uint64_t>(max_chunk, 1) * 16;
                if (ctx->pack_host_cap < 2 * slot_bytes) {
                    if (ctx->pack_host) { cudaFreeHost(ctx->pack_host); ctx->pack_host = nullptr; ctx->pack_host_cap = 0; }
                    CL(cudaHostAlloc(&ctx->pack_host, 2 * slot_bytes + (slot_bytes >> 2), cudaHostAllocDefault)); ctx->pack_host_cap = 2 * slot_bytes + (slot_bytes >> 2);
                }
                CL(ctx->pack_dev.reserve(2 * slot_bytes));
                for (cudaEvent_t& e : ctx->pack_ev) if (!e) CL(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                int n_threads = 8;                                     // measured at C2 size: 4 -> 101 ms end to end, 8 -> 67, 12 -> 65, 16 -> 66
                if (const char* e = getenv("PQ_PACK_THREADS")) { int v = atoi(e); if (v > 0) n_threads = v; }
                n_threads = std::max(1, std::min(n_threads, (int)std::thread::hardware_concurrency()));
                const size_t half = slot_bytes / 16;                   // peaks per slot: [m/z of the chunk | intensities of the chunk]
                // (No spinning wait kernels here: this producer needs the host to launch its copies, and a thread that sits in a device-wide
                // synchronisation — cudaFree of a growing buffer — while a wait kernel spins would block those launches for ever.  Events
                // instead: this thread records chunk c's event, then makes the compute stream wait for it, queues K0 and says so in `issued`.)
                stager.th = std::thread([ctx, st, sc, sp, pc, eager, stream_search, nck, chunk_spectra, n_res, h_src = std::move(h_src), chunk_first, off, mz, inten, slot_bytes, half, n_threads, &stager]() {
                    cudaSetDevice(ctx->P.device);
                    struct Fail { Stager* s; ~Fail() { if (!s->err.empty()) s->failed.store(true, std::memory_order_release); } } fail_guard{&stager};
                    for (size_t c = 0; c < nck; ++c) {
                        const uint32_t a = (uint32_t)(c * chunk_spectra), cnt = std::min(chunk_spectra, n_res - a);
                        const size_t slot = c & 1;
                        if (c >= 2 && cudaEventSynchronize(ctx->pack_ev[slot]) != cudaSuccess) { stager.err = "pack ring: event"; return; }      // the slot's previous chunk has been scattered
                        double* h_mz = reinterpret_cast<double*>(static_cast<uint8_t*>(ctx->pack_host) + slot * slot_bytes);
                        double* h_in = h_mz + half;
                        const uint64_t base = chunk_first[c], peaks = chunk_first[c + 1] - chunk_first[c];
                        // host threads copy contiguous runs of entries; the offsets follow from a prefix over the chunk
                        std::vector<uint64_t> pos(cnt + 1, 0);
                        for (uint32_t i = 0; i < cnt; ++i) pos[i + 1] = pos[i] + (off[h_src[a + i] + 1] - off[h_src[a + i]]);
                        auto work = [&](int t) {
                            const uint32_t lo = (uint32_t)((uint64_t)cnt * t / n_threads), hi = (uint32_t)((uint64_t)cnt * (t + 1) / n_threads);
                            for (uint32_t i = lo; i < hi; ++i) {
                                const uint64_t so = off[h_src[a + i]], k = pos[i + 1] - pos[i];
                                memcpy(h_mz + pos[i], mz + so, (size_t)k * 8); memcpy(h_in + pos[i], inten + so, (size_t)k * 8);
                            }
                        };
                        std::vector<std::thread> pool;
                        for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
                        work(0);
                        for (auto& t : pool) t.join();
                        double* d_mz = reinterpret_cast<double*>(ctx->pack_dev.as<uint8_t>() + slot * slot_bytes);
                        double* d_in = d_mz + half;
                        if (peaks) {
                            if (cudaMemcpyAsync(d_mz, h_mz, (size_t)peaks * 8, cudaMemcpyHostToDevice, sc) != cudaSuccess ||
                                cudaMemcpyAsync(d_in, h_in, (size_t)peaks * 8, cudaMemcpyHostToDevice, sc) != cudaSuccess) { stager.err = "pack ring: copy"; return; }
                        }
                        stager.launches += (uint64_t)launch_scatter_staged(ctx->gather_list.as<int32_t>() + a, (int32_t)cnt, ctx->pack_off.as<uint64_t>() + a, base, ctx->s_off.as<uint64_t>(),
                                                                            d_mz, d_in, ctx->s_mz.as<double>(), ctx->s_in.as<double>(), ctx->chunk_done.as<uint32_t>() + c, 1u, sc);
                        if (cudaEventRecord(ctx->pack_ev[slot], sc) != cudaSuccess || cudaEventRecord(ctx->chunk_ev[c], sc) != cudaSuccess ||
                            cudaStreamWaitEvent(st, ctx->chunk_ev[c], 0) != cudaSuccess) { stager.err = "pack ring: record"; return; }
                        if (eager)
                            stager.launches += (uint64_t)launch_prepare(sp, ctx->gather_list.as<int32_t>() + a, (int32_t)cnt, pc, ctx->blobs.as<uint8_t>(), ctx->blob_off.as<uint64_t>(),
                                                                         ctx->blob_len.as<uint32_t>(), ctx->max_peaks, st, 0, ctx->prep_id.as<int32_t>(), 0, ctx->k0_counters.as<uint32_t>() + 4 * c);
                        if (stream_search && cudaEventRecord(ctx->k0_ev[c], st) != cudaSuccess) { stager.err = "pack ring: K0 event"; return; }
                        stager.issued.store(c + 1, std::memory_order_release);
                    }
                });
            } else if (one_gather)
                k_launch += launch_gather_peaks_chunked(ctx->gather_list.as<int32_t>(), (int32_t)n_res, (int32_t)chunk_spectra, ctx->s_caller.as<int32_t>(), in_off.as<uint64_t>(),
                                                        ctx->s_off.as<uint64_t>(), static_cast<const double*>(d_h_mz), static_cast<const double*>(d_h_in),
                                                        ctx->s_mz.as<double>(), ctx->s_in.as<double>(), ctx->chunk_done.as<uint32_t>(), &gather_ctas, sc);
            for (size_t c = 0; c < nck && !cf_staged; ++c) {
                const uint32_t a = (uint32_t)(c * chunk_spectra), cnt = std::min(chunk_spectra, n_res - a);
                if (one_gather) k_launch += launch_wait_chunk(ctx->chunk_done.as<uint32_t>() + c, gather_ctas, st);
                else {
                    k_launch += launch_gather_peaks_from_host(ctx->gather_list.as<int32_t>() + a, (int32_t)cnt, ctx->s_caller.as<int32_t>(), in_off.as<uint64_t>(),
                                                              ctx->s_off.as<uint64_t>(), static_cast<const double*>(d_h_mz), static_cast<const double*>(d_h_in),
                                                              ctx->s_mz.as<double>(), ctx->s_in.as<double>(), sms, sc);
                    CL(cudaEventRecord(ctx->chunk_ev[c], sc));
                    if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, sc); tr_g.push_back(e); }
                    CL(cudaStreamWaitEvent(st, ctx->chunk_ev[c], 0));
                }
                if (eager)          // blob id = rank in mass order (prep_id); one K0 block per SM less than would fit: the next chunk's gather runs beside it
                    k_launch += launch_prepare(sp, ctx->gather_list.as<int32_t>() + a, (int32_t)cnt, pc, ctx->blobs.as<uint8_t>(), ctx->blob_off.as<uint64_t>(),
                                               ctx->blob_len.as<uint32_t>(), ctx->max_peaks, st, 0, ctx->prep_id.as<int32_t>(), 1, ctx->k0_counters.as<uint32_t>() + 4 * c);
                if (stream_search) CL(cudaEventRecord(ctx->k0_ev[c], st));
                if (trace) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tr_k.push_back(e); }
            }
            if (trace) {
                cudaStreamSynchronize(sc); cudaStreamSynchronize(st);
                for (size_t c = 0; c < tr_g.size(); ++c) {
                    float g = 0, k = 0; cudaEventElapsedTime(&g, tr0, tr_g[c]); cudaEventElapsedTime(&k, tr0, tr_k[c]);
                    fprintf(stderr, "[pq] load trace: chunk %3zu gather done %8.3f ms  K0 done %8.3f ms\n", c, g, k);
                    cudaEventDestroy(tr_g[c]); cudaEventDestroy(tr_k[c]);
                }
                cudaEventDestroy(tr0);
            }
            if (eager) { ctx->prepared_at_load = true; ctx->n_prepared = n_res; ctx->cnt.spectra_prepared = n_res; ctx->max_blob = blob_bytes(ctx->max_peaks); }
        }
        mk.mark("gather + K0 queued");
        if (early.th.joinable()) early.th.join();
        if (early.rc) return fail_load(early.rc, early.err.empty() ? ctx->err : early.err);
        ctx->cnt.other_kernel_launches += early.launches; ctx->cnt.d2h_bytes += early.d2h;
        if (early.digest) ctx->ref_digested = true;
        if (early.shuffles) {
            ctx->shuffles_all = true; ctx->shuffles_epoch = ctx->targets_epoch; ctx->shuf_pending = true;
            CL(cudaStreamWaitEvent(st, ctx->shuf_ev, 0));      // whatever is queued on the main stream from here on sees the shuffles
        }
        mk.mark("early work joined");
        if (stream_search && n_res) {
            // Steps 2-4 part by part while the rest of the spectra are still arriving
            LoadPlan plan{hj.njobs, hj.cand, n_res, chunk_spectra_cf, n_chunks_cf, &ctx->k0_ev};
            if (cf_staged) { plan.issued = &stager.issued; plan.failed = &stager.failed; }
            ctx->cnt.other_kernel_launches += (uint64_t)k_launch; k_launch = 0;
            rc = search_at_load(ctx, plan);
            if (rc) return fail_load(rc, ctx->err);
            searched = true;
        }
    } else {
        // ---- full upload: every peak, in chunks of whole caller spectra ----
        const uint64_t total_sorted = total + alias_peaks;
        CL(in_mz.reserve(std::max<uint64_t>(total, 1) * 8)); CL(in_in.reserve(std::max<uint64_t>(total, 1) * 8));
        CL(ctx->s_mz.reserve(std::max<uint64_t>(total_sorted, 1) * 8)); CL(ctx->s_in.reserve(std::max<uint64_t>(total_sorted, 1) * 8));
        sp = ctx->view();
        while (ctx->chunk_ev.size() < n_chunks + 1) { cudaEvent_t e; CL(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->chunk_ev.push_back(e); }
        {
            Timer t(ctx, &ctx->cnt.ms_prepare);
            k_launch += launch_resident_counts(nv, ctx->s_npeaks.as<uint32_t>(), nullptr, ctx->scratch.as<uint64_t>(), st);
            int32_t rc = scan_exclusive(ctx, ctx->scratch.as<uint64_t>(), ctx->s_off.as<uint64_t>(), (int)nv + 1);
            if (rc) return fail_load(rc, ctx->err);
        }
        // The peak copies are queued after the small precursor work: queued ahead of it (pageable precursor arrays are staged by the
        // driver) they were seen to hold those kernels back until the last chunk had landed (60 ms per load instead of 36 ms).
        for (size_t c = 0; c < n_chunks; ++c) {
            const uint64_t a = off[cut[c]], b = off[cut[c + 1]];
            if (b > a && !peaks_on_device) {
                CL(cudaMemcpyAsync(in_mz.as<double>() + a, mz + a, (b - a) * 8, cudaMemcpyHostToDevice, sc));
                CL(cudaMemcpyAsync(in_in.as<double>() + a, inten + a, (b - a) * 8, cudaMemcpyHostToDevice, sc));
            }
            CL(cudaEventRecord(ctx->chunk_ev[c], sc));
        }
        mk.mark("chunk copies queued");
        ctx->cnt.h2d_bytes += (uint64_t)n * 28 + 8 + (uint64_t)n0 * 4 + (peaks_on_device ? 0 : total * 16);
        ctx->residency = 0; ctx->n_resident = (uint32_t)nv;
        {
            Timer t(ctx, &ctx->cnt.ms_prepare);
            // prepare_at_load: blob v belongs to virtual spectrum v (the list K0 walks is s_inv), sizes from the peak counts
            if (eager) {
                CL(ctx->blob_off.reserve((size_t)(nv + 2) * 8)); CL(ctx->blob_len.reserve((size_t)(nv + 2) * 4));
                CL(ctx->prep_list.reserve((size_t)(nv + 2) * 8));
                uint64_t* sizes = ctx->prep_list.as<uint64_t>();
                k_launch += launch_blob_sizes(sp, ctx->s_inv.as<int32_t>(), (int32_t)nv, sizes, st);
                int32_t rc = scan_exclusive(ctx, sizes, ctx->blob_off.as<uint64_t>(), (int)nv + 1);
                if (rc) return fail_load(rc, ctx->err);
                uint64_t total_blob = 0;
                CL(cudaMemcpyAsync(&total_blob, ctx->blob_off.as<uint64_t>() + nv, 8, cudaMemcpyDeviceToHost, st));
                CL(cudaStreamSynchronize(st));
                if (ctx->blobs.cap < total_blob + 256) {                // growing: only if it fits beside what is still to come
                    size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b);
                    if ((double)(total_blob + 256 - ctx->blobs.cap) > 0.6 * (double)free_b) eager = false;
                }
                if (eager && ctx->blobs.reserve(total_blob + 256) != cudaSuccess) { cudaGetLastError(); eager = false; }
            }
            size_t z_lo = 0;                                             // charge-less spectra (ascending caller index) of the chunks seen so far
            for (size_t c = 0; c < n_chunks; ++c) {
                CL(cudaStreamWaitEvent(st, ctx->chunk_ev[c], 0));
                const int64_t a = cut[c], cnt = cut[c + 1] - cut[c];
                size_t z_hi = z_lo; while (z_hi < zeros.size() && zeros[z_hi] < cut[c + 1]) ++z_hi;
                // the chunk's spectra are virtual a .. a+cnt-1, their z = 3 readings virtual n+z_lo .. n+z_hi-1
                const int64_t first[2] = {a, n + (int64_t)z_lo}, count[2] = {cnt, (int64_t)(z_hi - z_lo)};
                for (int part = 0; part < 2; ++part) {
                    if (count[part] <= 0) continue;
                    k_launch += launch_gather_peaks_by_virtual(first[part], count[part], n, ctx->d_zero.as<int32_t>(), ctx->s_inv.as<int32_t>(), in_off.as<uint64_t>(),
                                                               ctx->s_off.as<uint64_t>(), in_mz.as<double>(), in_in.as<double>(), ctx->s_mz.as<double>(), ctx->s_in.as<double>(), st);
                    if (eager)
                        k_launch += launch_prepare(sp, ctx->s_inv.as<int32_t>() + first[part], (int32_t)count[part], pc, ctx->blobs.as<uint8_t>(),
                                                   ctx->blob_off.as<uint64_t>() + first[part], ctx->blob_len.as<uint32_t>() + first[part], ctx->max_peaks, st,
                                                   (uint32_t)std::max(min_peaks_eff(ctx) + 1, 0));
                }
                z_lo = z_hi;
            }
            if (eager) { ctx->prepared_at_load = true; ctx->n_prepared = (uint32_t)nv; ctx->cnt.spectra_prepared = usable; ctx->max_blob = blob_bytes(ctx->max_peaks); }
        }
        mk.mark("chunk work queued");
    }
    ctx->cnt.other_kernel_launches += (uint64_t)k_launch;
    CL(cudaGetLastError());
    // a protein database staged before the spectra (pq_stage_reference): its digest does not depend on them and runs now, on a
    // third stream, beside the work above
    if (digest_here && !searched && !ctx->ref_digested) {
        if (cf) { std::string err; if (!upload_target_keys_for_reference(ctx->st_aux, ctx->targets.seqs, ctx->reference.dev, err)) return fail_load(PQ_ERR_CUDA, "reference digest: " + err); }
        int32_t rc = digest_staged_reference(ctx, ctx->st_aux); if (rc) return fail_load(rc, ctx->err);
    }
    mk.mark("digest (staged reference)");
    // targets given before the spectra: their table is built while the peaks are on their way (device work, a few launches)
    if (ctx->have_targets && !cf) { int32_t rc = finalize_targets(ctx); if (rc) return fail_load(rc, ctx->err); }
    if (stager.th.joinable()) stager.th.join();
    if (!stager.err.empty()) return fail_load(PQ_ERR_CUDA, stager.err);
    ctx->cnt.other_kernel_launches += stager.launches;
    CL(cudaStreamSynchronize(st)); CL(cudaStreamSynchronize(sc));
    mk.mark("wait for the upload");
#undef CL
    return PQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int32_t pq_set_targets(pq_ctx* ctx, int32_t n_seq, const uint32_t* off, const char* res) {
    if (!ctx || n_seq < 0 || (n_seq > 0 && (!off || !res))) return fail(ctx, PQ_ERR_BAD_ARG, "null argument");
    cudaSetDevice(ctx->P.device);
    PeptideTable& T = ctx->targets;
    T.seqs.clear(); T.forms.clear(); ctx->t_off.assign(1, 0u); ctx->t_flat.clear();
    ctx->have_targets = false; ctx->target_table_ready = false; ctx->n_target_forms = 0; ++ctx->targets_epoch; ctx->n_targeted = 0; ctx->tg_pos.clear();
    for (int32_t i = 0; i < n_seq; ++i) {
        if (off[i + 1] < off[i]) return fail(ctx, PQ_ERR_BAD_ARG, "target offsets must be non-decreasing");
        std::string s(res + off[i], res + off[i + 1]);
        if (s.size() < 2 || s.size() > (size_t)kMaxLen) return fail(ctx, PQ_ERR_UNSUPPORTED, "target length must be 2..60: " + s);
        for (char c : s) if (c < 'A' || c > 'Z' || ctx->codes.residue_mass[c - 'A'] == 0.0) return fail(ctx, PQ_ERR_BAD_ARG, "target has a non-standard residue: " + s);
        T.seqs.push_back(s); ctx->t_flat += s; ctx->t_off.push_back((uint32_t)ctx->t_flat.size());
    }
    // The isoform expansion and the mass-sorted device table are built on the device by the next pq_load_spectra (it needs the
    // target masses to decide which spectra to bring over) or, failing that, by pq_step2_match_targets: finalize_targets().
    ctx->forms_expanded = false; ctx->cnt.target_forms = 0; ctx->ref_digested = false;     // the digest is 'minus the targets'
    ctx->have_targets = true; ctx->step_done = 0; reset_psm_state(ctx);
    return PQ_OK;
}

}  // extern "C"

namespace {
std::string mod_string(const pq_ctx* ctx, const FormRec& f) {
    if (f.n == 0) return "-";
    std::string s; char buf[96];
    for (int i = 0; i < f.n; ++i) {
        snprintf(buf, sizeof buf, "%d@%d[%.4f]", ctx->codes.mods[(size_t)f.mod[i]].id, (int)f.site[i], ctx->codes.mods[(size_t)f.mod[i]].delta);
        if (i) s += ";";
        s += buf;
    }
    return s;
}
// target forms keep two orders: public (generation) and device (mass-sorted)
// caller-visible form order = generation order (sequence order, then calcPeptideIsoforms order); the device table is the
// same forms sorted by mass; row_of_form / form_of_row translate.
void ensure_forms(pq_ctx* ctx) {
    if (ctx->forms_expanded || !ctx->have_targets) return;
    PeptideTable& T = ctx->targets;
    T.forms.clear();
    for (size_t i = 0; i < T.seqs.size(); ++i) ctx->codes.expand_forms(T.seqs[i], (int32_t)i, T.forms);
    ctx->cnt.target_forms = T.forms.size();
    ctx->forms_expanded = true;
}

// The target table (PeptideInput.calcPeptideIsoforms for every target, mass-sorted rows, side arrays of K7) is built on the
// device (k6_reference.cu, build_target_table_on_device).  Queued on the context stream; synchronised once inside (form count).
int32_t finalize_targets(pq_ctx* ctx) {
    if (ctx->target_table_ready) return PQ_OK;
    PeptideTable& T = ctx->targets;
    Marks mk("finalize_targets");
    TargetDeviceTable& D = ctx->tdev;
    D.rows = &T.rows; D.mass = &T.mass; D.form_of_row = &ctx->t_form_of_row; D.seq_of_form = &ctx->t_seq_of_form; D.form_mods = &ctx->t_form_mods;
    D.seq_off = &ctx->t_seq_off; D.seq_letters = &ctx->t_seq_letters;
    std::string err; uint64_t launches = 0;
    if (!build_target_table_on_device(ctx->st, ctx->codes, (uint32_t)T.seqs.size(), ctx->t_off.data(), ctx->t_flat.data(), D, &launches, err))
        return fail(ctx, PQ_ERR_CUDA, "target table: " + err);
    ctx->cnt.other_kernel_launches += launches;
    const size_t n = D.n_forms;
    ctx->n_target_forms = D.n_forms; T.n_rows = n; T.device_built = true; ctx->cnt.target_forms = n;
    ctx->cnt.h2d_bytes += ctx->t_flat.size() * 2 + T.seqs.size() * 16 + 4;
    if (ctx->P.score_method == 2 && n > 0) {
        CU(T.pred.reserve(n * 8));
        ctx->cnt.other_kernel_launches += (uint64_t)launch_predict(T.rows.as<uint8_t>(), (uint32_t)n, ctx->d_mods.as<ModTable>(), ctx->P.itol, T.pred.as<int32_t>(), ctx->st);
    }
    mk.mark("device build (queued)");
    ctx->target_table_ready = true;
    return PQ_OK;
}
}  // namespace

extern "C" {

int32_t pq_set_targeted_psms(pq_ctx* ctx, int64_t n, const int32_t* target_seq, const int32_t* spectrum) {
    if (!ctx || n < 0 || (n > 0 && (!target_seq || !spectrum))) return fail(ctx, PQ_ERR_BAD_ARG, "null argument");
    if (!ctx->have_targets) return fail(ctx, PQ_ERR_STATE, "pq_set_targets first");
    if (n > 0x7fffffffLL) return fail(ctx, PQ_ERR_UNSUPPORTED, "too many targeted pairs");
    cudaSetDevice(ctx->P.device);
    ctx->n_targeted = 0; ctx->tg_pos.clear();
    ctx->step_done = 0; reset_psm_state(ctx);                     // Step 2 has to run (again) with the list
    if (n == 0) return PQ_OK;
    const auto& seqs = ctx->targets.seqs;
    const size_t nseq = seqs.size();
    // the reference keys its map by the sequence text: targets given twice share one entry
    std::vector<uint32_t> canon(nseq); std::unordered_map<std::string, uint32_t> first;
    for (size_t i = 0; i < nseq; ++i) canon[i] = first.emplace(seqs[i], (uint32_t)i).first->second;
    std::vector<uint64_t> keys((size_t)n); std::vector<uint8_t> listed(nseq, 0);
    for (int64_t i = 0; i < n; ++i) {
        if (target_seq[i] < 0 || (size_t)target_seq[i] >= nseq || spectrum[i] < 0) return fail(ctx, PQ_ERR_BAD_ARG, "targeted pair " + std::to_string(i) + " out of range");
        const uint32_t q = canon[(size_t)target_seq[i]];
        keys[(size_t)i] = ((uint64_t)q << 32) | (uint32_t)spectrum[i]; listed[q] = 1;
    }
    std::vector<uint64_t> sorted(keys); std::sort(sorted.begin(), sorted.end()); sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
    ctx->tg_pos.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) ctx->tg_pos[(size_t)i] = (uint32_t)(std::lower_bound(sorted.begin(), sorted.end(), keys[(size_t)i]) - sorted.begin());
    const size_t m = sorted.size();
    CU(ctx->tg_keys.reserve(m * 8)); CU(ctx->tg_listed.reserve(nseq + 64)); CU(ctx->tg_canon.reserve(nseq * 4 + 64)); CU(ctx->tg_seen.reserve(m + 64));
    CU(ctx->tg_best.reserve(m * 8)); CU(ctx->tg_type.reserve(m * 4));
    CU(cudaMemcpyAsync(ctx->tg_keys.p, sorted.data(), m * 8, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(ctx->tg_listed.p, listed.data(), nseq, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemcpyAsync(ctx->tg_canon.p, canon.data(), nseq * 4, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaMemsetAsync(ctx->tg_seen.p, 0, m, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->cnt.h2d_bytes += m * 8 + nseq * 5;
    ctx->n_targeted = (uint32_t)m;
    return PQ_OK;
}

int32_t pq_get_targeted_psm_status(pq_ctx* ctx, int32_t* type, int64_t capacity, int64_t* n_pairs) {
    if (!ctx || !n_pairs) return PQ_ERR_BAD_ARG;
    *n_pairs = (int64_t)ctx->tg_pos.size();
    if (!type) return PQ_OK;
    if (capacity < *n_pairs) return fail(ctx, PQ_ERR_CAPACITY, "status buffer too small");
    if (ctx->tg_pos.empty()) return PQ_OK;
    if (ctx->step_done < 2) return fail(ctx, PQ_ERR_STATE, "pq_step2_match_targets first");
    cudaSetDevice(ctx->P.device);
    ctx->cnt.other_kernel_launches += (uint64_t)launch_targeted_status(ctx->d_psm.as<DevPsm>(), ctx->n_psm, ctx->targeted_view(), ctx->tg_best.as<unsigned long long>(),
                                                                       ctx->tg_type.as<int32_t>(), ctx->st);
    std::vector<int32_t> h(ctx->n_targeted);
    CU(cudaMemcpyAsync(h.data(), ctx->tg_type.p, h.size() * 4, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->cnt.d2h_bytes += h.size() * 4;
    for (size_t i = 0; i < ctx->tg_pos.size(); ++i) type[i] = h[ctx->tg_pos[i]];
    return PQ_OK;
}

int32_t pq_get_target_forms(pq_ctx* ctx, pq_form* out, int64_t capacity, int64_t* n_forms) {
    if (!ctx || !n_forms) return PQ_ERR_BAD_ARG;
    PeptideTable& T = ctx->targets;
    ensure_forms(ctx);
    *n_forms = (int64_t)T.forms.size();
    if (!out) return PQ_OK;
    if (capacity < *n_forms) return fail(ctx, PQ_ERR_CAPACITY, "form buffer too small");
    for (size_t i = 0; i < T.forms.size(); ++i) ctx->codes.to_public(T.seqs[(size_t)T.forms[i].seq], T.forms[i], out + i);
    return PQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int32_t pq_step2_match_targets(pq_ctx* ctx, int64_t* n_psms) {
    if (!ctx) return PQ_ERR_BAD_ARG;
    if (!ctx->have_targets) return fail(ctx, PQ_ERR_STATE, "pq_set_targets first");
    cudaSetDevice(ctx->P.device);
    if (ctx->pending_step[2]) { ctx->pending_step[2] = false; if (n_psms) *n_psms = (int64_t)ctx->n_psm; return PQ_OK; }      // searched while the spectra were loaded
    Marks mk("step2");
    int32_t rc = finalize_targets(ctx);
    if (rc) return rc;
    rc = ensure_counters(ctx);
    if (rc) return rc;
    reset_psm_state(ctx); ctx->step_done = 0;
    if (!ctx->prepared_at_load) ctx->n_prepared = 0;
    if (n_psms) *n_psms = 0;
    const int64_t n = ctx->nS;
    PeptideTable& T = ctx->targets;
    if (n == 0 || T.n() == 0) { ctx->step_done = 2; return PQ_OK; }
    cudaStream_t st = ctx->st;
    const uint32_t job_cap = (uint32_t)(n * ctx->P.n_isotope);
    const bool cf = ctx->residency == 1;
    CU(ctx->jobs.reserve((size_t)job_cap * sizeof(Job)));
    CU(ctx->njobs.reserve(64));
    CU(ctx->need.reserve((size_t)(n + 1) * 4));
    if (!cf) { CU(ctx->prep_id.reserve((size_t)(n + 1) * 4)); CU(ctx->prep_list.reserve((size_t)(n + 1) * 4)); }
    CU(cudaMemsetAsync(ctx->njobs.p, 0, 32, st));
    CU(cudaMemsetAsync(ctx->need.p, 0, (size_t)(n + 1) * 4, st));
    if (ctx->n_targeted) CU(cudaMemsetAsync(ctx->tg_seen.p, 0, ctx->n_targeted, st));
    WindowConfig W = window_config(ctx, true);
    SpectraView sp = ctx->view();
    struct { uint32_t njobs, nprep; unsigned long long cand; uint32_t missing, pad; } hn = {0, 0, 0, 0, 0};
    int32_t h_nprep = 0;
    {
        Timer t(ctx, &ctx->cnt.ms_window);
        int k = launch_enumerate_targets(sp, T.mass.as<double>(), (uint32_t)T.n(), W, ctx->jobs.as<Job>(), ctx->njobs.as<uint32_t>(), ctx->need.as<int32_t>(), st);
        if (cf) k += launch_count_missing(n, ctx->need.as<int32_t>(), ctx->s_res.as<int32_t>(), ctx->njobs.as<uint32_t>() + 4, st);
        else { rc = scan_exclusive(ctx, ctx->need.as<int32_t>(), ctx->prep_id.as<int32_t>(), (int)n + 1); if (rc) return rc; }
        ctx->cnt.other_kernel_launches += (uint64_t)k + 1;
    }
    CU(cudaMemcpyAsync(&hn, ctx->njobs.p, 24, cudaMemcpyDeviceToHost, st));
    if (!cf) CU(cudaMemcpyAsync(&h_nprep, ctx->prep_id.as<int32_t>() + n, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ctx->cnt.d2h_bytes += 28;
    const uint32_t h_njobs = hn.njobs;
    const uint64_t cand = hn.cand;
    mk.mark("enumerate");
    if (cf && hn.missing)
        return fail(ctx, PQ_ERR_STATE, "the spectra were loaded candidate-first for another target list: " + std::to_string(hn.missing) +
                                       " spectra in the precursor windows of the present targets have no peaks on the device; call pq_load_spectra again");
    if (h_njobs == 0) { ctx->step_done = 2; return PQ_OK; }
    if (cand > 0xfffffff0ull) return fail(ctx, PQ_ERR_UNSUPPORTED, "too many Step-2 candidates for one launch");
    // list of spectra to prepare (ascending sorted index) + blob offsets
    {
        Timer t(ctx, &ctx->cnt.ms_prepare);
        int k = 0;
        PrepConfig pc; pc.itol = ctx->P.itol; pc.method = ctx->P.score_method;
        pc.mvh_bins = (int32_t)jround((double)(int)(2000.0 - 200.0) / (ctx->P.itol * 2.0));
        if (cf) {
            // candidate-first: the resident spectra are the ones with a target in their window; blob id = rank among them (prep_id of the load)
            if (!ctx->prepared_at_load && ctx->n_resident) {          // prepare_at_load = 0: K0 runs here, once per load
                uint64_t total_blob = 0;
                CU(cudaMemcpyAsync(&total_blob, ctx->blob_off.as<uint64_t>() + ctx->n_resident, 8, cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                CU(ctx->blobs.reserve(total_blob + 256));
                k += launch_prepare(sp, ctx->prep_list.as<int32_t>(), (int32_t)ctx->n_resident, pc, ctx->blobs.as<uint8_t>(), ctx->blob_off.as<uint64_t>(),
                                    ctx->blob_len.as<uint32_t>(), ctx->max_peaks, st);
                ctx->n_prepared = ctx->n_resident; ctx->cnt.spectra_prepared = ctx->n_resident; ctx->max_blob = blob_bytes(ctx->max_peaks);
                if (ctx->P.prepare_at_load != 0) ctx->prepared_at_load = true;      // (the load could not reserve the blobs then; they exist now)
            }
            k += launch_fix_job_blobs(ctx->jobs.as<Job>(), h_njobs, ctx->prep_id.as<int32_t>(), st);
        } else if (ctx->prepared_at_load) {
            // every usable spectrum was prepared during the upload; blob id = virtual id
            k += launch_fix_job_blobs(ctx->jobs.as<Job>(), h_njobs, ctx->s_vid.as<int32_t>(), st);
        } else {
        size_t tmp = 0;
        cub::CountingInputIterator<int32_t> iota(0);
        cub::DeviceSelect::Flagged(nullptr, tmp, iota, ctx->need.as<int32_t>(), ctx->prep_list.as<int32_t>(), ctx->njobs.as<uint32_t>() + 1, (int)n, st);
        CU(ctx->scan_tmp.reserve(tmp + 64));
        CU(cub::DeviceSelect::Flagged(ctx->scan_tmp.p, tmp, iota, ctx->need.as<int32_t>(), ctx->prep_list.as<int32_t>(), ctx->njobs.as<uint32_t>() + 1, (int)n, st));
        CU(ctx->scratch.reserve((size_t)(h_nprep + 2) * 8));
        CU(ctx->blob_off.reserve((size_t)(h_nprep + 2) * 8)); CU(ctx->blob_len.reserve((size_t)(h_nprep + 2) * 4));
        launch_blob_sizes(sp, ctx->prep_list.as<int32_t>(), h_nprep, ctx->scratch.as<uint64_t>(), st);
        rc = scan_exclusive(ctx, ctx->scratch.as<uint64_t>(), ctx->blob_off.as<uint64_t>(), h_nprep + 1);
        if (rc) return rc;
        uint64_t total_blob = 0;
        CU(cudaMemcpyAsync(&total_blob, ctx->blob_off.as<uint64_t>() + h_nprep, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(ctx->blobs.reserve(total_blob + 256));
        k += launch_prepare(sp, ctx->prep_list.as<int32_t>(), h_nprep, pc, ctx->blobs.as<uint8_t>(), ctx->blob_off.as<uint64_t>(), ctx->blob_len.as<uint32_t>(), ctx->max_peaks, st);
        k += launch_fix_job_blobs(ctx->jobs.as<Job>(), h_njobs, ctx->prep_id.as<int32_t>(), st);
        ctx->n_prepared = (uint32_t)h_nprep; ctx->cnt.spectra_prepared = (uint64_t)h_nprep;
        ctx->max_blob = blob_bytes(ctx->max_peaks);
        }
        // order the jobs by (charge, first target row): lanes of a warp then score peptides of equal length and charge
        CU(ctx->k_key_a.reserve((size_t)h_njobs * 8)); CU(ctx->k_key_b.reserve((size_t)h_njobs * 8));
        CU(ctx->k_idx_a.reserve((size_t)h_njobs * 4)); CU(ctx->k_idx_b.reserve((size_t)h_njobs * 4));
        CU(ctx->jobs2.reserve((size_t)h_njobs * sizeof(Job)));
        k += launch_job_keys(ctx->jobs.as<Job>(), h_njobs, ctx->k_key_a.as<uint64_t>(), ctx->k_idx_a.as<uint32_t>(), st);
        rc = sort_pairs(ctx, ctx->k_key_a.as<uint64_t>(), ctx->k_key_b.as<uint64_t>(), ctx->k_idx_a.as<uint32_t>(), ctx->k_idx_b.as<uint32_t>(), (int)h_njobs, 44);
        if (rc) return rc;
        k += launch_gather_jobs(ctx->jobs.as<Job>(), ctx->k_idx_b.as<uint32_t>(), h_njobs, ctx->jobs2.as<Job>(), st);
        std::swap(ctx->jobs, ctx->jobs2);
        ctx->cnt.other_kernel_launches += (uint64_t)k + 3;
    }
    CU(cudaGetLastError());
    mk.mark("prepare launch");
    rc = run_score(ctx, kJobTargets, h_njobs, T.pred.as<int32_t>(), T.rows.as<uint8_t>(), T.mass.as<double>(), (uint32_t)cand, 1);
    if (rc) return rc;
    uint32_t nh = 0;
    CU(cudaMemcpyAsync(&nh, ctx->hit_count.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ctx->cnt.d2h_bytes += 4;
    mk.mark("score (waits for K0 + K1)");
    const Hit* hits = ctx->hits.as<Hit>();
    const uint8_t* skip_job = nullptr;
    if (ctx->nZero > 0) {
        // spectra loaded without a charge: the z = 3 reading only counts when the z = 2 reading saved no PSM
        // (FindCandidateSpectraAndScoreWorker.java:96-228): drop its hits and its pair count otherwise
        CU(ctx->z_hit.reserve((size_t)n + 64)); CU(ctx->z_skip.reserve((size_t)h_njobs + 64)); CU(ctx->z_keep.reserve((size_t)nh + 64)); CU(ctx->hits_f.reserve(((size_t)nh + 1) * sizeof(Hit)));
        CU(cudaMemsetAsync(ctx->z_hit.p, 0, (size_t)n, st));
        int k = launch_chargeless_filter(hits, nh, ctx->jobs.as<Job>(), h_njobs, sp, ctx->nCaller, ctx->d_zero.as<int32_t>(), ctx->s_vid.as<int32_t>(), ctx->s_inv.as<int32_t>(),
                                         ctx->z_hit.as<uint8_t>(), ctx->z_skip.as<uint8_t>(), ctx->z_keep.as<uint8_t>(), st);
        if (nh) {
            size_t tmp = 0;
            uint32_t* d_cnt = ctx->njobs.as<uint32_t>() + 6;
            cub::DeviceSelect::Flagged(nullptr, tmp, hits, ctx->z_keep.as<uint8_t>(), ctx->hits_f.as<Hit>(), d_cnt, (int)nh, st);
            CU(ctx->scan_tmp.reserve(tmp + 64));
            CU(cub::DeviceSelect::Flagged(ctx->scan_tmp.p, tmp, hits, ctx->z_keep.as<uint8_t>(), ctx->hits_f.as<Hit>(), d_cnt, (int)nh, st));
            CU(cudaMemcpyAsync(&nh, d_cnt, 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            hits = ctx->hits_f.as<Hit>();
        }
        skip_job = ctx->z_skip.as<uint8_t>();
        ctx->cnt.other_kernel_launches += (uint64_t)k + 2;
    }
    ctx->cnt.other_kernel_launches += (uint64_t)launch_job_counters(ctx->jobs.as<Job>(), ctx->results.as<JobResult>(), h_njobs, sp.npeaks, 0, ctx->d_counters.as<DevCounters>(), st, skip_job);
    // hits -> PSM table in (mass-sorted spectrum, form) order, spectrum segments, caller-order permutation: all on the device
    if (nh) {
        CU(ctx->d_psm.reserve((size_t)nh * sizeof(DevPsm))); CU(ctx->d_seg_start.reserve(((size_t)nh + 2) * 4)); CU(ctx->d_out_perm.reserve((size_t)nh * 4));
        std::string err; int launches = 0;
        if (!psm_table_from_hits(hits, nh, sp, T.mass.as<double>(), T.rows.as<uint8_t>(), ctx->aux(), ctx->k_key_a, ctx->k_key_b, ctx->k_idx_a,
                                 ctx->k_idx_b, ctx->k_head, ctx->scan_tmp, ctx->d_psm.as<DevPsm>(), ctx->d_seg_start.as<uint32_t>(),
                                 ctx->d_scalars.as<uint32_t>(), ctx->d_out_perm.as<uint32_t>(), st, &launches, err))
            return fail(ctx, PQ_ERR_CUDA, "psm table: " + err);
        ctx->cnt.other_kernel_launches += (uint64_t)launches;
        // segment count + the first / last matched spectrum (mass range of the reference table, GeneratePeptideformWorker.java:40-52)
        uint32_t nseg = 0; int32_t s_first = 0, s_last = 0;
        CU(cudaMemcpyAsync(&nseg, ctx->d_scalars.p, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&s_first, &ctx->d_psm.as<DevPsm>()[0].spectrum_sorted, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&s_last, &ctx->d_psm.as<DevPsm>()[nh - 1].spectrum_sorted, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaMemcpyAsync(&ctx->psm_mass_lo, ctx->s_mass.as<double>() + s_first, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&ctx->psm_mass_hi, ctx->s_mass.as<double>() + s_last, 8, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        ctx->cnt.d2h_bytes += 28;
        ctx->n_psm = nh; ctx->n_seg = nseg;
    }
    ctx->cnt.psms = ctx->n_psm;
    mk.mark("psm table");
    ctx->step_done = 2;
    if (n_psms) *n_psms = (int64_t)ctx->n_psm;
    return PQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

// digest of the staged proteins (K6 part 1) on stream st; synchronised at return
int32_t digest_staged_reference(pq_ctx* ctx, cudaStream_t st, bool count_time) {
    auto t_begin = std::chrono::steady_clock::now();
    PeptideTable& R = ctx->reference;
    std::string err; uint64_t launches = 0;
    if (!digest_reference_on_device(st, ctx->P, ctx->codes, ctx->targets.seqs, R.dev, &launches, err)) return fail(ctx, PQ_ERR_CUDA, "reference digest: " + err);   // (target keys already uploaded)
    ctx->cnt.other_kernel_launches += launches;
    ctx->ref_digested = true;
    if (count_time) ctx->cnt.ms_host_reference += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return PQ_OK;
}
}  // namespace

extern "C" {

int32_t pq_stage_reference(pq_ctx* ctx, int32_t n_proteins, const uint64_t* off, const char* res) {
    if (!ctx || n_proteins < 0 || (n_proteins > 0 && (!off || !res))) return fail(ctx, PQ_ERR_BAD_ARG, "null argument");
    cudaSetDevice(ctx->P.device);
    auto t_begin = std::chrono::steady_clock::now();
    ctx->ref_staged = false; ctx->ref_digested = false; ctx->have_reference = false;
    ctx->pending_ref = false; ctx->pending_step[3] = ctx->pending_step[4] = false; ctx->step_done = std::min(ctx->step_done, 2);      // results of another database
    std::string err;
    if (!upload_proteins_for_reference(ctx->st, n_proteins, off, res, ctx->reference.dev, err)) return fail(ctx, PQ_ERR_CUDA, "reference upload: " + err);
    ctx->cnt.h2d_bytes += (n_proteins ? off[n_proteins] : 0) + ((uint64_t)n_proteins + 1) * 8;
    ctx->ref_staged = true;
    ctx->cnt.ms_host_reference += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return PQ_OK;
}

int32_t pq_build_reference(pq_ctx* ctx, int32_t n_proteins, const uint64_t* off, const char* res) {
    if (!ctx || n_proteins < 0 || (n_proteins > 0 && (!off || !res))) return fail(ctx, PQ_ERR_BAD_ARG, "null argument");
    if (ctx->step_done < 2) return fail(ctx, PQ_ERR_STATE, "pq_step2_match_targets first (the reference mass window comes from the matched spectra)");
    const bool use_staged = n_proteins == 0 && !off && !res && ctx->ref_staged;      // the database handed over by pq_stage_reference
    cudaSetDevice(ctx->P.device);
    if (use_staged && ctx->pending_ref && ctx->have_reference) { ctx->pending_ref = false; return PQ_OK; }      // cut out of the superset table when the spectra were loaded
    ctx->pending_ref = false; ctx->pending_step[3] = ctx->pending_step[4] = false; ctx->step_done = std::min(ctx->step_done, 2);
    PeptideTable& R = ctx->reference;
    R.seqs.clear(); R.forms.clear(); R.have_mask = false; R.have_lad = false; R.device_built = false; R.materialized = true;
    auto t_begin = std::chrono::steady_clock::now();
    Marks mk("build_reference");
    if (!use_staged) { ctx->ref_staged = false; ctx->ref_digested = false; }
    {
        // device path (K6): digest, unique, form expansion and the mass sort all run on the GPU
        double mn = 1.7976931348623157e308, mx = 0.0;      // GeneratePeptideformWorker.set_mass_range (:40-52)
        if (ctx->n_psm) { mn = ctx->psm_mass_lo; mx = ctx->psm_mass_hi; }
        const double lo = mn - 250.0 - 5.0, hi = mx + 250.0 + 5.0;
        R.dev.rows = &R.rows; R.dev.mass = &R.mass;
        std::string err;
        uint64_t launches = 0;
        if (use_staged) {
            if (!ctx->ref_digested) {
                if (!upload_target_keys_for_reference(ctx->st, ctx->targets.seqs, R.dev, err)) return fail(ctx, PQ_ERR_CUDA, "reference digest: " + err);
                int32_t rc = digest_staged_reference(ctx, ctx->st, false); if (rc) return rc;
            }
            if (!expand_reference_on_device(ctx->st, ctx->P, ctx->codes, lo, hi, R.dev, &launches, err)) return fail(ctx, PQ_ERR_CUDA, "reference build: " + err);
        } else {
            if (!build_reference_on_device(ctx->st, ctx->P, ctx->codes, n_proteins, off, res, ctx->targets.seqs, lo, hi, R.dev, &launches, err))
                return fail(ctx, PQ_ERR_CUDA, "reference build: " + err);
            ctx->cnt.h2d_bytes += (n_proteins ? off[n_proteins] : 0) + ((uint64_t)n_proteins + 1) * 8;
        }
        ctx->cnt.other_kernel_launches += launches;
        R.device_built = true; R.materialized = false; R.n_rows = R.dev.n_forms;
        mk.mark("device build");
        R.h_mass.clear();                 // the host copy of the masses is pulled by pq_get_reference_forms when it is asked for
        if (R.n_rows) {
            if (ctx->P.score_method == 2) {
                CU(R.pred.reserve(R.n_rows * 8));
                ctx->cnt.other_kernel_launches += (uint64_t)launch_predict(R.rows.as<uint8_t>(), (uint32_t)R.n_rows, ctx->d_mods.as<ModTable>(), ctx->P.itol, R.pred.as<int32_t>(), ctx->st);
            }
            CU(cudaStreamSynchronize(ctx->st));
        }
        mk.mark("predicted peaks");
        ctx->cnt.reference_sequences = R.dev.n_unique; ctx->cnt.reference_forms = R.n_rows;
        ctx->cnt.ms_host_reference += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
        ctx->have_reference = true;
        return PQ_OK;
    }
}

int32_t pq_get_reference_forms(pq_ctx* ctx, pq_form* out, int64_t capacity, int64_t* n_forms) {
    if (!ctx || !n_forms) return PQ_ERR_BAD_ARG;
    PeptideTable& T = ctx->reference;
    *n_forms = (int64_t)T.n();
    if (!out) return PQ_OK;
    if (capacity < *n_forms) return fail(ctx, PQ_ERR_CAPACITY, "form buffer too small");
    if (T.device_built) {
        // the table lives on the device only; bring the rows back and decode them for the caller
        cudaSetDevice(ctx->P.device);
        const size_t n = T.n();
        if (!n) return PQ_OK;
        std::vector<uint8_t> rows(n * kRowBytes); std::vector<uint32_t> sid(n);
        if (T.h_mass.size() != n) {
            T.h_mass.resize(n);
            CU(cudaMemcpyAsync(T.h_mass.data(), T.mass.p, n * 8, cudaMemcpyDeviceToHost, ctx->st));
            ctx->cnt.d2h_bytes += n * 8;
        }
        CU(cudaMemcpyAsync(rows.data(), T.rows.p, rows.size(), cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(sid.data(), T.dev.seq_id.p, n * 4, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        ctx->cnt.d2h_bytes += rows.size() + n * 4;
        std::string seq; FormRec fr;
        for (size_t i = 0; i < n; ++i) {
            ctx->codes.decode_row(rows.data() + i * kRowBytes, seq, fr);
            fr.seq = (int32_t)sid[i]; fr.mass = T.h_mass[i];
            ctx->codes.to_public(seq, fr, out + i);
        }
        return PQ_OK;
    }
    for (size_t i = 0; i < T.forms.size(); ++i) ctx->codes.to_public(T.seqs[(size_t)T.forms[i].seq], T.forms[i], out + i);
    return PQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int32_t pq_step3_competitors(pq_ctx* ctx) {
    if (!ctx) return PQ_ERR_BAD_ARG;
    if (ctx->step_done < 2 || !ctx->have_reference) return fail(ctx, PQ_ERR_STATE, "pq_step2_match_targets and pq_build_reference first");
    cudaSetDevice(ctx->P.device);
    if (ctx->pending_step[3]) { ctx->pending_step[3] = false; ctx->step_done = std::max(ctx->step_done, 3); return PQ_OK; }      // searched while the spectra were loaded
    ctx->pending_step[4] = false; ctx->step_done = std::min(ctx->step_done, 3);
    Marks mk("step3");
    PeptideTable& R = ctx->reference;
    cudaStream_t st = ctx->st;
    ctx->nh3 = 0; ctx->comp3_ready = false; ctx->comp3_n = 0;
    const uint32_t n_seg = ctx->n_seg;
    if (n_seg == 0 || R.n() == 0) {
        // no matched spectrum, or an empty reference table: every PSM keeps n_db = total_db = 0
        if (n_seg) { CU(ctx->d_agg.reserve((size_t)n_seg * sizeof(SegAgg))); CU(cudaMemsetAsync(ctx->d_agg.p, 0, (size_t)n_seg * sizeof(SegAgg), st)); }
        ctx->step_done = 3;
        return PQ_OK;
    }
    WindowConfig W = window_config(ctx, false);
    const uint32_t n_jobs = n_seg * (uint32_t)ctx->P.n_isotope;
    CU(ctx->jobs.reserve((size_t)n_jobs * sizeof(Job)));
    CU(ctx->d_agg.reserve((size_t)n_seg * sizeof(SegAgg)));
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(ctx->d_scalars.as<uint8_t>() + 64);
    CU(cudaMemsetAsync(d_total, 0, 8, st));
    // one job per (matched spectrum, isotope error); target best score per spectrum = DatabaseInput.get_spectra2score "max"
    ctx->cnt.other_kernel_launches += (uint64_t)launch_step3_jobs(ctx->d_psm.as<DevPsm>(), ctx->d_seg_start.as<uint32_t>(), n_seg, R.mass.as<double>(), (uint32_t)R.n(),
                                                                  W, ctx->s_mass.as<double>(), ctx->jobs.as<Job>(), ctx->d_agg.as<SegAgg>(), d_total, st);
    unsigned long long cand = 0;
    CU(cudaMemcpyAsync(&cand, d_total, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ctx->cnt.d2h_bytes += 8;
    mk.mark("build jobs");
    if (cand > 0xfffffff0ull) return fail(ctx, PQ_ERR_UNSUPPORTED, "too many Step-3 candidates for one launch");
    int32_t rc = ensure_ladders(ctx, R, R.rows.as<uint8_t>(), (uint32_t)R.n(), st, true);
    if (rc) return rc;
    rc = run_score(ctx, kJobCompetitors, n_jobs, R.pred.as<int32_t>(), R.rows.as<uint8_t>(), R.mass.as<double>(), (uint32_t)std::max<unsigned long long>(cand, 1), 1, &R);
    if (rc) return rc;
    int k = launch_job_counters(ctx->jobs.as<Job>(), ctx->results.as<JobResult>(), n_jobs, ctx->s_npeaks.as<uint32_t>(), 1, ctx->d_counters.as<DevCounters>(), st);
    k += launch_step3_aggregate(ctx->results.as<JobResult>(), n_seg, ctx->P.n_isotope, ctx->d_seg_start.as<uint32_t>(), ctx->d_psm.as<DevPsm>(), ctx->d_agg.as<SegAgg>(), st);
    ctx->cnt.other_kernel_launches += (uint64_t)k;
    // detail rows stay on the device until pq_get_competitors(3) asks for them: keep this launch's hit rows
    uint32_t nh = 0;
    CU(cudaMemcpyAsync(&nh, ctx->hit_count.p, 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ctx->cnt.d2h_bytes += 4;
    if (nh) { CU(ctx->hits3.reserve((size_t)nh * sizeof(Hit))); CU(cudaMemcpyAsync(ctx->hits3.p, ctx->hits.p, (size_t)nh * sizeof(Hit), cudaMemcpyDeviceToDevice, st)); }
    ctx->nh3 = nh;
    CU(cudaGetLastError());
    mk.mark("score + aggregate");
    ctx->step_done = 3;
    return PQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

template <class K, class V> int32_t sort_pairs_on(pq_ctx* ctx, cudaStream_t st, DevBuf& tmpbuf, const K* kin, K* kout, const V* vin, V* vout, int n, int end_bit) {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, kin, kout, vin, vout, n, 0, end_bit, st);
    CU(tmpbuf.reserve(tmp + 64));
    CU(cub::DeviceRadixSort::SortPairs(tmpbuf.p, tmp, kin, kout, vin, vout, n, 0, end_bit, st));
    return PQ_OK;
}
template <class F> int32_t select_iota_on(pq_ctx* ctx, cudaStream_t st, DevBuf& tmpbuf, const F* flags, uint32_t* out, uint32_t* d_count, int n) {
    size_t tmp = 0;
    cub::CountingInputIterator<uint32_t> iota(0);
    cub::DeviceSelect::Flagged(nullptr, tmp, iota, flags, out, d_count, n, st);
    CU(tmpbuf.reserve(tmp + 64));
    CU(cub::DeviceSelect::Flagged(tmpbuf.p, tmp, iota, flags, out, d_count, n, st));
    return PQ_OK;
}

// Shuffles (K2) for the sequences / forms flagged in sh_seq_flag / sh_form_flag — or, all = true, for every target form with
// all three list charges: plans (group order, round count), layout, generation, de-duplication, modification placement, cell
// lists, MVH predicted-peak counts.  Synchronises st once (sizes).
int32_t generate_shuffles(pq_ctx* ctx, cudaStream_t st, bool all, ShuffleTotals& tot, bool timed, pq_counters* sink) {
    pq_counters& cnt = sink ? *sink : ctx->cnt;      // a helper thread counts into its own block
    const TargetAux A = ctx->aux();
    const size_t nseq = A.n_seq, nform = A.n_forms;
    if (all) {
        CU(ctx->sh_seq_flag.reserve(nseq + 64)); CU(ctx->sh_form_flag.reserve(nform * 4 + 64));
        CU(ctx->sh_plans.reserve(nseq * sizeof(ShuffleSeq))); CU(ctx->sh_seqs.reserve(nseq * sizeof(ShuffleSeq))); CU(ctx->sh_forms.reserve(nform * sizeof(ShuffleForm)));
        CU(ctx->sh_seq_slot.reserve(nseq * 4 + 64)); CU(ctx->sh_form_slot.reserve(nform * 4 + 64)); CU(ctx->sh_row_base.reserve(nform * 4 + 64));
        CU(ctx->sh_totals.reserve(64)); CU(ctx->sh_kcnt.reserve(nseq * 4 + 64));
        launch_fill_u8(ctx->sh_seq_flag.as<uint8_t>(), (uint32_t)nseq, 1, st);
        launch_fill_u32(ctx->sh_form_flag.as<uint32_t>(), (uint32_t)nform, (uint32_t)kListMaxCharge, st);
        CU(cudaMemsetAsync(ctx->sh_totals.p, 0, sizeof(ShuffleTotals), st));
    }
    {
        cudaEvent_t dummy = nullptr; (void)dummy;
        int k = launch_shuffle_plans(ctx->sh_seq_flag.as<uint8_t>(), A, ctx->P.n_random, ctx->sh_plans.as<ShuffleSeq>(), st);
        k += launch_shuffle_layout(ctx->sh_seq_flag.as<uint8_t>(), ctx->sh_form_flag.as<uint32_t>(), nullptr, 0, ctx->sh_plans.as<ShuffleSeq>(), A,
                                   ctx->sh_seq_slot.as<uint32_t>(), ctx->sh_form_slot.as<uint32_t>(), ctx->sh_row_base.as<uint32_t>(), ctx->sh_seqs.as<ShuffleSeq>(),
                                   ctx->sh_forms.as<ShuffleForm>(), ctx->sh_totals.as<ShuffleTotals>(), st);
        cnt.other_kernel_launches += (uint64_t)k;
    }
    CU(cudaMemcpyAsync(&tot, ctx->sh_totals.p, sizeof tot, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cnt.d2h_bytes += sizeof tot;
    ctx->have_shuffles = true; ctx->have_lists = false;
    if (tot.n_seqs == 0) return PQ_OK;
    const uint32_t raw_rows = tot.raw_rows, out_rows = tot.out_rows, max_num = (uint32_t)std::max(ctx->P.n_random, 1);
    CU(ctx->sh_raw.reserve((size_t)std::max(raw_rows, 1u) * kRowBytes)); CU(ctx->sh_kidx.reserve((size_t)std::max(raw_rows, 1u) * 4));
    CU(ctx->sh_rows.reserve((size_t)std::max(out_rows, 1u) * kRowBytes));
    if (ctx->P.score_method == 2) { CU(ctx->sh_pred.reserve((size_t)std::max(out_rows, 1u) * 8)); CU(cudaMemsetAsync(ctx->sh_rows.p, 0, (size_t)out_rows * kRowBytes, st)); }
    // cell lists of the shuffle rows (k_cells.cuh): only the bounded count pass reads them
    ctx->have_lists = !score_config(ctx, 0).score_all && tot.list_units > 0 && !getenv("PQ_NO_CELL_LISTS");
    if (ctx->have_lists && ctx->sh_lists.reserve((size_t)tot.list_units * 16) != cudaSuccess) { cudaGetLastError(); ctx->have_lists = false; }      // no room: K1 walks the ladders
    auto launch = [&]() {
        int k = launch_shuffles(ctx->sh_seqs.as<ShuffleSeq>(), tot.n_seqs, ctx->sh_forms.as<ShuffleForm>(), tot.n_forms,
                                ctx->d_shufmods.as<ShuffleMods>(), ctx->sh_raw.as<uint8_t>(), ctx->sh_kidx.as<uint32_t>(), ctx->sh_kcnt.as<uint32_t>(),
                                ctx->sh_rows.as<uint8_t>(), max_num, ctx->P.shuffle_seed, ctx->d_mods.as<ModTable>(), ctx->have_lists ? ctx->sh_lists.as<uint4>() : nullptr, st);
        if (ctx->P.score_method == 2) k += launch_predict(ctx->sh_rows.as<uint8_t>(), out_rows, ctx->d_mods.as<ModTable>(), ctx->P.itol, ctx->sh_pred.as<int32_t>(), st);
        cnt.other_kernel_launches += (uint64_t)k;
    };
    if (timed) { Timer t(ctx, &ctx->cnt.ms_shuffle); launch(); } else launch();
    CU(cudaGetLastError());
    return PQ_OK;
}

// Step-4 scoring of the still-valid PSMs of one PSM table (or part of one): one job per valid PSM (StatisticScoreWorker), ordered by
// form so consecutive CTAs share shuffle rows in L2; K1 over the shuffle rows; p-values into the table.  Nothing is synchronised.
struct Step4Work { DevBuf *valid_flag, *list, *order, *keys, *keys2, *jobs, *results, *hits, *hit_count, *tmp; uint32_t* d_cnt; uint32_t* job_counter; };
int32_t step4_score(pq_ctx* ctx, cudaStream_t st, DevPsm* psms, uint32_t n_psm, uint32_t n_valid, const Step4Work& W, bool timed) {
    if (!n_valid) return PQ_OK;
    CU(W.list->reserve((size_t)n_valid * 4 + 64)); CU(W.order->reserve((size_t)n_valid * 4 + 64));
    CU(W.keys->reserve((size_t)n_valid * 4 + 64)); CU(W.keys2->reserve((size_t)n_valid * 4 + 64));
    int32_t rc = select_iota_on(ctx, st, *W.tmp, W.valid_flag->as<uint8_t>(), W.list->as<uint32_t>(), W.d_cnt, (int)n_psm);
    if (rc) return rc;
    int k = launch_psm_form_keys(psms, W.list->as<uint32_t>(), n_valid, W.keys->as<uint32_t>(), st);
    rc = sort_pairs_on(ctx, st, *W.tmp, W.keys->as<uint32_t>(), W.keys2->as<uint32_t>(), W.list->as<uint32_t>(), W.order->as<uint32_t>(), (int)n_valid, 32);
    if (rc) return rc;
    CU(W.jobs->reserve((size_t)n_valid * sizeof(Job)));
    k += launch_step4_jobs(psms, W.order->as<uint32_t>(), n_valid, ctx->sh_form_slot.as<uint32_t>(), ctx->sh_forms.as<ShuffleForm>(),
                           ctx->sh_kcnt.as<uint32_t>(), W.jobs->as<Job>(), st);
    ctx->cnt.other_kernel_launches += (uint64_t)k + 4;
    ScoreIO io{st, W.jobs->as<Job>(), W.results, W.hits, W.hit_count, W.job_counter, timed};
    rc = run_score_io(ctx, io, kJobShuffles, n_valid, ctx->sh_pred.as<int32_t>(), ctx->sh_rows.as<uint8_t>(), nullptr, 1, 0);
    if (rc) return rc;
    k = launch_job_counters(W.jobs->as<Job>(), W.results->as<JobResult>(), n_valid, ctx->s_npeaks.as<uint32_t>(), 2, ctx->d_counters.as<DevCounters>(), st);
    k += launch_step4_apply(W.jobs->as<Job>(), W.results->as<JobResult>(), W.order->as<uint32_t>(), n_valid, psms, st);
    ctx->cnt.other_kernel_launches += (uint64_t)k;
    CU(cudaGetLastError());
    return PQ_OK;
}

// Steps 2-4 while the spectra are still arriving (pq_params.search_at_load).  The resident spectra are cut into parts in arrival
// order; as soon as K0 has prepared a part, its Step 2 (targets), Step 3 (competitors from a reference table built for the mass
// range of ALL resident spectra, a superset of the final one) and Step 4 (shuffles generated up front for every target form) run
// on a stream of their own, beside the gather and K0 of the later parts.  The parts' PSM tables are then concatenated, the
// reference table cut down to the mass window of the matched spectra (GeneratePeptideformWorker.java:40-52) and the row numbers
// shifted accordingly: the context ends up exactly as after pq_step2 .. pq_step4 called one by one.
int32_t search_at_load(pq_ctx* ctx, const LoadPlan& plan) {
    Marks mk("search_at_load");
    cudaStream_t sa = ctx->st_aux, sp_ = ctx->st_part;
    PeptideTable& T = ctx->targets; PeptideTable& R = ctx->reference; PeptideTable& S = ctx->ref_super;
    const int64_t nv = ctx->nS;
    SpectraView sp = ctx->view();
    // ---- the Step-2 jobs by part: queued first (they need nothing of the reference), read back once the reference table is on its way ----
    const uint32_t n_jobs = plan.n_jobs;
    uint32_t chunks_per_part = 6;      // measured at C2 size: 2 -> 48.4 ms end to end, 4 -> 44.4, 6 -> 43.7, 8 -> 43.6 (every part costs ~40 small launches and five host round trips)
    if (const char* e = getenv("PQ_CHUNKS_PER_PART")) { int v = atoi(e); if (v > 0) chunks_per_part = (uint32_t)v; }
    // the first part is shorter: the reference table and the jobs are ready well before six chunks have arrived (4 -> 41.4 ms, 6 -> 42.4 ms when every part has that size)
    uint32_t first_chunks = std::min<uint32_t>(4, chunks_per_part);
    if (const char* e = getenv("PQ_FIRST_PART_CHUNKS")) { int v = atoi(e); if (v > 0) first_chunks = std::min<uint32_t>((uint32_t)v, chunks_per_part); }
    const uint32_t shift_chunks = chunks_per_part - first_chunks;
    const uint32_t part_spectra = plan.chunk_spectra * chunks_per_part, part_shift = plan.chunk_spectra * shift_chunks;
    const uint32_t n_parts = (plan.n_res + part_shift + part_spectra - 1) / part_spectra;
    if (ctx->parts.size() < n_parts) ctx->parts.resize(n_parts);
    std::vector<uint32_t> first(n_parts + 1, 0); std::vector<unsigned long long> pcand(n_parts, 0);
    int k_parts = 0; uint32_t* d_first = nullptr; unsigned long long* d_pc = nullptr;
    if (n_jobs) {
        CU(ctx->up_rank.reserve((size_t)nv * 4 + 64)); CU(ctx->part_first.reserve(((size_t)n_parts + 2) * 16));
        CU(ctx->k_key_a.reserve((size_t)n_jobs * 8)); CU(ctx->k_key_b.reserve((size_t)n_jobs * 8)); CU(ctx->k_idx_a.reserve((size_t)n_jobs * 4)); CU(ctx->k_idx_b.reserve((size_t)n_jobs * 4));
        CU(ctx->jobs2.reserve((size_t)n_jobs * sizeof(Job)));
        k_parts = launch_scatter_rank(ctx->gather_list.as<int32_t>(), plan.n_res, ctx->up_rank.as<uint32_t>(), sp_);
        k_parts += launch_fix_job_blobs(ctx->jobs.as<Job>(), n_jobs, ctx->prep_id.as<int32_t>(), sp_);
        k_parts += launch_part_job_keys(ctx->jobs.as<Job>(), n_jobs, ctx->up_rank.as<uint32_t>(), part_spectra, part_shift, ctx->k_key_a.as<uint64_t>(), ctx->k_idx_a.as<uint32_t>(), sp_);
        int32_t rc = sort_pairs_on(ctx, sp_, ctx->parts[0].tmp, ctx->k_key_a.as<uint64_t>(), ctx->k_key_b.as<uint64_t>(), ctx->k_idx_a.as<uint32_t>(), ctx->k_idx_b.as<uint32_t>(), (int)n_jobs, 64);
        if (rc) return rc;
        k_parts += launch_gather_jobs(ctx->jobs.as<Job>(), ctx->k_idx_b.as<uint32_t>(), n_jobs, ctx->jobs2.as<Job>(), sp_);
        std::swap(ctx->jobs, ctx->jobs2);
        d_first = ctx->part_first.as<uint32_t>();
        d_pc = reinterpret_cast<unsigned long long*>(ctx->part_first.as<uint8_t>() + ((size_t)n_parts + 2) * 8);
        CU(cudaMemsetAsync(d_pc, 0, (size_t)n_parts * 8, sp_));
        k_parts += launch_part_bounds(ctx->k_key_b.as<uint64_t>(), n_jobs, n_parts, d_first, sp_);
        k_parts += launch_part_cand(ctx->jobs.as<Job>(), n_jobs, ctx->k_key_b.as<uint64_t>(), d_pc, sp_);
    }
    // ---- reference: digest (if not done), superset expansion over the resident mass range ----
    if (!ctx->ref_digested) {
        std::string err;
        if (!upload_target_keys_for_reference(sa, ctx->targets.seqs, R.dev, err)) return fail(ctx, PQ_ERR_CUDA, "reference digest: " + err);
        int32_t rc = digest_staged_reference(ctx, sa); if (rc) return rc;
    }
    mk.mark("digest");
    {
        int32_t s_lo = 0, s_hi = 0; double m_lo = 0, m_hi = 0;
        CU(cudaMemcpyAsync(&s_lo, ctx->prep_list.as<int32_t>(), 4, cudaMemcpyDeviceToHost, sa));
        CU(cudaMemcpyAsync(&s_hi, ctx->prep_list.as<int32_t>() + (plan.n_res - 1), 4, cudaMemcpyDeviceToHost, sa));
        CU(cudaStreamSynchronize(sa));
        CU(cudaMemcpyAsync(&m_lo, ctx->s_mass.as<double>() + s_lo, 8, cudaMemcpyDeviceToHost, sa));
        CU(cudaMemcpyAsync(&m_hi, ctx->s_mass.as<double>() + s_hi, 8, cudaMemcpyDeviceToHost, sa));
        CU(cudaStreamSynchronize(sa));
        std::string err; uint64_t launches = 0;
        R.dev.rows = &S.rows; R.dev.mass = &S.mass;
        std::swap(R.dev.seq_id, ctx->super_seq);
        const bool ok = expand_reference_on_device(sa, ctx->P, ctx->codes, m_lo - 250.0 - 5.0, m_hi + 250.0 + 5.0, R.dev, &launches, err);
        std::swap(R.dev.seq_id, ctx->super_seq);
        R.dev.rows = &R.rows; R.dev.mass = &R.mass;
        if (!ok) return fail(ctx, PQ_ERR_CUDA, "reference build: " + err);
        ctx->cnt.other_kernel_launches += launches;
        ctx->super_rows = R.dev.n_forms; R.dev.n_forms = 0;
        S.have_lad = false;
        { int32_t rc = ensure_ladders(ctx, S, S.rows.as<uint8_t>(), ctx->super_rows, sa, false); if (rc) return rc; }
        if (ctx->P.score_method == 2 && ctx->super_rows) {
            CU(S.pred.reserve((size_t)ctx->super_rows * 8));
            ctx->cnt.other_kernel_launches += (uint64_t)launch_predict(S.rows.as<uint8_t>(), ctx->super_rows, ctx->d_mods.as<ModTable>(), ctx->P.itol, S.pred.as<int32_t>(), sa);
        }
    }
    mk.mark("superset reference");
    // ---- shuffles for every target form ----
    if (!(ctx->shuffles_all && ctx->shuffles_epoch == ctx->targets_epoch && ctx->have_shuffles)) {
        if (ctx->P.n_random > 20000) return fail(ctx, PQ_ERR_UNSUPPORTED, "more than 20000 shuffles per peptide");
        ShuffleTotals tot; memset(&tot, 0, sizeof tot);
        int32_t rc = generate_shuffles(ctx, sa, true, tot, false);
        if (rc) return rc;
        ctx->shuffles_all = true; ctx->shuffles_epoch = ctx->targets_epoch;
    }
    CU(cudaStreamSynchronize(sa));
    mk.mark("shuffles for all forms");
    if (n_jobs) {
        CU(cudaMemcpyAsync(first.data(), d_first, ((size_t)n_parts + 1) * 4, cudaMemcpyDeviceToHost, sp_));
        CU(cudaMemcpyAsync(pcand.data(), d_pc, (size_t)n_parts * 8, cudaMemcpyDeviceToHost, sp_));
        CU(cudaStreamSynchronize(sp_));
        ctx->cnt.other_kernel_launches += (uint64_t)k_parts + 3;
    }
    mk.mark("jobs by part");
    WindowConfig W = window_config(ctx, false);
    const TargetAux A = ctx->aux();
    uint64_t N = 0, NS = 0, NH3 = 0;
    for (uint32_t p = 0; p < n_parts; ++p) {
        pq_ctx::PartBufs& B = ctx->parts[p];
        B.n_psm = 0; B.n_seg = 0; B.nh3 = 0;
        // the parts alternate between two streams: while this thread waits for the small results of part p (hit count, segment count,
        // candidate totals), the Step-4 launch of part p - 1 is still running on the other stream
        cudaStream_t sq = (p & 1u) && !getenv("PQ_ONE_PART_STREAM") ? ctx->st_part2 : sp_;
        const uint32_t j0 = first[p], nj = first[p + 1] - first[p];
        // the part is complete when K0 has finished its last chunk
        const size_t last_chunk = std::min<size_t>((size_t)(p + 1) * chunks_per_part - shift_chunks, plan.n_chunks) - 1;
        if (plan.issued) {                       // a second thread records these events: wait (on the host) until this one has been recorded
            while (plan.issued->load(std::memory_order_acquire) <= last_chunk) {
                if (plan.failed && plan.failed->load(std::memory_order_acquire)) return fail(ctx, PQ_ERR_CUDA, "the upload of the spectra failed");
                std::this_thread::yield();
            }
        }
        CU(cudaStreamWaitEvent(sq, (*plan.k0_done)[last_chunk], 0));
        if (!nj) continue;
        if (pcand[p] > 0xfffffff0ull) return fail(ctx, PQ_ERR_UNSUPPORTED, "too many Step-2 candidates for one launch");
        CU(B.scal.reserve(256));
        CU(cudaMemsetAsync(B.scal.p, 0, 256, sq));
        uint32_t* scal = B.scal.as<uint32_t>();
        // ---- Step 2 ----
        ScoreIO io2{sq, ctx->jobs.as<Job>() + j0, &B.results, &B.hits, &B.hit_count, scal + 12, false};
        int32_t rc = run_score_io(ctx, io2, kJobTargets, nj, T.pred.as<int32_t>(), T.rows.as<uint8_t>(), T.mass.as<double>(), (uint32_t)std::max<unsigned long long>(pcand[p], 1), 1);
        if (rc) return rc;
        uint32_t nh = 0;
        CU(cudaMemcpyAsync(&nh, B.hit_count.p, 4, cudaMemcpyDeviceToHost, sq));
        CU(cudaStreamSynchronize(sq));
        ctx->cnt.other_kernel_launches += (uint64_t)launch_job_counters(ctx->jobs.as<Job>() + j0, B.results.as<JobResult>(), nj, sp.npeaks, 0, ctx->d_counters.as<DevCounters>(), sq);
        if (!nh) continue;
        CU(B.psm.reserve((size_t)nh * sizeof(DevPsm))); CU(B.seg_start.reserve(((size_t)nh + 2) * 4));
        {
            std::string err; int launches = 0;
            if (!psm_table_from_hits(B.hits.as<Hit>(), nh, sp, T.mass.as<double>(), T.rows.as<uint8_t>(), A, B.key_a, B.key_b, B.idx_a, B.idx_b, B.head, B.tmp,
                                     B.psm.as<DevPsm>(), B.seg_start.as<uint32_t>(), scal, nullptr, sq, &launches, err))
                return fail(ctx, PQ_ERR_CUDA, "psm table: " + err);
            ctx->cnt.other_kernel_launches += (uint64_t)launches;
        }
        uint32_t nseg = 0;
        CU(cudaMemcpyAsync(&nseg, scal, 4, cudaMemcpyDeviceToHost, sq));
        CU(cudaMemcpyAsync(&B.s_first, &B.psm.as<DevPsm>()[0].spectrum_sorted, 4, cudaMemcpyDeviceToHost, sq));
        CU(cudaMemcpyAsync(&B.s_last, &B.psm.as<DevPsm>()[nh - 1].spectrum_sorted, 4, cudaMemcpyDeviceToHost, sq));
        CU(cudaStreamSynchronize(sq));
        B.n_psm = nh; B.n_seg = nseg;
        // ---- Step 3 (superset table) ----
        CU(B.agg.reserve((size_t)nseg * sizeof(SegAgg)));
        if (ctx->super_rows == 0) {
            CU(cudaMemsetAsync(B.agg.p, 0, (size_t)nseg * sizeof(SegAgg), sq));
        } else {
            const uint32_t nj3 = nseg * (uint32_t)ctx->P.n_isotope;
            CU(B.jobs.reserve((size_t)nj3 * sizeof(Job)));
            unsigned long long* d_total = reinterpret_cast<unsigned long long*>(scal + 16);
            ctx->cnt.other_kernel_launches += (uint64_t)launch_step3_jobs(B.psm.as<DevPsm>(), B.seg_start.as<uint32_t>(), nseg, S.mass.as<double>(), ctx->super_rows, W,
                                                                          ctx->s_mass.as<double>(), B.jobs.as<Job>(), B.agg.as<SegAgg>(), d_total, sq);
            unsigned long long cand3 = 0;
            CU(cudaMemcpyAsync(&cand3, d_total, 8, cudaMemcpyDeviceToHost, sq));
            CU(cudaStreamSynchronize(sq));
            if (cand3 > 0xfffffff0ull) return fail(ctx, PQ_ERR_UNSUPPORTED, "too many Step-3 candidates for one launch");
            ScoreIO io3{sq, B.jobs.as<Job>(), &B.results, &B.hits, &B.hit_count, scal + 12, false};
            rc = run_score_io(ctx, io3, kJobCompetitors, nj3, S.pred.as<int32_t>(), S.rows.as<uint8_t>(), S.mass.as<double>(), (uint32_t)std::max<unsigned long long>(cand3, 1), 1, &S);
            if (rc) return rc;
            int k = launch_job_counters(B.jobs.as<Job>(), B.results.as<JobResult>(), nj3, sp.npeaks, 1, ctx->d_counters.as<DevCounters>(), sq);
            k += launch_step3_aggregate(B.results.as<JobResult>(), nseg, ctx->P.n_isotope, B.seg_start.as<uint32_t>(), B.psm.as<DevPsm>(), B.agg.as<SegAgg>(), sq);
            ctx->cnt.other_kernel_launches += (uint64_t)k;
            uint32_t nh3 = 0;
            CU(cudaMemcpyAsync(&nh3, B.hit_count.p, 4, cudaMemcpyDeviceToHost, sq));
            CU(cudaStreamSynchronize(sq));
            if (nh3) { CU(B.hits3.reserve((size_t)nh3 * sizeof(Hit))); CU(cudaMemcpyAsync(B.hits3.p, B.hits.p, (size_t)nh3 * sizeof(Hit), cudaMemcpyDeviceToDevice, sq)); }
            B.nh3 = nh3;
        }
        // ---- Step 4 ----
        CU(B.valid_flag.reserve((size_t)nh + 64));
        ctx->cnt.other_kernel_launches += (uint64_t)launch_valid_flags(B.psm.as<DevPsm>(), nh, B.valid_flag.as<uint8_t>(), scal + 5, sq);
        uint32_t n_valid = 0;
        CU(cudaMemcpyAsync(&n_valid, scal + 5, 4, cudaMemcpyDeviceToHost, sq));
        CU(cudaStreamSynchronize(sq));
        Step4Work W4{&B.valid_flag, &B.list, &B.order, &B.keys, &B.keys2, &B.jobs, &B.results, &B.hits, &B.hit_count, &B.tmp, scal + 4, scal + 12};
        if (ctx->shuf_pending) CU(cudaStreamWaitEvent(sq, ctx->shuf_ev, 0));
        rc = step4_score(ctx, sq, B.psm.as<DevPsm>(), nh, n_valid, W4, false);
        if (rc) return rc;
        N += nh; NS += nseg; NH3 += B.nh3;
    }
    CU(cudaStreamSynchronize(ctx->st_part2));
    if (ctx->shuf_pending) { CU(cudaStreamSynchronize(ctx->st_shuf)); ctx->shuf_pending = false; }
    mk.mark("parts queued");
    // ---- merge ----
    ctx->n_psm = 0; ctx->n_seg = 0; ctx->nh3 = 0; ctx->comp3_ready = false; ctx->comp3_n = 0;
    R.seqs.clear(); R.forms.clear(); R.have_mask = false; R.have_lad = false; R.device_built = true; R.materialized = false; R.h_mass.clear(); R.n_rows = 0;
    if (N) {
        CU(ctx->d_psm.reserve((size_t)N * sizeof(DevPsm))); CU(ctx->d_seg_start.reserve(((size_t)NS + 2) * 4)); CU(ctx->d_out_perm.reserve((size_t)N * 4));
        CU(ctx->d_agg.reserve((size_t)NS * sizeof(SegAgg))); CU(ctx->hits3.reserve(std::max<size_t>((size_t)NH3, 1) * sizeof(Hit)));
        // mass window of the matched spectra -> rows [a, b) of the superset are the reference table
        int32_t s_first = 0x7fffffff, s_last = -1;
        for (uint32_t p = 0; p < n_parts; ++p) if (ctx->parts[p].n_psm) { s_first = std::min(s_first, ctx->parts[p].s_first); s_last = std::max(s_last, ctx->parts[p].s_last); }
        CU(cudaMemcpyAsync(&ctx->psm_mass_lo, ctx->s_mass.as<double>() + s_first, 8, cudaMemcpyDeviceToHost, sp_));
        CU(cudaMemcpyAsync(&ctx->psm_mass_hi, ctx->s_mass.as<double>() + s_last, 8, cudaMemcpyDeviceToHost, sp_));
        CU(cudaStreamSynchronize(sp_));
        uint32_t ab[2] = {0, 0};
        if (ctx->super_rows) {
            uint32_t* d_ab = ctx->part_first.as<uint32_t>();
            launch_window_rows(S.mass.as<double>(), ctx->super_rows, ctx->psm_mass_lo - 250.0 - 5.0, ctx->psm_mass_hi + 250.0 + 5.0, d_ab, sp_);
            CU(cudaMemcpyAsync(ab, d_ab, 8, cudaMemcpyDeviceToHost, sp_));
            CU(cudaStreamSynchronize(sp_));
        }
        const uint32_t a = ab[0], nr = ab[1] - ab[0];
        if (nr) {
            CU(R.rows.reserve((size_t)nr * kRowBytes)); CU(R.mass.reserve((size_t)nr * 8)); CU(R.dev.seq_id.reserve((size_t)nr * 4));
            CU(cudaMemcpyAsync(R.rows.p, S.rows.as<uint8_t>() + (size_t)a * kRowBytes, (size_t)nr * kRowBytes, cudaMemcpyDeviceToDevice, sp_));
            CU(cudaMemcpyAsync(R.mass.p, S.mass.as<double>() + a, (size_t)nr * 8, cudaMemcpyDeviceToDevice, sp_));
            CU(cudaMemcpyAsync(R.dev.seq_id.p, ctx->super_seq.as<uint32_t>() + a, (size_t)nr * 4, cudaMemcpyDeviceToDevice, sp_));
            if (ctx->P.score_method == 2) { CU(R.pred.reserve((size_t)nr * 8)); CU(cudaMemcpyAsync(R.pred.p, S.pred.as<int32_t>() + 2 * (size_t)a, (size_t)nr * 8, cudaMemcpyDeviceToDevice, sp_)); }
        }
        R.n_rows = nr; R.dev.n_forms = nr;
        uint32_t psm_off = 0, seg_off = 0, h3_off = 0; int k = 0;
        for (uint32_t p = 0; p < n_parts; ++p) {
            pq_ctx::PartBufs& B = ctx->parts[p];
            if (!B.n_psm) continue;
            const bool last = psm_off + B.n_psm == N;
            k += launch_merge_psms(B.psm.as<DevPsm>(), B.n_psm, seg_off, ctx->d_psm.as<DevPsm>() + psm_off, sp_);
            k += launch_merge_segs(B.seg_start.as<uint32_t>(), B.agg.as<SegAgg>(), B.n_seg, psm_off, a, ctx->d_seg_start.as<uint32_t>() + seg_off, ctx->d_agg.as<SegAgg>() + seg_off,
                                   (uint32_t)N, last ? 1 : 0, sp_);
            k += launch_merge_hits(B.hits3.as<Hit>(), B.nh3, a, ctx->hits3.as<Hit>() + h3_off, sp_);
            psm_off += B.n_psm; seg_off += B.n_seg; h3_off += B.nh3;
        }
        std::string err;
        if (!build_out_perm(ctx->d_psm.as<DevPsm>(), (uint32_t)N, ctx->k_key_a, ctx->k_key_b, ctx->k_idx_a, ctx->scan_tmp, ctx->d_out_perm.as<uint32_t>(), sp_, err))
            return fail(ctx, PQ_ERR_CUDA, "psm table: " + err);
        ctx->cnt.other_kernel_launches += (uint64_t)k + 6;
        ctx->n_psm = (uint32_t)N; ctx->n_seg = (uint32_t)NS; ctx->nh3 = (uint32_t)NH3;
    }
    CU(cudaStreamSynchronize(sp_));
    ctx->cnt.psms = ctx->n_psm;
    ctx->cnt.reference_sequences = R.dev.n_unique; ctx->cnt.reference_forms = R.n_rows;
    ctx->have_reference = true; ctx->have_shuffles = true;
    ctx->step_done = 4;
    ctx->pending_step[2] = ctx->pending_step[3] = ctx->pending_step[4] = true; ctx->pending_ref = true;
    mk.mark("merge");
    return PQ_OK;
}

}  // namespace

extern "C" {

int32_t pq_step4_shuffles(pq_ctx* ctx) {
    if (!ctx) return PQ_ERR_BAD_ARG;
    if (ctx->step_done < 2) return fail(ctx, PQ_ERR_STATE, "pq_step2_match_targets first");
    cudaSetDevice(ctx->P.device);
    if (ctx->pending_step[4]) { ctx->pending_step[4] = false; ctx->step_done = 4; return PQ_OK; }     // searched while the spectra were loaded
    cudaStream_t st = ctx->st;
    Marks mk("step4");
    const uint32_t n_psm = ctx->n_psm;
    if (n_psm == 0) { ctx->step_done = 4; return PQ_OK; }
    if (ctx->P.n_random > 20000) return fail(ctx, PQ_ERR_UNSUPPORTED, "more than 20000 shuffles per peptide");
    const TargetAux A = ctx->aux();
    const size_t nseq = A.n_seq, nform = A.n_forms;
    const bool pre = ctx->shuffles_all && ctx->shuffles_epoch == ctx->targets_epoch && ctx->have_shuffles;      // generated for every form already
    CU(ctx->sh_valid_flag.reserve((size_t)n_psm + 64)); CU(ctx->sh_totals.reserve(64));
    ShuffleTotals tot; memset(&tot, 0, sizeof tot);
    uint32_t n_valid = 0;
    if (!pre) {
        // StatisticScoreCalc.run (:22-41): sequences with at least one still-valid PSM get shuffles
        ctx->have_shuffles = false; ctx->shuffles_all = false;
        CU(ctx->sh_seq_flag.reserve(nseq + 64)); CU(ctx->sh_form_flag.reserve(nform * 4 + 64));
        CU(ctx->sh_plans.reserve(nseq * sizeof(ShuffleSeq))); CU(ctx->sh_seqs.reserve(nseq * sizeof(ShuffleSeq))); CU(ctx->sh_forms.reserve(nform * sizeof(ShuffleForm)));
        CU(ctx->sh_seq_slot.reserve(nseq * 4 + 64)); CU(ctx->sh_form_slot.reserve(nform * 4 + 64)); CU(ctx->sh_row_base.reserve(nform * 4 + 64));
        CU(ctx->sh_kcnt.reserve(nseq * 4 + 64));
        CU(cudaMemsetAsync(ctx->sh_seq_flag.p, 0, nseq, st)); CU(cudaMemsetAsync(ctx->sh_form_flag.p, 0, nform * 4, st));
        CU(cudaMemsetAsync(ctx->sh_totals.p, 0, sizeof(ShuffleTotals), st));
        {
            Timer t(ctx, &ctx->cnt.ms_shuffle);
            ctx->cnt.other_kernel_launches += (uint64_t)launch_flag_valid(ctx->d_psm.as<DevPsm>(), n_psm, A.seq_of_form, ctx->sh_seq_flag.as<uint8_t>(), ctx->sh_form_flag.as<uint32_t>(),
                                                                          ctx->sh_valid_flag.as<uint8_t>(), ctx->sh_totals.as<ShuffleTotals>(), st);
        }
        int32_t rc = generate_shuffles(ctx, st, false, tot, true);
        if (rc) return rc;
        n_valid = tot.n_valid;
        mk.mark("shuffle plans + layout + K2 (launch)");
        if (tot.n_seqs == 0 || n_valid == 0) { ctx->step_done = 4; return PQ_OK; }
    } else {
        uint32_t* d_nv = ctx->d_scalars.as<uint32_t>() + 5;
        CU(cudaMemsetAsync(d_nv, 0, 4, st));
        ctx->cnt.other_kernel_launches += (uint64_t)launch_valid_flags(ctx->d_psm.as<DevPsm>(), n_psm, ctx->sh_valid_flag.as<uint8_t>(), d_nv, st);
        CU(cudaMemcpyAsync(&n_valid, d_nv, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (n_valid == 0) { ctx->step_done = 4; return PQ_OK; }
    }
    Step4Work W{&ctx->sh_valid_flag, &ctx->sh_list, &ctx->sh_order, &ctx->sh_keys, &ctx->sh_keys2, &ctx->jobs, &ctx->results, &ctx->hits, &ctx->hit_count, &ctx->scan_tmp,
                ctx->d_scalars.as<uint32_t>() + 4, ctx->d_scalars.as<uint32_t>() + 12};
    int32_t rc = step4_score(ctx, st, ctx->d_psm.as<DevPsm>(), n_psm, n_valid, W, true);
    if (rc) return rc;
    mk.mark("jobs + score (launch)");
    CU(cudaStreamSynchronize(st));            // synchronous at return (header contract): a kernel fault surfaces here, not in a later call
    mk.mark("wait");
    ctx->step_done = 4;
    return PQ_OK;
}

int32_t pq_get_shuffles(pq_ctx* ctx, int32_t target_seq, char* out, int64_t capacity, int32_t* n_shuffles) {
    if (!ctx || !n_shuffles) return PQ_ERR_BAD_ARG;
    *n_shuffles = 0;
    if (!ctx->have_shuffles || target_seq < 0 || (size_t)target_seq >= ctx->targets.seqs.size()) return PQ_OK;
    cudaSetDevice(ctx->P.device);
    uint8_t flag = 0; uint32_t slot = 0;
    CU(cudaMemcpy(&flag, ctx->sh_seq_flag.as<uint8_t>() + target_seq, 1, cudaMemcpyDeviceToHost));
    if (!flag) return PQ_OK;
    CU(cudaMemcpy(&slot, ctx->sh_seq_slot.as<uint32_t>() + target_seq, 4, cudaMemcpyDeviceToHost));
    ShuffleSeq S; uint32_t nk = 0;
    CU(cudaMemcpy(&S, ctx->sh_seqs.as<ShuffleSeq>() + slot, sizeof S, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&nk, ctx->sh_kcnt.as<uint32_t>() + slot, 4, cudaMemcpyDeviceToHost));
    *n_shuffles = (int32_t)nk;
    if (!out) return PQ_OK;
    int64_t L1 = (int64_t)S.len + 1;
    if (capacity < (int64_t)nk * L1) return fail(ctx, PQ_ERR_CAPACITY, "shuffle buffer too small");
    std::vector<uint8_t> raw((size_t)S.num * kRowBytes); std::vector<uint32_t> kidx(S.num);
    if (S.num) {
        CU(cudaMemcpy(raw.data(), ctx->sh_raw.as<uint8_t>() + (size_t)S.out_base * kRowBytes, raw.size(), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(kidx.data(), ctx->sh_kidx.as<uint32_t>() + S.out_base, (size_t)S.num * 4, cudaMemcpyDeviceToHost));
    }
    for (uint32_t k = 0; k < nk; ++k) {
        const uint8_t* r = raw.data() + (size_t)kidx[k] * kRowBytes;
        for (uint32_t i = 0; i < S.len; ++i) out[k * L1 + i] = (char)('A' + r[kRowHeader + i]);
        out[k * L1 + S.len] = 0;
    }
    return PQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
int32_t pq_rank_and_validate(pq_ctx* ctx) {
    if (!ctx) return PQ_ERR_BAD_ARG;
    cudaSetDevice(ctx->P.device);
    // rank inside each spectrum segment (pg/RankPSM.java:107-149) + CParameter.passed_psm_validation, on the device
    ctx->cnt.other_kernel_launches += (uint64_t)launch_rank(ctx->d_psm.as<DevPsm>(), ctx->n_psm, ctx->d_seg_start.as<uint32_t>(), ctx->aux(), ctx->st);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->st));
    return PQ_OK;
}

int32_t pq_get_psms(pq_ctx* ctx, pq_psm* out, int64_t capacity, int64_t* n_psms) {
    if (!ctx || !n_psms) return PQ_ERR_BAD_ARG;
    *n_psms = (int64_t)ctx->n_psm;
    if (!out) return PQ_OK;
    if (capacity < *n_psms) return fail(ctx, PQ_ERR_CAPACITY, "psm buffer too small");
    if (!ctx->n_psm) return PQ_OK;
    cudaSetDevice(ctx->P.device);
    const size_t bytes = (size_t)ctx->n_psm * sizeof(pq_psm);
    CU(ctx->d_export.reserve(bytes));
    ctx->cnt.other_kernel_launches += (uint64_t)launch_export_psms(ctx->d_psm.as<DevPsm>(), ctx->d_out_perm.as<uint32_t>(), ctx->n_psm, ctx->d_export.as<pq_psm>(), ctx->st);
    int32_t rc = copy_to_caller(ctx, out, ctx->d_export.p, bytes);
    if (rc) return rc;
    ctx->cnt.d2h_bytes += bytes;
    return PQ_OK;
}

int32_t pq_get_delta_scores(pq_ctx* ctx, double* ref_delta, double* mod_delta, int64_t capacity, int64_t* n_psms) {
    if (!ctx || !n_psms) return PQ_ERR_BAD_ARG;
    *n_psms = (int64_t)ctx->n_psm;
    if (!ref_delta || !mod_delta) return PQ_OK;
    if (capacity < *n_psms) return fail(ctx, PQ_ERR_CAPACITY, "delta score buffers too small");
    if (!ctx->n_psm) return PQ_OK;
    cudaSetDevice(ctx->P.device);
    cudaStream_t st = ctx->st;
    const size_t n = ctx->n_psm;
    // best ptm.txt row per spectrum (rows of Step 5 live on the host, sorted by spectrum): a short sorted key list for the kernel
    std::vector<int32_t> mk; std::vector<double> mb;
    if (ctx->step_done >= 5)
        for (const pq_competitor& r : ctx->comp5) {
            if (mk.empty() || mk.back() != r.spectrum) { mk.push_back(r.spectrum); mb.push_back(r.score); }
            else if (mb.back() < r.score) mb.back() = r.score;
        }
    for (size_t i = 1; i < mk.size(); ++i) if (mk[i - 1] >= mk[i]) return fail(ctx, PQ_ERR_STATE, "ptm rows are not ordered by spectrum");
    CU(ctx->d_delta.reserve(2 * n * 8)); CU(ctx->d_modkey.reserve(mk.size() * 4 + 4)); CU(ctx->d_modbest.reserve(mb.size() * 8 + 8));
    if (!mk.empty()) {
        CU(cudaMemcpyAsync(ctx->d_modkey.p, mk.data(), mk.size() * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(ctx->d_modbest.p, mb.data(), mb.size() * 8, cudaMemcpyHostToDevice, st));
        ctx->cnt.h2d_bytes += mk.size() * 12;
    }
    const SegAgg* agg = (ctx->step_done >= 3 && ctx->n_seg) ? ctx->d_agg.as<SegAgg>() : nullptr;
    ctx->cnt.other_kernel_launches += (uint64_t)launch_delta_scores(ctx->d_psm.as<DevPsm>(), ctx->d_out_perm.as<uint32_t>(), ctx->n_psm, agg, ctx->d_modkey.as<int32_t>(),
                                                                    ctx->d_modbest.as<double>(), (uint32_t)mk.size(), ctx->d_delta.as<double>(), ctx->d_delta.as<double>() + n, st);
    CU(cudaStreamSynchronize(st));                      // mk / mb leave scope
    int32_t rc = copy_to_caller(ctx, ref_delta, ctx->d_delta.p, n * 8);
    if (rc) return rc;
    rc = copy_to_caller(ctx, mod_delta, ctx->d_delta.as<double>() + n, n * 8);
    if (rc) return rc;
    ctx->cnt.d2h_bytes += 2 * n * 8;
    return PQ_OK;
}

int32_t pq_find_in_proteome(pq_ctx* ctx, int32_t n_seq, const uint32_t* seq_off, const char* res, int32_t n_prot, const uint64_t* prot_off,
                            const char* prot_res, int32_t* nhit) {
    if (!ctx || n_seq < 0 || n_prot < 0 || (n_seq > 0 && (!seq_off || !res || !nhit)) || (n_prot > 0 && (!prot_off || !prot_res)))
        return fail(ctx, PQ_ERR_BAD_ARG, "null argument");
    cudaSetDevice(ctx->P.device);
    std::string err; int launches = 0;
    if (!find_in_proteome(n_seq, seq_off, res, n_prot, prot_off, prot_res, nhit, ctx->st, &ctx->cnt.h2d_bytes, &ctx->cnt.d2h_bytes, &launches, err))
        return fail(ctx, err.find("target") != std::string::npos ? PQ_ERR_BAD_ARG : PQ_ERR_CUDA, "novelty filter: " + err);
    ctx->cnt.other_kernel_launches += (uint64_t)launches;
    return PQ_OK;
}

int32_t pq_export_psms_device(pq_ctx* ctx, uint64_t* device_address, int64_t* n_psms) {
    if (!ctx || !device_address || !n_psms) return PQ_ERR_BAD_ARG;
    *n_psms = (int64_t)ctx->n_psm; *device_address = 0;
    if (!ctx->n_psm) return PQ_OK;
    cudaSetDevice(ctx->P.device);
    CU(ctx->d_export.reserve((size_t)ctx->n_psm * sizeof(pq_psm)));
    ctx->cnt.other_kernel_launches += (uint64_t)launch_export_psms(ctx->d_psm.as<DevPsm>(), ctx->d_out_perm.as<uint32_t>(), ctx->n_psm, ctx->d_export.as<pq_psm>(), ctx->st);
    CU(cudaStreamSynchronize(ctx->st));       // the caller's collective runs on another stream
    *device_address = (uint64_t)(uintptr_t)ctx->d_export.p;
    return PQ_OK;
}

int32_t pq_get_competitors(pq_ctx* ctx, int32_t step, pq_competitor* out, int64_t capacity, int64_t* n_rows) {
    if (!ctx || !n_rows) return PQ_ERR_BAD_ARG;
    if (step == 5) {
        auto& v = ctx->comp5;
        *n_rows = (int64_t)v.size();
        if (!out) return PQ_OK;
        if (capacity < *n_rows) return fail(ctx, PQ_ERR_CAPACITY, "competitor buffer too small");
        if (!v.empty()) memcpy(out, v.data(), v.size() * sizeof(pq_competitor));
        return PQ_OK;
    }
    // Step 3 detail rows: better-or-equal competitors with score >= 20 (the K1 hit rows) plus the best competitor of each
    // spectrum when its score > 0 (pg/DBSearchWorker.java:71-76, 93-98), sorted by (spectrum, reference row) on the device.
    // Built once per Step 3 (first call), the size query and the fill share it.
    *n_rows = 0;
    if (ctx->step_done < 3 || ctx->n_seg == 0 || !ctx->have_reference || ctx->reference.n() == 0) return PQ_OK;
    cudaSetDevice(ctx->P.device);
    cudaStream_t st = ctx->st;
    if (!ctx->comp3_ready) {
        const size_t cap = (size_t)ctx->nh3 + ctx->n_seg;
        CU(ctx->c_rows.reserve(cap * sizeof(pq_competitor))); CU(ctx->c_rows2.reserve(cap * sizeof(pq_competitor)));
        CU(ctx->c_keys.reserve(cap * 8)); CU(ctx->c_keys2.reserve(cap * 8)); CU(ctx->c_idx.reserve(cap * 4)); CU(ctx->c_idx2.reserve(cap * 4));
        uint32_t* d_cnt = ctx->d_scalars.as<uint32_t>() + 8;
        CU(cudaMemsetAsync(d_cnt, 0, 4, st));
        int k = launch_comp3_rows(ctx->hits3.as<Hit>(), ctx->nh3, ctx->d_agg.as<SegAgg>(), ctx->n_seg, ctx->d_seg_start.as<uint32_t>(), ctx->d_psm.as<DevPsm>(),
                                  ctx->s_caller.as<int32_t>(), ctx->reference.mass.as<double>(), ctx->c_rows.as<pq_competitor>(), ctx->c_keys.as<uint64_t>(),
                                  ctx->c_idx.as<uint32_t>(), d_cnt, st);
        uint32_t extra = 0;
        CU(cudaMemcpyAsync(&extra, d_cnt, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        const uint32_t n = ctx->nh3 + extra;
        if (n) {
            int32_t rc = sort_pairs(ctx, ctx->c_keys.as<uint64_t>(), ctx->c_keys2.as<uint64_t>(), ctx->c_idx.as<uint32_t>(), ctx->c_idx2.as<uint32_t>(), (int)n, 64);
            if (rc) return rc;
            k += launch_gather_comp(ctx->c_rows.as<pq_competitor>(), ctx->c_idx2.as<uint32_t>(), n, ctx->c_rows2.as<pq_competitor>(), st);
        }
        ctx->cnt.other_kernel_launches += (uint64_t)k + (n ? 4 : 0);
        ctx->cnt.d2h_bytes += 4;
        ctx->comp3_n = n; ctx->comp3_ready = true;
    }
    *n_rows = (int64_t)ctx->comp3_n;
    if (!out) return PQ_OK;
    if (capacity < *n_rows) return fail(ctx, PQ_ERR_CAPACITY, "competitor buffer too small");
    if (!ctx->comp3_n) return PQ_OK;
    int32_t rc = copy_to_caller(ctx, out, ctx->c_rows2.p, (size_t)ctx->comp3_n * sizeof(pq_competitor));
    if (rc) return rc;
    ctx->cnt.d2h_bytes += (uint64_t)ctx->comp3_n * sizeof(pq_competitor);
    return PQ_OK;
}

int32_t pq_step5_ptm(pq_ctx* ctx, int32_t n_ptm, const pq_ptm* ptms) {
    if (!ctx || n_ptm < 0 || (n_ptm > 0 && !ptms)) return fail(ctx, PQ_ERR_BAD_ARG, "null argument");
    if (ctx->step_done < 3 || !ctx->have_reference) return fail(ctx, PQ_ERR_STATE, "run Steps 2-4 and pq_build_reference first");
    if (n_ptm > 2048) return fail(ctx, PQ_ERR_UNSUPPORTED, "more than 2048 PTM specificities");
    cudaSetDevice(ctx->P.device);
    cudaStream_t st = ctx->st;
    Marks mk("step5");
    int32_t rc = pq_rank_and_validate(ctx);       // doPTMValidation reads psm_rank.txt (ModificationDB.java:1340-1384)
    if (rc) return rc;
    rc = ensure_host_spectra(ctx);
    if (rc) return rc;
    ctx->comp5.clear();
    PeptideTable& R = ctx->reference; PeptideTable& T = ctx->targets;
    // Step 5 is driven from the host: mirror the device PSM table (table order: PSMs of one spectrum are contiguous)
    {
        std::vector<DevPsm> hp(ctx->n_psm);
        if (ctx->n_psm) { CU(cudaMemcpyAsync(hp.data(), ctx->d_psm.p, (size_t)ctx->n_psm * sizeof(DevPsm), cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st)); }
        ctx->cnt.d2h_bytes += (uint64_t)ctx->n_psm * sizeof(DevPsm);
        ctx->psms.resize(ctx->n_psm);
        for (size_t i = 0; i < hp.size(); ++i) {
            const DevPsm& d = hp[i]; HostPsm& h = ctx->psms[i];
            h.spectrum_sorted = d.spectrum_sorted; h.spectrum = d.spectrum; h.form = d.form; h.len = d.len; h.blob = d.blob; h.score = d.score; h.charge = d.charge;
            h.n_db = d.n_db; h.rank = d.rank; h.n_ptm = d.n_ptm; h.p_value = d.p_value;
        }
    }
    // spectra to validate: PSMs with passed_psm_validation(len, p, n_db, rank); target = MIN score per spectrum (:1369-1383)
    std::vector<size_t> val;
    for (size_t i = 0; i < ctx->psms.size(); ++i) {
        HostPsm& p = ctx->psms[i];
        p.n_ptm = -1;
        int len = p.len;
        bool pv = (len <= 8 && p.p_value <= 0.05) || (len > 8 && p.p_value <= 0.01);
        if (pv && p.n_db == 0 && p.rank == 1) val.push_back(i);
    }
    std::vector<Job> hj; std::vector<size_t> first_psm;
    for (size_t k = 0; k < val.size(); ++k) {           // psms are sorted by spectrum
        const HostPsm& p = ctx->psms[val[k]];
        if (!hj.empty() && hj.back().spectrum == p.spectrum_sorted) { if (p.score < hj.back().ref_score) hj.back().ref_score = p.score; continue; }
        Job J; J.blob = p.blob; J.cand_begin = 0; J.cand_count = 0; J.out = (uint32_t)hj.size(); J.mass = ctx->h_mass[(size_t)p.spectrum_sorted];
        J.ref_score = p.score; J.charge = p.charge; J.iso = 0; J.spectrum = p.spectrum_sorted; J.aux = 0;
        hj.push_back(J);
    }
    std::stable_sort(hj.begin(), hj.end(), [](const Job& a, const Job& b) { return a.spectrum < b.spectrum; });   // mass order: neighbouring CTAs share table rows in L2
    for (size_t j = 0; j < hj.size(); ++j) hj[j].out = (uint32_t)j;
    std::vector<JobResult> hr(hj.size());
    std::vector<Hit> hh;
    if (!hj.empty() && n_ptm > 0 && R.n() > 0) {
        if (!R.have_mask) {
            CU(R.mask.reserve(R.n() * 4));
            ctx->cnt.other_kernel_launches += (uint64_t)launch_row_masks(R.rows.as<uint8_t>(), (uint32_t)R.n(), R.mask.as<uint32_t>(), st);
            R.have_mask = true;
        }
        DevBuf d_ptm;
        CU(d_ptm.reserve((size_t)n_ptm * sizeof(pq_ptm)));
        CU(cudaMemcpyAsync(d_ptm.p, ptms, (size_t)n_ptm * sizeof(pq_ptm), cudaMemcpyHostToDevice, st));
        CU(ctx->jobs.reserve(hj.size() * sizeof(Job)));
        CU(cudaMemcpyAsync(ctx->jobs.p, hj.data(), hj.size() * sizeof(Job), cudaMemcpyHostToDevice, st));
        CU(ctx->results.reserve(hj.size() * sizeof(JobResult)));
        CU(ctx->hit_count.reserve(64));
        ctx->cnt.h2d_bytes += hj.size() * sizeof(Job) + (size_t)n_ptm * sizeof(pq_ptm);
        uint32_t cap = (uint32_t)std::max<size_t>(1u << 20, hj.size() * 64);
        for (int attempt = 0; attempt < 4; ++attempt) {
            CU(ctx->hits.reserve((size_t)cap * sizeof(Hit)));
            CU(cudaMemsetAsync(ctx->hit_count.p, 0, 4, st));
            PtmLaunch L; memset(&L, 0, sizeof L);
            L.jobs = ctx->jobs.as<Job>(); L.n_jobs = (uint32_t)hj.size(); L.blobs = ctx->blobs.as<uint8_t>(); L.blob_off = ctx->blob_off.as<uint64_t>(); L.blob_len = ctx->blob_len.as<uint32_t>();
            L.rows = R.rows.as<uint8_t>(); L.row_mass = R.mass.as<double>(); L.row_mask = R.mask.as<uint32_t>(); L.n_rows = (uint32_t)R.n();
            L.ptms = d_ptm.as<pq_ptm>(); L.n_ptm = n_ptm; L.n_isotope = ctx->P.n_isotope;
            for (int i = 0; i < PQ_MAX_ISOTOPE; ++i) L.isotope[i] = ctx->P.isotope[i];
            L.mods = ctx->d_mods.as<ModTable>(); L.log10_fact = ctx->d_l10.as<double>(); L.ln_fact = ctx->d_lnf.as<double>();
            L.results = ctx->results.as<JobResult>(); L.hits = ctx->hits.as<Hit>(); L.hit_count = ctx->hit_count.as<uint32_t>(); L.hit_capacity = cap;
            L.max_blob_bytes = ctx->max_blob;
            ScoreConfig cfg = score_config(ctx, 1);
            { Timer t(ctx, &ctx->cnt.ms_score, &ctx->cnt.ms_score_step[3]); ctx->cnt.score_kernel_launches += (uint64_t)launch_ptm(L, cfg, st); }
            CU(cudaGetLastError());
            uint32_t nh = 0;
            CU(cudaMemcpyAsync(&nh, ctx->hit_count.p, 4, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(hr.data(), ctx->results.p, hj.size() * sizeof(JobResult), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if (nh > cap) { cap = nh + 1024; continue; }       // rare: more rows than reserved, run again with room
            hh.resize(nh);
            if (nh) { CU(cudaMemcpyAsync(hh.data(), ctx->hits.p, (size_t)nh * sizeof(Hit), cudaMemcpyDeviceToHost, st)); CU(cudaStreamSynchronize(st)); }
            ctx->cnt.d2h_bytes += hj.size() * sizeof(JobResult) + (uint64_t)nh * sizeof(Hit) + 4;
            break;
        }
    }
    mk.mark("score + results d2h");
    std::unordered_map<int32_t, size_t> at;
    for (size_t j = 0; j < hj.size(); ++j) {
        at[hj[j].spectrum] = j;
        ctx->cnt.pairs_step5 += hr[j].n_in_window; ctx->cnt.pairs_step5_bounded += hr[j].pad;
        uint64_t ab = (uint64_t)hr[j].n_in_window * (16ull * ctx->h_npeaks[(size_t)hj[j].spectrum] + 96ull);
        ctx->cnt.score_algorithmic_bytes += ab; ctx->cnt.bytes_score_step[3] += ab;
    }
    for (const Hit& h : hh) ctx->comp5.push_back({ctx->h_caller[(size_t)h.spectrum], (int32_t)h.cand, (int32_t)((uint32_t)h.aux >> 8), (int32_t)(h.aux & 0xFF), h.score, h.mass});
    std::sort(ctx->comp5.begin(), ctx->comp5.end(), [](const pq_competitor& a, const pq_competitor& b) {
        if (a.spectrum != b.spectrum) return a.spectrum < b.spectrum;
        if (a.ref_form != b.ref_form) return a.ref_form < b.ref_form;
        if (a.ptm != b.ptm) return a.ptm < b.ptm;
        return a.ptm_site < b.ptm_site;
    });
    // PSMT.summariseResult (:920-1038): a validated PSM gets n_ptm = #PTM rows scoring > (>= with -hc) its score, others -1
    for (size_t i : val) {
        HostPsm& p = ctx->psms[i];
        auto it = at.find(p.spectrum_sorted);
        if (it == at.end()) { p.n_ptm = 0; continue; }
        if (p.score == hj[it->second].ref_score) p.n_ptm = (int32_t)hr[it->second].n_better;
        else { int n = 0; for (const Hit& h : hh) if (h.spectrum == p.spectrum_sorted && (ctx->P.hc ? h.score >= p.score : h.score > p.score)) ++n; p.n_ptm = n; }
    }
    if (ctx->n_psm) {       // n_ptm back into the device table, then rank/accept again
        std::vector<int32_t> np(ctx->n_psm);
        for (size_t i = 0; i < np.size(); ++i) np[i] = ctx->psms[i].n_ptm;
        CU(ctx->k_idx_a.reserve((size_t)ctx->n_psm * 4));
        CU(cudaMemcpyAsync(ctx->k_idx_a.p, np.data(), np.size() * 4, cudaMemcpyHostToDevice, st));
        ctx->cnt.other_kernel_launches += (uint64_t)launch_set_nptm(ctx->d_psm.as<DevPsm>(), ctx->n_psm, ctx->k_idx_a.as<int32_t>(), st);
        CU(cudaStreamSynchronize(st));
        ctx->cnt.h2d_bytes += np.size() * 4;
    }
    ctx->step_done = 5;
    return pq_rank_and_validate(ctx);
}

// ---------------------------------------------------------------------------------------------------------------------
int32_t pq_score_pairs(pq_ctx* ctx, int64_t n_pairs, const int32_t* spectrum_index, const pq_form* forms, double* score,
                       int32_t* count_a, int32_t* count_b) {
    if (!ctx || n_pairs < 0 || (n_pairs > 0 && (!spectrum_index || !forms || !score))) return fail(ctx, PQ_ERR_BAD_ARG, "null argument");
    if (n_pairs == 0) return PQ_OK;
    if (n_pairs > 0x7ffffff0LL) return fail(ctx, PQ_ERR_UNSUPPORTED, "too many pairs for one call");
    cudaSetDevice(ctx->P.device);
    { int32_t rc0 = ensure_counters(ctx); if (rc0) return rc0; rc0 = ensure_host_spectra(ctx); if (rc0) return rc0; }
    cudaStream_t st = ctx->st;
    const int64_t n = ctx->nS;
    // caller index -> sorted index (a spectrum loaded without a charge is scored as its z = 2 reading)
    std::vector<int32_t> sorted_of((size_t)ctx->nCaller, -1);
    for (int64_t i = 0; i < n; ++i) if (ctx->h_alias[(size_t)i] != 2) sorted_of[(size_t)ctx->h_caller[i]] = (int32_t)i;
    // group pairs by spectrum; rows in that order
    std::vector<uint32_t> order((size_t)n_pairs);
    std::iota(order.begin(), order.end(), 0u);
    for (int64_t i = 0; i < n_pairs; ++i) {
        if (spectrum_index[i] < 0 || spectrum_index[i] >= ctx->nCaller) return fail(ctx, PQ_ERR_BAD_ARG, "spectrum index out of range");
        if (!ctx->h_resident[(size_t)sorted_of[(size_t)spectrum_index[i]]])
            return fail(ctx, PQ_ERR_STATE, "the peaks of spectrum " + std::to_string(spectrum_index[i]) + " are not on the device (candidate-first load: only spectra with a "
                                           "target in their precursor window were brought over); load with pq_params.candidate_first = 0 to score arbitrary pairs");
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return sorted_of[(size_t)spectrum_index[a]] < sorted_of[(size_t)spectrum_index[b]]; });
    std::vector<uint8_t> rows((size_t)n_pairs * kRowBytes);
    std::vector<int32_t> pred;
    std::string seq; FormRec fr;
    for (int64_t k = 0; k < n_pairs; ++k) {
        if (!ctx->codes.from_public(forms[order[(size_t)k]], seq, fr) || !ctx->codes.encode_row(seq, fr, rows.data() + (size_t)k * kRowBytes))
            return fail(ctx, PQ_ERR_BAD_ARG, "pair " + std::to_string(order[(size_t)k]) + ": form cannot be encoded");
    }
    // spectra to prepare = distinct spectra of the list
    std::vector<int32_t> list; std::vector<Job> hj;
    for (int64_t k = 0; k < n_pairs;) {
        int32_t s = sorted_of[(size_t)spectrum_index[order[(size_t)k]]];
        int64_t e = k;
        while (e < n_pairs && sorted_of[(size_t)spectrum_index[order[(size_t)e]]] == s) ++e;
        Job J; J.blob = (uint32_t)list.size(); J.cand_begin = (uint32_t)k; J.cand_count = (uint32_t)(e - k); J.out = (uint32_t)k;
        J.mass = ctx->h_mass[(size_t)s]; J.ref_score = 0; J.charge = ctx->h_charge[(size_t)s]; J.iso = 0; J.spectrum = s; J.aux = 0;
        hj.push_back(J); list.push_back(s);
        k = e;
    }
    DevBuf d_rows, d_score, d_a, d_b, d_list, d_pred;
    CU(d_rows.reserve(rows.size())); CU(d_score.reserve((size_t)n_pairs * 8)); CU(d_a.reserve((size_t)n_pairs * 4)); CU(d_b.reserve((size_t)n_pairs * 4));
    CU(d_list.reserve(list.size() * 4 + 8));
    CU(cudaMemcpyAsync(d_rows.p, rows.data(), rows.size(), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_list.p, list.data(), list.size() * 4, cudaMemcpyHostToDevice, st));
    SpectraView sp = ctx->view();
    // the call prepares its spectra into its own blob pool: the step pipeline's blobs stay as they are
    DevBuf& p_blobs = ctx->pp_blobs; DevBuf& p_off = ctx->pp_blob_off; DevBuf& p_len = ctx->pp_blob_len;
    CU(ctx->scratch.reserve((list.size() + 2) * 8)); CU(p_off.reserve((list.size() + 2) * 8)); CU(p_len.reserve((list.size() + 2) * 4));
    launch_blob_sizes(sp, d_list.as<int32_t>(), (int32_t)list.size(), ctx->scratch.as<uint64_t>(), st);
    int32_t rc = scan_exclusive(ctx, ctx->scratch.as<uint64_t>(), p_off.as<uint64_t>(), (int)list.size() + 1);
    if (rc) return rc;
    uint64_t total_blob = 0;
    CU(cudaMemcpyAsync(&total_blob, p_off.as<uint64_t>() + list.size(), 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    CU(p_blobs.reserve(total_blob + 256));
    PrepConfig pc; pc.itol = ctx->P.itol; pc.method = ctx->P.score_method;
    pc.mvh_bins = (int32_t)jround((double)(int)(2000.0 - 200.0) / (ctx->P.itol * 2.0));
    {
        Timer t(ctx, &ctx->cnt.ms_prepare);
        ctx->cnt.other_kernel_launches += (uint64_t)launch_prepare(sp, d_list.as<int32_t>(), (int32_t)list.size(), pc, p_blobs.as<uint8_t>(), p_off.as<uint64_t>(), p_len.as<uint32_t>(), ctx->max_peaks, st) + 2;
    }
    CU(ctx->jobs.reserve(hj.size() * sizeof(Job)));
    CU(cudaMemcpyAsync(ctx->jobs.p, hj.data(), hj.size() * sizeof(Job), cudaMemcpyHostToDevice, st));
    CU(ctx->results.reserve(hj.size() * sizeof(JobResult))); CU(ctx->hits.reserve(64)); CU(ctx->hit_count.reserve(64));
    CU(cudaMemsetAsync(ctx->hit_count.p, 0, 4, st));
    ScoreLaunch L; memset(&L, 0, sizeof L);
    L.kind = kJobPairs; L.jobs = ctx->jobs.as<Job>(); L.n_jobs = (uint32_t)hj.size();
    L.blobs = p_blobs.as<uint8_t>(); L.blob_off = p_off.as<uint64_t>(); L.blob_len = p_len.as<uint32_t>(); L.rows = d_rows.as<uint8_t>(); L.row_mass = nullptr;
    L.mods = ctx->d_mods.as<ModTable>(); L.log10_fact = ctx->d_l10.as<double>(); L.ln_fact = ctx->d_lnf.as<double>(); L.row_pred = nullptr; L.n_codes = (uint32_t)ctx->codes.n_codes;
    L.results = ctx->results.as<JobResult>(); L.hits = ctx->hits.as<Hit>(); L.hit_count = ctx->hit_count.as<uint32_t>(); L.hit_capacity = 0;
    L.pair_score = d_score.as<double>(); L.pair_a = d_a.as<int32_t>(); L.pair_b = d_b.as<int32_t>(); L.max_blob_bytes = blob_bytes(ctx->max_peaks); L.job_counter = ctx->d_scalars.as<uint32_t>() + 12;
    if (ctx->P.score_method == 2) {
        CU(d_pred.reserve((size_t)n_pairs * 8));
        ctx->cnt.other_kernel_launches += (uint64_t)launch_predict(d_rows.as<uint8_t>(), (uint32_t)n_pairs, ctx->d_mods.as<ModTable>(), ctx->P.itol, d_pred.as<int32_t>(), st);
        L.row_pred = d_pred.as<int32_t>();
    }
    ScoreConfig cfg = score_config(ctx, 0);
    {
        Timer t(ctx, &ctx->cnt.ms_score);
        ctx->cnt.score_kernel_launches += (uint64_t)launch_score(L, cfg, st);
    }
    CU(cudaGetLastError());
    std::vector<double> hs((size_t)n_pairs); std::vector<int32_t> ha((size_t)n_pairs), hb((size_t)n_pairs);
    CU(cudaMemcpyAsync(hs.data(), d_score.p, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ha.data(), d_a.p, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(hb.data(), d_b.p, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int64_t k = 0; k < n_pairs; ++k) {
        uint32_t o = order[(size_t)k];
        score[o] = hs[(size_t)k]; if (count_a) count_a[o] = ha[(size_t)k]; if (count_b) count_b[o] = hb[(size_t)k];
    }
    return PQ_OK;
}

}  // extern "C"

// pq_internal.h — device views and kernel launchers shared between the translation units of libpepquery_b200.so.
#pragma once
#include <string>
#include <vector>

#include "pq_common.h"

namespace pq {

// grow-only device buffer
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept { if (this != &o) { if (p) cudaFree(p); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; } return *this; }
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Spectra as they sit in HBM: CSR peak arrays in precursor-mass order (SURVEY.md §8 a15 / north_star item 1).
struct SpectraView {
    int64_t n;                    // (virtual) spectra: a spectrum loaded without a charge counts twice, as z = 2 and as z = 3
    const double* prec_mz;        // [n] sorted by neutral mass
    const int32_t* charge;        // [n] charge in use
    const double* mass;           // [n] neutral precursor mass  mz*z - z*proton, ascending
    const uint64_t* off;          // [n+1] peak offsets into mz / inten; a spectrum whose peaks are not resident has length 0 here
    const double* mz;             // peaks, file order inside a spectrum
    const double* inten;
    const int32_t* caller_index;  // [n] sorted position -> index in the caller's order
    const uint32_t* npeaks;       // [n] peak count of the spectrum (resident or not)
    const uint8_t* alias;         // [n] 0: charge given; 1: charge-less searched as z = 2; 2: charge-less searched as z = 3
};

struct PrepConfig {
    double itol;
    int32_t method;               // 1 hyperscore, 2 MVH
    int32_t mvh_bins;             // round(1800 / (2*itol))
};

// One unit of scoring work: one spectrum x a run of candidate peptide rows.
struct Job {
    uint32_t blob;                // prepared-spectrum id
    uint32_t cand_begin;          // first candidate row (in the table the launch names)
    uint32_t cand_count;
    uint32_t out;                 // result slot of the job
    double   mass;                // precursor neutral mass this job matches against (isotope-corrected)
    double   ref_score;           // Step 3: best target score of the spectrum; Step 4: the PSM's score
    int32_t  charge;              // precursor charge
    int32_t  iso;                 // isotope error of `mass`
    int32_t  spectrum;            // sorted spectrum index
    int32_t  aux;                 // Step 4: target form; Step 5: ptm index
};

// What a launch does with the scores of a job.
enum JobKind : int32_t { kJobTargets = 2, kJobCompetitors = 3, kJobShuffles = 4, kJobPtm = 5, kJobPairs = 1 };

struct JobResult {            // per job, reduced in the CTA
    uint32_t n_in_window;     // candidates passing the precursor window (= pairs scored)
    uint32_t n_better;        // Step 3: score >= ref_score; Step 4: score >= ref_score; Step 5: better per -hc rule
    double   best_score;      // max score over the job (-1 if none)
    uint32_t best_cand;       // lowest candidate row attaining it
    uint32_t pad;             // Step 4: pairs whose comparison with ref_score was decided by the score upper bound
};

struct Hit {                  // appended rows (Step 2 PSMs, Step 3 detail rows, Step 5 ptm rows)
    int32_t  spectrum;        // sorted spectrum index
    uint32_t cand;            // candidate row
    double   score;
    double   mass;            // precursor mass of the job
    int32_t  count_a, count_b;
    int32_t  iso;
    int32_t  aux;             // Step 5: (ptm << 8) | site
    uint32_t blob;            // prepared-spectrum id of the job
    uint32_t pad;
};

struct ScoreConfig {
    double itol;
    double tol;
    int32_t tol_ppm;
    int32_t method;
    double min_score;
    int32_t fast, hc;
    int32_t mvh_bins;
    int32_t check_window;     // candidates must pass the bucket + tolerance predicate (Steps 2, 3, 5)
    int32_t score_all;        // Step 4: compute every shuffle's score in full (no upper-bound decision)
    int32_t prefetch_next;    // one-stage K1 jobs: take the next job early and pull its blob towards L2 (off: PQ_NO_BLOB_PREFETCH)
    int32_t no_bucket;        // Step 2 of a library search: only |mass error| <= tol, no +-1 rule on the 0.1-Da buckets (MsLibrarySearchWorker.java:118-120)
};

struct ShuffleForm;
// SpectraInput.pep2targeted_spectra on the device: sorted keys (canonical target sequence << 32 | caller spectrum index)
struct TargetedView {
    const uint64_t* keys; uint32_t n;                 // n = 0: no list
    const uint8_t* seq_listed;                        // [sequences] the (canonical) sequence appears in the list
    const uint32_t* canon;                            // [sequences] first sequence with the same letters
    uint8_t* seen;                                    // [n] the pair met in a precursor window (PsmType.spectrum_matched)
    const uint32_t* form_of_row; const int32_t* seq_of_form; const int32_t* caller_index; const uint8_t* alias;
};
struct ScoreLaunch {
    uint32_t n_codes = kMaxCodes;   // residue codes in use (26 letters + the (residue, modification) pairs of this search)
    JobKind kind;
    const Job* jobs; uint32_t n_jobs;
    const uint8_t* blobs; const uint64_t* blob_off; const uint32_t* blob_len;   // prepared spectra (slot offsets, used bytes)
    const uint8_t* rows; const double* row_mass;           // candidate table (64-byte rows) + masses (may be null)
    const ModTable* mods;
    const double* log10_fact;                              // [65]  log10(n!)
    const double* ln_fact;                                 // [n]   ln(n!)  (MVH)
    const int32_t* row_pred;                               // MVH: predicted-peak counts per row, [2*row]: maxc 1 / 2
    JobResult* results;                                    // [n_jobs]
    Hit* hits; uint32_t* hit_count; uint32_t hit_capacity; // appended rows
    double* pair_score; int32_t* pair_a; int32_t* pair_b;  // kJobPairs: per-candidate outputs
    uint32_t max_blob_bytes;
    uint32_t* job_counter;                                 // device word, zeroed by the launcher: next job to hand out
    // Step 4: cell lists of the shuffle rows (k_cells.cuh); null = walk the ladder in the count pass
    const ShuffleForm* sh_forms; const uint32_t* form_slot; const uint4* lists;
    TargetedView tgt;                                      // Step 2: targeted PSM list
    // Step 3: precomputed fragment ladders of the candidate rows (k1_ladder.cu); null = walk the ladder per pair
    const uint4* lad; uint32_t lad_stride;                 // row r at lad + r * lad_stride (16-byte units)
    const uint8_t* lad_len;                                // peptide length of every row (one byte each)
    // Step 5
    double ptm_delta; int32_t ptm_position; int32_t ptm_residue_code;
};

// kernel launchers (each returns the number of kernels it launched; errors via cudaGetLastError by the caller)
int launch_prepare(const SpectraView& sp, const int32_t* list, int32_t n_list, const PrepConfig& cfg, uint8_t* blobs,
                   const uint64_t* blob_off, uint32_t* blob_len, uint32_t max_peaks, cudaStream_t st, uint32_t min_peaks = 0,
                   const int32_t* blob_index = nullptr, int leave_blocks_per_sm = 0, uint32_t* work_counters = nullptr);
int launch_score(const ScoreLaunch& L, const ScoreConfig& cfg, cudaStream_t st);
int launch_score_ladder(const ScoreLaunch& L, const ScoreConfig& cfg, cudaStream_t st);      // Step 3 over rows with precomputed ladders (k1_ladder.cu)
// the ladder tables of a row table whose peptides have at most max_len residues: one fixed-size slot per row
bool build_row_ladders(cudaStream_t st, const uint8_t* rows, uint32_t n, int max_len, const ModTable* mods, DevBuf& lad, DevBuf& lad_len, uint32_t* stride_units,
                       uint64_t* launches, std::string& err);

// Step-2 window enumeration on device: for every usable spectrum x isotope error, the run of target rows whose mass
// can fall in the precursor window.  Fills jobs (capacity n_spectra * n_isotope) and returns the count through d_n_jobs.
struct WindowConfig { double tol; int32_t tol_ppm; int32_t n_isotope; int32_t isotope[PQ_MAX_ISOTOPE]; int32_t min_peaks; int32_t no_bucket; };
int launch_enumerate_targets(const SpectraView& sp, const double* table_mass, uint32_t n_rows, const WindowConfig& cfg,
                             Job* jobs, uint32_t* d_n_jobs, int32_t* need_prepare /* [n] flags */, cudaStream_t st);

// Shuffle generation (K2): replays java.util.Random per target sequence, de-duplicates, then places the
// modifications of each target form on each shuffle (second Random stream) and writes candidate rows.
struct ShuffleSeq {            // one per target sequence that needs shuffles
    uint32_t len;              // L
    uint32_t n_groups;
    uint32_t num;              // rounds to run (min(n_random, nComb))
    uint32_t out_base;         // first raw-shuffle slot
    uint8_t  group_code[32];   // residue letter code (0..25) per group, in the reference's group order
    uint8_t  group_count[32];
    uint8_t  seq[kMaxLen + 4]; // target letters (codes 0..25)
};
struct ShuffleForm {           // one per target form that needs shuffles
    uint32_t seq_slot;         // index into ShuffleSeq[]
    uint32_t row_base;         // first output row
    uint32_t n_mods;
    int8_t   mod[PQ_MAX_FORM_MODS];   // modification ids (index into ShuffleMods), target list order
    // cell lists of the form's shuffle rows (k_cells.cuh), unit-major: unit u of row k at lists[list_base + u * list_rows + k]
    uint32_t len;              // peptide length L
    uint32_t list_nch;         // fragment charges held in the lists (1..kListMaxCharge): the largest a still-valid PSM of the form needs
    uint32_t list_chunks;      // units per side
    uint32_t list_rows;        // rows reserved (= rounds planned for the sequence)
    uint64_t list_base;
};
struct ShuffleMods {           // the search's modifications as K2 needs them
    int32_t n_mods;            // var then fixed
    int32_t n_var;
    int32_t position[2 * PQ_MAX_MODS];
    uint32_t residue_mask[2 * PQ_MAX_MODS];    // bit r set: letter code r is a target residue
    uint8_t  code_of[2 * PQ_MAX_MODS][26];     // residue code for (mod, letter)
    uint8_t  term_code[2 * PQ_MAX_MODS];       // terminal code for site-0 / site-L+1 modifications
    int64_t shuffle_seed, mod_seed;
};
int launch_shuffles(const ShuffleSeq* seqs, uint32_t n_seqs, const ShuffleForm* forms, uint32_t n_forms,
                    const ShuffleMods* mods, uint8_t* raw /* n_seqs*num*64 */, uint32_t* kept_index, uint32_t* kept_count,
                    uint8_t* rows, uint32_t max_num, int64_t shuffle_seed, const ModTable* table, uint4* lists /* may be null */, cudaStream_t st);

// packing helpers (k3_window.cu)
int launch_virtual_meta(int64_t nv, int64_t n, const double* prec_mz, const int32_t* charge, const int32_t* zero_idx, double* mass, int32_t* vz,
                        uint64_t* key, int32_t* idx, cudaStream_t st);
int launch_gather_meta(int64_t nv, int64_t n, const int32_t* order, const double* prec_mz_in, const int32_t* charge_in, const int32_t* zero_idx,
                       const int32_t* vz, const double* mass_in, const uint64_t* off_in, double* prec_mz, int32_t* charge, double* mass,
                       int32_t* caller, uint32_t* npeaks, uint8_t* alias, cudaStream_t st);
int launch_flags_by_virtual(int64_t nv, const int32_t* inv, const int32_t* flag_sorted, uint8_t* flag_virtual, cudaStream_t st);
int launch_resident_counts(int64_t nv, const uint32_t* npeaks, const int32_t* resident, uint64_t* count, cudaStream_t st);
int launch_gather_peaks_from_host(const int32_t* list, int32_t n_list, const int32_t* caller, const uint64_t* off_in, const uint64_t* off_out,
                                  const double* h_mz, const double* h_in, double* mz_out, double* in_out, int sms, cudaStream_t st);
int launch_gather_peaks_chunked(const int32_t* list, int32_t n_list, int32_t chunk, const int32_t* caller, const uint64_t* off_in, const uint64_t* off_out,
                                const double* h_mz, const double* h_in, double* mz_out, double* in_out, uint32_t* chunk_done /* zeroed, one word per chunk */,
                                uint32_t* n_ctas, cudaStream_t st);
int launch_wait_chunk(const uint32_t* flag, uint32_t expected, cudaStream_t st);      // the stream goes on once *flag >= expected
// candidate-first from pageable caller arrays: what to pack (caller spectrum + peak count per gather-list entry), and a packed chunk into the CSR
int launch_gather_sources(const int32_t* list, int32_t n, const int32_t* caller, const uint32_t* npeaks, int32_t* src, uint64_t* cnt /* [n + 1] */, cudaStream_t st);
int launch_scatter_staged(const int32_t* list, int32_t n, const uint64_t* g_off, uint64_t base, const uint64_t* off_out, const double* st_mz, const double* st_in,
                          double* mz_out, double* in_out, uint32_t* flag, uint32_t flag_value, cudaStream_t st);
int launch_invert_order(int64_t n, const int32_t* order, int32_t* inv, cudaStream_t st);
int launch_gather_peaks_by_virtual(int64_t first, int64_t count, int64_t n, const int32_t* zero_idx, const int32_t* inv, const uint64_t* off_in,
                                   const uint64_t* off_out, const double* mz_in, const double* in_in, double* mz_out, double* in_out, cudaStream_t st);
int launch_count_missing(int64_t n, const int32_t* need, const int32_t* resident, uint32_t* out, cudaStream_t st);
int launch_chargeless_filter(const Hit* hits, uint32_t nh, const Job* jobs, uint32_t n_jobs, const SpectraView& sp, int64_t n_caller, const int32_t* zero_idx,
                             const int32_t* vid, const int32_t* inv, uint8_t* has_hit /* zeroed */, uint8_t* skip_job, uint8_t* keep_hit, cudaStream_t st);
int launch_blob_sizes(const SpectraView& sp, const int32_t* list, int32_t n_list, uint64_t* sizes, cudaStream_t st);
int launch_fix_job_blobs(Job* jobs, uint32_t n_jobs, const int32_t* prep_id, cudaStream_t st);
int launch_job_keys(const Job* jobs, uint32_t n, uint64_t* key, uint32_t* idx, cudaStream_t st);
int launch_gather_jobs(const Job* in, const uint32_t* idx, uint32_t n, Job* out, cudaStream_t st);
// Step 5 (k5_ptm.cu): one job per validated spectrum; candidates are enumerated inside the kernel
struct PtmLaunch {
    const Job* jobs; uint32_t n_jobs;                      // mass = precursor neutral mass, ref_score = target PSM score
    const uint8_t* blobs; const uint64_t* blob_off; const uint32_t* blob_len;
    const uint8_t* rows; const double* row_mass; const uint32_t* row_mask; uint32_t n_rows;   // reference table
    const pq_ptm* ptms; int32_t n_ptm;
    int32_t n_isotope; int32_t isotope[PQ_MAX_ISOTOPE];
    const ModTable* mods; const double* log10_fact; const double* ln_fact;
    JobResult* results; Hit* hits; uint32_t* hit_count; uint32_t hit_capacity;
    uint32_t max_blob_bytes;
};
int launch_ptm(const PtmLaunch& L, const ScoreConfig& cfg, cudaStream_t st);
int launch_row_masks(const uint8_t* rows, uint32_t n_rows, uint32_t* mask, cudaStream_t st);

// ---- device-resident step pipeline (k7_steps.cu) --------------------------------------------------------------------
// The PSM table lives on the device from Step 2 to the final export; the host only sees counts.
struct DevPsm {
    int32_t  spectrum_sorted;   // index into the mass-sorted spectra
    int32_t  spectrum;          // caller's spectrum index
    int32_t  form;              // public target form index
    uint32_t row;               // target table row
    uint32_t blob;              // prepared-spectrum id
    int32_t  charge, iso, len;
    double   score, exp_mass, pep_mass, p_value;
    int32_t  ca, cb, n_db, total_db, n_random, total_random, valid, rank, n_ptm, confident;
    uint32_t seg;               // spectrum segment (PSMs of one spectrum are contiguous)
    uint32_t pad;
};
struct SegAgg {                 // Step-3 result per matched spectrum
    int32_t matched, better;
    double  best, ref_best;
    uint32_t best_row, pad;
};
struct DevCounters { unsigned long long pairs[4]; unsigned long long bytes[4]; unsigned long long bounded; };
struct FormMods { uint32_t n_mods; int8_t mod[PQ_MAX_FORM_MODS]; uint8_t site[PQ_MAX_FORM_MODS]; };     // a target form's modifications, reference list order
struct ShuffleTotals { uint32_t n_seqs, raw_rows, n_forms, out_rows, n_valid, pad; uint64_t list_units; };

struct TargetAux {              // per-target-table device arrays
    const uint32_t* form_of_row;      // [rows]  mass-sorted row -> public form
    const int32_t*  seq_of_form;      // [forms]
    const uint16_t* piece_rank;       // [(var + fixed mods) * 64]: rank of the text "id@site[delta]" among all such pieces (RankPSM string tie-break)
    const FormMods* form_mods;        // [forms]
    const uint32_t* seq_off;          // [seqs+1]
    const uint8_t*  seq_letters;      // letter codes 0..25
    uint32_t n_seq, n_forms;
};

// Step 2: hits -> sorted PSM table, segments, output permutation.  Returns false + err on CUDA failure.
bool psm_table_from_hits(const Hit* hits, uint32_t nh, const SpectraView& sp, const double* t_mass, const uint8_t* t_rows, const TargetAux& A,
                         DevBuf& key_a, DevBuf& key_b, DevBuf& idx_a, DevBuf& idx_b, DevBuf& head, DevBuf& tmp,
                         DevPsm* psms, uint32_t* seg_start, uint32_t* d_nseg, uint32_t* out_perm, cudaStream_t st, int* launches, std::string& err);
// merging the parts of a search that ran during the upload (search_at_load)
int launch_merge_psms(const DevPsm* in, uint32_t n, uint32_t seg_off, DevPsm* out, cudaStream_t st);
int launch_merge_segs(const uint32_t* seg_start_in, const SegAgg* agg_in, uint32_t n_seg, uint32_t psm_off, uint32_t row_off, uint32_t* seg_start_out, SegAgg* agg_out,
                      uint32_t end_value, int last, cudaStream_t st);
int launch_merge_hits(const Hit* in, uint32_t n, uint32_t row_off, Hit* out, cudaStream_t st);
int launch_part_job_keys(const Job* jobs, uint32_t n, const uint32_t* rank_in_upload, uint32_t part_spectra, uint32_t shift, uint64_t* key, uint32_t* idx, cudaStream_t st);
int launch_part_bounds(const uint64_t* sorted_key, uint32_t n, uint32_t n_parts, uint32_t* first, cudaStream_t st);
int launch_part_cand(const Job* jobs, uint32_t n, const uint64_t* sorted_key, unsigned long long* cand, cudaStream_t st);
int launch_window_rows(const double* mass, uint32_t n, double lo, double hi, uint32_t* ab, cudaStream_t st);
int launch_scatter_rank(const int32_t* list, uint32_t n, uint32_t* rank, cudaStream_t st);
int launch_fill_u8(uint8_t* p, uint32_t n, uint8_t v, cudaStream_t st);
int launch_fill_u32(uint32_t* p, uint32_t n, uint32_t v, cudaStream_t st);
bool build_out_perm(const DevPsm* psms, uint32_t n, DevBuf& key_a, DevBuf& key_b, DevBuf& idx_a, DevBuf& tmp, uint32_t* out_perm, cudaStream_t st, std::string& err);
int launch_job_counters(const Job* jobs, const JobResult* res, uint32_t n_jobs, const uint32_t* sp_npeaks, int which, DevCounters* c, cudaStream_t st,
                        const uint8_t* skip_job = nullptr);
int launch_step3_jobs(const DevPsm* psms, const uint32_t* seg_start, uint32_t n_seg, const double* r_mass, uint32_t n_rows, const WindowConfig& W,
                      const double* sp_mass, Job* jobs, SegAgg* agg, unsigned long long* total_cand, cudaStream_t st);
int launch_step3_aggregate(const JobResult* res, uint32_t n_seg, int n_iso, const uint32_t* seg_start, DevPsm* psms, SegAgg* agg, cudaStream_t st);
int launch_comp3_rows(const Hit* hits, uint32_t nh, const SegAgg* agg, uint32_t n_seg, const uint32_t* seg_start, const DevPsm* psms,
                      const int32_t* caller_index, const double* r_mass, pq_competitor* rows, uint64_t* keys, uint32_t* idx, uint32_t* count, cudaStream_t st);
int launch_gather_comp(const pq_competitor* in, const uint32_t* idx, uint32_t n, pq_competitor* out, cudaStream_t st);
int launch_flag_valid(const DevPsm* psms, uint32_t n, const int32_t* seq_of_form, uint8_t* seq_flag, uint32_t* form_flag /* largest fragment-charge count of a valid PSM, 0 = none */,
                      uint8_t* valid_flag, ShuffleTotals* totals /* zeroed by the caller */, cudaStream_t st);
int launch_valid_flags(const DevPsm* psms, uint32_t n, uint8_t* valid_flag, uint32_t* n_valid /* zeroed by the caller */, cudaStream_t st);
int launch_shuffle_plans(const uint8_t* seq_flag, const TargetAux& A, int32_t n_random, ShuffleSeq* plans, cudaStream_t st);
int launch_shuffle_layout(const uint8_t* seq_flag, const uint32_t* form_flag, const uint8_t* valid_flag, uint32_t n_psm, const ShuffleSeq* plans, const TargetAux& A,
                          uint32_t* seq_slot, uint32_t* form_slot, uint32_t* row_base, ShuffleSeq* seqs, ShuffleForm* forms, ShuffleTotals* totals, cudaStream_t st);
int launch_step4_jobs(const DevPsm* psms, const uint32_t* order, uint32_t n_valid, const uint32_t* form_slot, const ShuffleForm* forms, const uint32_t* kept_count,
                      Job* jobs, cudaStream_t st);
int launch_step4_apply(const Job* jobs, const JobResult* res, const uint32_t* order, uint32_t n_valid, DevPsm* psms, cudaStream_t st);
int launch_psm_form_keys(const DevPsm* psms, const uint32_t* list, uint32_t n, uint32_t* keys, cudaStream_t st);
int launch_rank(DevPsm* psms, uint32_t n, const uint32_t* seg_start, const TargetAux& A, cudaStream_t st);
int launch_export_psms(const DevPsm* psms, const uint32_t* out_perm, uint32_t n, pq_psm* out, cudaStream_t st);
// Step-1 novelty filter (k8_novelty.cu): nhit[i] = 1 when target i occurs in a protein (I = L); false + err on refusal / CUDA error
bool find_in_proteome(int32_t n_seq, const uint32_t* seq_off, const char* res, int32_t n_prot, const uint64_t* prot_off, const char* prot_res,
                      int32_t* nhit, cudaStream_t st, uint64_t* h2d, uint64_t* d2h, int* launches, std::string& err);
int launch_delta_scores(const DevPsm* psms, const uint32_t* out_perm, uint32_t n, const SegAgg* agg, const int32_t* mod_key, const double* mod_best,
                        uint32_t n_mod, double* ref_delta, double* mod_delta, cudaStream_t st);
int launch_set_nptm(DevPsm* psms, uint32_t n, const int32_t* n_ptm, cudaStream_t st);
int launch_targeted_status(const DevPsm* psms, uint32_t n_psm, const TargetedView& V, unsigned long long* best /* [V.n] scratch */, int32_t* type_sorted /* [V.n] */, cudaStream_t st);

// MVH predicted-peak counts per row (k4_predict.cu): out[2*row] for precursor charge 2, out[2*row+1] otherwise
int launch_predict(const uint8_t* rows, uint32_t n_rows, const ModTable* mods, double itol, int32_t* out, cudaStream_t st);

}  // namespace pq

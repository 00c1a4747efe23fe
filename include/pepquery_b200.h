/*
 * pepquery_b200.h — C ABI of the B200-native PepQuery pair-scoring path.
 *
 * This is the drop-in boundary.  The reference (PepQuery 2.0.4,
 * 100 % Java) has no FFI seam; the seam is created at its batch fan-out loops.
 * Every entry point below names the reference loop it replaces (file:line under
 * /root/reference/src/main/java/).  A Java host binds these through JNI or
 * Panama FFM; tests bind them through ctypes.
 *
 * Rules of the boundary
 *   - plain C: pointers, sizes, POD structs; no C++/torch types;
 *   - the caller owns every input and output buffer; the library copies inputs to
 *     the device during the call and never keeps a host pointer after return;
 *   - every function returns int32 (PQ_OK = 0, <0 = pq_status); text through
 *     pq_last_error(); nothing throws or exits across the boundary;
 *   - there is NO CPU fallback: without a CUDA device pq_create fails with
 *     PQ_ERR_NO_DEVICE;
 *   - one pq_ctx per GPU; calls on one ctx are serialised by the caller; every call
 *     is synchronous at return.
 */
#ifndef PEPQUERY_B200_H
#define PEPQUERY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQ_VERSION_MAJOR 0
#define PQ_VERSION_MINOR 1

#define PQ_MAX_ISOTOPE 8
#define PQ_MAX_MODS 8
#define PQ_MAX_PEP_LEN 63      /* residues per peptide form handled on device (reference default max 45) */
#define PQ_MAX_FORM_MODS 16    /* modification matches carried per peptide form */

typedef enum pq_status {
    PQ_OK = 0,
    PQ_ERR_BAD_ARG = -1,
    PQ_ERR_CAPACITY = -2,     /* caller buffer too small; required size returned through the size out-param */
    PQ_ERR_CUDA = -3,
    PQ_ERR_NO_DEVICE = -4,    /* no CUDA device: there is no CPU fallback */
    PQ_ERR_STATE = -5,        /* call order violated (e.g. step3 before step2) */
    PQ_ERR_UNSUPPORTED = -6,
    PQ_ERR_NOMEM = -7
} pq_status;

/* Where a modification may sit.  Mirrors compomics ModificationType as used by
 * ModificationUtils.getPossibleModificationSites (call sites pg/PeptideInput.java:127,224,515;
 * OpenModificationSearch/ModificationDB.java:284). */
typedef enum pq_mod_position {
    PQ_MOD_ANYWHERE = 0,        /* modaa: every listed residue, site = 1-based residue index   */
    PQ_MOD_PEP_NTERM = 1,       /* modn_peptide: site 0                                         */
    PQ_MOD_PEP_CTERM = 2,       /* modc_peptide: site L+1                                       */
    PQ_MOD_PEP_NTERM_AA = 3,    /* modnaa_peptide: site 1 if residue 1 is listed                */
    PQ_MOD_PEP_CTERM_AA = 4,    /* modcaa_peptide: site L if residue L is listed                */
    PQ_MOD_PROT_NTERM = 5,      /* protein-terminal types: never placed on digest peptides here */
    PQ_MOD_PROT_CTERM = 6,
    PQ_MOD_PROT_NTERM_AA = 7,
    PQ_MOD_PROT_CTERM_AA = 8
} pq_mod_position;

/* A resolved modification: what pg/ModificationGear.java:100-132 hands to the search
 * (the Java host keeps the id -> name mapping; the shim only needs residue, mass, loss). */
typedef struct pq_mod {
    int32_t id;                 /* caller's id (e.g. the -fixMod/-varMod number); echoed in results    */
    int32_t position;           /* pq_mod_position                                                     */
    char    residues[24];       /* NUL-terminated residue letters for residue-specific positions       */
    double  delta;              /* monoisotopic mass shift                                             */
    double  neutral_loss;       /* mass of the modification-defined neutral loss (0 = none); only the
                                   MVH predicted-peak count uses it (PSMMatch/MVHMatch.java:150-199)   */
} pq_mod;

/* POD mirror of the pg/CParameter.java statics the hot path reads. */
typedef struct pq_params {
    double  tol;                /* precursor tolerance (CParameter.tol, default 10)                    */
    int32_t tol_ppm;            /* 1 = ppm (CParameter.tolu "ppm"), 0 = Da                             */
    int32_t score_method;       /* 1 = hyperscore, 2 = MVH (CParameter.scoreMethod)                    */
    double  itol;               /* fragment tolerance in Da (CParameter.itol, default 0.6)             */
    double  min_score;          /* CParameter.minScore, default 12                                     */
    int32_t min_peaks;          /* CParameter.minPeaks, default 10: spectra with <= this are skipped   */
    int32_t n_isotope;          /* entries of isotope[]; {0} by default (CParameter.isotope_error)     */
    int32_t isotope[PQ_MAX_ISOTOPE];
    int32_t max_var_mods;       /* CParameter.maxVarMods, default 3                                    */
    int32_t max_mods_per_aa;    /* CParameter.maxModsPerAA, default 1                                  */
    int32_t fast;               /* CParameter.fast_model                                               */
    int32_t hc;                 /* CParameter.unrestrictedSearchWithEqualAndBetterScore (-hc)          */
    int32_t n_random;           /* CParameter.nRandomPeptides (-n), reference default 10000            */
    int32_t enzyme;             /* -e index of pg/DatabaseInput.java:145-267: 1 Trypsin (default) ... 17
                                   LysargiNase; 0 = NoEnzyme: unspecific digestion, every stretch of a protein
                                   within the length and mass bounds (max_missed is not used then)        */
    int32_t max_missed;         /* CParameter.maxMissedCleavages, default 2                            */
    int32_t min_len, max_len;   /* 7 / 45                                                              */
    double  min_mass, max_mass; /* 500 / 10000 (digest mass window)                                    */
    int64_t shuffle_seed;       /* 12457 (pg/PeptideInput.java:433)                                    */
    int64_t mod_seed;           /* 2022  (pg/PeptideInput.java:500)                                    */
    int32_t n_fixed, n_var;
    pq_mod  fixed[PQ_MAX_MODS];
    pq_mod  var[PQ_MAX_MODS];
    int32_t device;             /* CUDA device ordinal                                                 */
    int32_t score_all_shuffles; /* 1 = compute every shuffle's score (hyperscore or MVH) in full.  0 (default): Step 4 only needs
                                   "shuffleScore >= targetScore" (pg/StatisticScoreWorker.java:51), so a shuffle whose
                                   score upper bound (from its count of possible matches) stays below the target's
                                   score is decided without resolving its matches; counts and p-values are identical */
    int32_t prepare_at_load;    /* 1 (default): pq_load_spectra also runs the peptide-independent spectrum preparation
                                   (HyperscoreMatch/MVHMatch.SpectrumPreProcessing + the annotator's window tables) for
                                   every usable spectrum, chunk by chunk while the next chunk of peaks is still crossing
                                   PCIe, so the upload hides it.  0: pq_step2_match_targets prepares only the spectra that
                                   have a target in their precursor window (less HBM, nothing overlapped).  Same results */
    int32_t candidate_first;    /* 1 (default): when pq_set_targets came before pq_load_spectra and the m/z and intensity arrays are in
                                   pinned (page-locked, device-mapped) memory, only spectra with a target form in their precursor window
                                   have their peaks moved to the device (the only ones Steps 2-5 read; the reference keeps exactly
                                   those in SpectraInput.spectraMap, pg/SpectraInput.java:243-254): a kernel gathers them from the
                                   caller's arrays over PCIe.  A later pq_set_targets with other targets then needs a new
                                   pq_load_spectra (PQ_ERR_STATE otherwise).  0: every peak is uploaded.  Same results */
    int32_t search_at_load;     /* 1 (default): when pq_load_spectra can tell everything a search needs — targets set, the database handed
                                   over by pq_stage_reference, candidate-first upload — it also runs Steps 2-4 on each part of the spectra as
                                   soon as that part has arrived, beside the rest of the upload.  pq_step2_match_targets,
                                   pq_build_reference(ctx, 0, NULL, NULL), pq_step3_competitors and pq_step4_shuffles then return what is
                                   already there.  0: every step runs when it is called.  Same results */
    int32_t library_mode;       /* 1: the spectra come from an indexed MS/MS library (pq_read_ms_library) and Step 2 applies the rules of
                                   msio/MsLibrarySearchWorker.java:42-153 instead of FindCandidateSpectraAndScoreWorker's: a spectrum needs
                                   n_peaks >= min_peaks (loadSpectra :239; > min_peaks otherwise) and a candidate only |mass error| <= tol
                                   (CParameter.get_mass_error), without the +-1 rule on 0.1-Da buckets.  0 (default) */
} pq_params;

/* One peptide form (sequence + placed modifications) as the caller sees it. */
typedef struct pq_form {
    int32_t seq_index;          /* index of the parent sequence in the table it came from              */
    int32_t length;
    double  mass;               /* H2O + residues + mods                                               */
    int32_t n_mods;
    uint8_t mod_site[PQ_MAX_FORM_MODS];   /* 0 = N-term, 1..L residues, L+1 = C-term                   */
    int8_t  mod_index[PQ_MAX_FORM_MODS];  /* >=0: var[idx]; <0: fixed[-idx-1]; in reference list order  */
    char    seq[PQ_MAX_PEP_LEN + 1];
} pq_form;

/* One target PSM and everything Steps 2-5 attach to it.  Field names follow the columns of
 * psm.txt / psm_rank.txt (pg/PeptideSearchMT.java:1634, pg/RankPSM.java:80, PSMT:1007). */
typedef struct pq_psm {
    int32_t spectrum;           /* index in the caller's spectrum order (pq_load_spectra)              */
    int32_t form;               /* target form index (pq_get_target_forms order)                       */
    double  score;              /* hyperscore or MVH                                                   */
    double  exp_mass;           /* precursor neutral mass (isotope-corrected mass that matched)        */
    double  pep_mass;
    double  tol_ppm;            /* signed, CParameter.get_mass_error                                   */
    int32_t isotope_error;
    int32_t count_b;            /* hyperscore: matched b / y ions; MVH: matched peaks / predicted      */
    int32_t count_y;
    int32_t n_db;               /* Step 3: reference forms scoring >= best target score of the spectrum */
    int32_t total_db;           /* Step 3: reference forms inside the precursor window                 */
    int32_t n_random;           /* Step 4: shuffles scoring >= target; -1 = not evaluated              */
    int32_t total_random;       /* Step 4: distinct shuffles produced; -1 = not evaluated              */
    int32_t valid;              /* 0 once Step 3 found a better reference match                        */
    double  p_value;            /* (n_random+1)/(total_random+1); 100 = not evaluated                  */
    int32_t rank;               /* pg/RankPSM.java:36-148, 1-based within the spectrum                 */
    int32_t n_ptm;              /* Step 5: PTM forms scoring > (>= with hc) the target; -1 = not run   */
    int32_t confident;          /* CParameter.passed_psm_validation(len,p,n_db,rank,n_ptm)             */
    int32_t charge;             /* precursor charge the PSM was scored with (2 or 3 for a spectrum loaded with charge 0) */
} pq_psm;

/* Row of detail.txt / ptm.txt: a competitor that matched a target spectrum. */
typedef struct pq_competitor {
    int32_t spectrum;
    int32_t ref_form;           /* reference form index (pq_get_reference_forms order)                 */
    int32_t ptm;                /* Step 5: index into the PTM table, -1 for Step 3 rows                */
    int32_t ptm_site;           /* Step 5: site carrying the PTM                                       */
    double  score;
    double  pep_mass;
} pq_competitor;

/* Status of a targeted PSM (pg/PsmType.java; TargetedPSM.get_psm_status, pg/TargetedPSM.java:29-46). */
typedef enum pq_psm_type {
    PQ_PSM_NO_SPECTRUM_MATCH = 0,        /* the peptide never met the spectrum in a precursor window                     */
    PQ_PSM_SPECTRUM_MATCHED = 1,
    PQ_PSM_LOW_SCORE = 2,                /* it did, but no form scored above min_score                                   */
    PQ_PSM_HIGH_P_VALUE = 3,
    PQ_PSM_BETTER_REF_PEP = 4,           /* n_db >= 1                                                                    */
    PQ_PSM_BETTER_REF_PEP_WITH_MOD = 5,  /* n_ptm >= 1                                                                   */
    PQ_PSM_NOT_TOP_RANK = 6,             /* rank >= 2                                                                    */
    PQ_PSM_CONFIDENT = 7,
    PQ_PSM_INVALID_PEPTIDE = 8           /* the Java host's call (it owns the Step-1 verdict); never returned from here  */
} pq_psm_type;

/* Unimod-like PTM specificity for Step 5 (OpenModificationSearch/ModificationDB.java:334-533). */
typedef struct pq_ptm {
    int32_t id;
    int32_t position;           /* pq_mod_position                                                     */
    char    residue;            /* 0 for terminal-only specificities                                   */
    char    pad[7];
    double  delta;
} pq_ptm;

/* Work and time counters (what the Java log lines at PSMT:476,574,589 report). */
typedef struct pq_counters {
    uint64_t pairs_step2, pairs_step3, pairs_step4, pairs_step5;   /* scorePeptide2Spectrum calls the
                                                                       reference would have made       */
    uint64_t spectra_loaded, spectra_usable, spectra_prepared;
    uint64_t target_forms, reference_sequences, reference_forms, psms;
    uint64_t score_kernel_launches, other_kernel_launches;
    uint64_t score_algorithmic_bytes;   /* sum over scored pairs of 16*P_s + 64 + 32   */
    double   ms_prepare, ms_window, ms_score, ms_shuffle, ms_total; /* CUDA-event time on the ctx stream */
    uint64_t h2d_bytes, d2h_bytes;
    double   ms_score_step[4];          /* K1 launch time of Steps 2,3,4,5 (CUDA events around the launch)       */
    uint64_t bytes_score_step[4];       /* algorithmic bytes of those launches                                   */
    double   ms_host_reference;         /* wall time of the reference-table build (digest + forms + upload)      */
    uint64_t pairs_step4_bounded;       /* Step-4 pairs decided by the score upper bound (subset of pairs_step4)   */
    uint64_t pairs_step5_bounded;       /* Step-5 pairs decided by the score upper bound (subset of pairs_step5)   */
    double   ms_ladder;                 /* fragment-ladder tables of the reference rows (built once per reference table, read by Step 3) */
} pq_counters;

typedef struct pq_ctx pq_ctx;

int32_t     pq_version(void);
const char* pq_last_error(const pq_ctx* ctx);          /* ctx may be NULL: error of the last failed pq_create */
void        pq_default_params(pq_params* p);           /* defaults of pg/CParameter.java:84-463 with -n 1000   */

int32_t pq_create(const pq_params* params, pq_ctx** out);
int32_t pq_destroy(pq_ctx* ctx);

/* Spectrum loader.  Replaces the MGF->task loop of pg/SpectraInput.java:142-198 (readMSMSfast):
 * packs peak lists into precursor-mass-sorted CSR device arrays; a precursor window is a binary search on the sorted masses
 * followed by the reference's exact predicate (0.1-Da bucket within +-1 and |delta| <= tol) per candidate.
 * Peaks may be in any order; spectra with n_peaks <= min_peaks are kept but never scored (SpectraInput.java:183).
 * charge[i] >= 1, or 0 for a spectrum without a charge: it is searched as z = 2 and, only if that saves no PSM, as z = 3
 * (pg/FindCandidateSpectraAndScoreWorker.java:96-228); pq_psm.charge tells which.
 * With pq_set_targets called first and the m/z and intensity arrays in pinned memory only the spectra that have a target in
 * their precursor window are moved to the device (pq_params.candidate_first); otherwise every peak is uploaded in chunks.
 * Either way the work that does not depend on the matched spectra runs under the transfer (spectrum preparation with
 * prepare_at_load; the digest of a database handed over by pq_stage_reference).  A load that fails leaves the context
 * without spectra and with nothing in flight. */
int32_t pq_load_spectra(pq_ctx* ctx, int64_t n_spectra, const double* precursor_mz, const int32_t* charge,
                        const uint64_t* peak_offset /* n+1 */, const double* mz, const double* intensity);

/* The same from MGF text (replaces the compomics MgfFileIterator loop of pg/SpectraInput.java:150-195): the text goes
 * to the device in chunks and is parsed THERE — line starts, line classes, prefix scans for spectrum and peak indices, numbers through
 * Clinger's exact fast path in FP64 (the same doubles as strtod / Double.parseDouble; the rare number of another shape is parsed by the
 * host's strtod and patched in) — straight into the arrays pq_load_spectra would have been given; then the loader runs as above
 * (candidate-first, search_at_load ...).  Spectra are numbered in file order; a block without CHARGE has charge 0.  Blocks must be well
 * formed (every BEGIN IONS closed by END IONS); otherwise PQ_ERR_BAD_ARG, and pq_parse_mgf + pq_load_spectra take the file. */
int32_t pq_load_mgf(pq_ctx* ctx, const char* text, uint64_t n_bytes, int64_t* n_spectra);
/* The device parser alone, with pq_parse_mgf's contract (first call with NULL outputs returns the sizes, the second fills the caller's
 * arrays — it re-uses what the first one parsed when given the same text).  Discards the spectra loaded in ctx. */
int32_t pq_parse_mgf_device(pq_ctx* ctx, const char* text, uint64_t n_bytes, int64_t* n_spectra, uint64_t* n_peaks, double* precursor_mz,
                            int32_t* charge, uint64_t* peak_offset, double* mz, double* intensity);

/* Target peptides.  Replaces CreatePeptideInputWorker -> PeptideInput.calcPeptideIsoforms
 * (pg/PeptideInput.java:114-237): expands each sequence into its modification forms (lazily: under the next
 * pq_load_spectra, else in pq_step2_match_targets / pq_get_target_forms). */
int32_t pq_set_targets(pq_ctx* ctx, int32_t n_seq, const uint32_t* seq_offset /* n+1 */, const char* residues);
int32_t pq_get_target_forms(pq_ctx* ctx, pq_form* out, int64_t capacity, int64_t* n_forms);
/* Targeted PSMs.  Replaces SpectraInput.pep2targeted_spectra (filled at pg/PeptideSearchMT.java:264-309 when the peptide input names
 * spectra): pairs (target sequence index, spectrum index in the pq_load_spectra order).  In Step 2 a target sequence that appears in
 * the list is only scored against the spectra named for it (pg/FindCandidateSpectraAndScoreWorker.java:60-74, :126-134, :188-197);
 * sequences that do not appear are unrestricted.  Call after pq_set_targets (which clears the list) and before
 * pq_step2_match_targets; n_pairs = 0 clears it. */
int32_t pq_set_targeted_psms(pq_ctx* ctx, int64_t n_pairs, const int32_t* target_seq, const int32_t* spectrum);
/* One pq_psm_type per targeted pair, in the order they were given: PeptideSearchMT.export_targeted_psm_status
 * (pg/PeptideSearchMT.java:1484-1589) from the ranked PSM table — the best-ranked row of (peptide sequence, spectrum) decides; without a
 * row: LOW_SCORE when Step 2 saw the pair in a precursor window (a spectrum that came with a charge), else NO_SPECTRUM_MATCH.
 * Call after pq_rank_and_validate (and pq_step5_ptm, if it runs). */
int32_t pq_get_targeted_psm_status(pq_ctx* ctx, int32_t* type, int64_t capacity, int64_t* n_pairs);

/* Reference database.  Replaces DatabaseInput.generate_reference_peptide_index
 * (pg/DatabaseInput.java:402-556) + DigestProteinWorker.run (OpenModificationSearch/DigestProteinWorker.java:64-137)
 * + GeneratePeptideformWorker.run (:25-52): digest, I->L, unique, minus targets, isoforms, mass window,
 * mass-sorted device table (window lookup = binary search + the exact +-1-bucket predicate).  Call after pq_step2 (the mass
 * window is taken from the matched spectra, GeneratePeptideformWorker.set_mass_range). */
int32_t pq_build_reference(pq_ctx* ctx, int32_t n_proteins, const uint64_t* seq_offset /* n+1 */, const char* residues);
/* Optional: hand the protein database over early (any time before pq_build_reference).  The bytes go to the device at once; the
 * part of the index build that does not depend on the spectra — digestion, unique sequences, minus the targets
 * (DatabaseInput.protein_digest :402-478) — then runs inside the next pq_load_spectra, beside the upload of the peaks, and
 * pq_build_reference(ctx, 0, NULL, NULL) only expands the isoforms inside the matched-spectra mass window.  Same table. */
int32_t pq_stage_reference(pq_ctx* ctx, int32_t n_proteins, const uint64_t* seq_offset /* n+1 */, const char* residues);
int32_t pq_get_reference_forms(pq_ctx* ctx, pq_form* out, int64_t capacity, int64_t* n_forms);

/* Step 2.  Replaces FindCandidateSpectraAndScoreWorker.run (pg/FindCandidateSpectraAndScoreWorker.java:33-100):
 * every usable spectrum x every target form in its precursor window; keeps score > min_score. */
int32_t pq_step2_match_targets(pq_ctx* ctx, int64_t* n_psms);
/* Step 3.  Replaces DBSearchWorker.run (pg/DBSearchWorker.java:25-113) driven by DatabaseInput.db_search
 * (pg/DatabaseInput.java:672-758): n_db / total_db per matched spectrum, valid flags. */
int32_t pq_step3_competitors(pq_ctx* ctx);
/* Step 4.  Replaces StatisticScoreCalc.run (pg/StatisticScoreCalc.java:19-63) + StatisticScoreWorker.run
 * (pg/StatisticScoreWorker.java:36-63) + PeptideInput.generateRandomPeptidesFast / getModificationPeptide
 * (pg/PeptideInput.java:376-557). */
int32_t pq_step4_shuffles(pq_ctx* ctx);
/* Rank + accept.  pg/RankPSM.java:36-148 and CParameter.passed_psm_validation (pg/CParameter.java:913-940). */
int32_t pq_rank_and_validate(pq_ctx* ctx);
/* Step 5.  Replaces ModificationDB.doPTMValidation (OpenModificationSearch/ModificationDB.java:1320-1466)
 * -> ModPepQueryAndScoringWorker.do_score (ModPepQueryAndScoringWorker.java:149-276). */
int32_t pq_step5_ptm(pq_ctx* ctx, int32_t n_ptm, const pq_ptm* ptms);

int32_t pq_get_psms(pq_ctx* ctx, pq_psm* out, int64_t capacity, int64_t* n_psms);
/* Step-1 novelty filter (pg/InputProcessor.java:379-448 with the default FM-index mapping, pg/PepMapping.java:60-107):
 * nhit[i] = 1 when target sequence i occurs as a substring of some protein of the reference database, I and L
 * indistinguishable, case ignored; 0 otherwise.  Exact substring semantics (not digest membership).  Independent of the
 * search state of ctx (only its device and stream are used). */
int32_t pq_find_in_proteome(pq_ctx* ctx, int32_t n_seq, const uint32_t* seq_offset /* n+1 */, const char* residues,
                            int32_t n_proteins, const uint64_t* protein_offset /* n+1 */, const char* protein_residues, int32_t* nhit);

/* ref_delta_score / mod_delta_score of psm_rank.txt (PeptideSearchMT.add_delta_score, pg/PeptideSearchMT.java:1365-1475), one
 * pair per PSM in pq_get_psms order: score minus the best detail.txt (Step 3) / ptm.txt (Step 5) row of the same spectrum;
 * the plain score where the spectrum has no such row or the step has not run.  NULL outputs: only n_psms is written. */
int32_t pq_get_delta_scores(pq_ctx* ctx, double* ref_delta_score, double* mod_delta_score, int64_t capacity, int64_t* n_psms);
/* Multi-GPU gather: exports the PSM records (same layout and order as pq_get_psms) into a device buffer
 * owned by the context and returns its device address, so that the one collective of the path (NCCL gather of per-shard
 * PSM records) reads them without a host round trip.  The address stays valid until the next call on this context. */
int32_t pq_export_psms_device(pq_ctx* ctx, uint64_t* device_address, int64_t* n_psms);
int32_t pq_get_competitors(pq_ctx* ctx, int32_t step /* 3 or 5 */, pq_competitor* out, int64_t capacity, int64_t* n_rows);
/* Shuffled sequences of one target sequence, in generation order (for parity checks of the RNG stream). */
int32_t pq_get_shuffles(pq_ctx* ctx, int32_t target_seq, char* out /* n * (L+1) bytes, NUL separated */,
                        int64_t capacity, int32_t* n_shuffles);

/* Per-pair entry.  Replaces PeptideSearchMT.scorePeptide2Spectrum (pg/PeptideSearchMT.java:1144-1251) for
 * callers that score an explicit list (pg/ScoreWorker.java:26-45, plot/, msio/).  forms[i] is scored
 * against spectrum_index[i]; outputs score and the two counts. */
int32_t pq_score_pairs(pq_ctx* ctx, int64_t n_pairs, const int32_t* spectrum_index, const pq_form* forms,
                       double* score, int32_t* count_a, int32_t* count_b);

int32_t pq_get_counters(pq_ctx* ctx, pq_counters* out);
int32_t pq_reset_counters(pq_ctx* ctx);
/* Device-side stopwatch: CUDA events recorded on the context's own stream (torch.cuda.Event cannot see that stream).
 * which = 0 records the start event, 1 the stop event; elapsed = stop - start after synchronising the stop event. */
int32_t pq_event_record(pq_ctx* ctx, int32_t which);
int32_t pq_event_elapsed_ms(pq_ctx* ctx, double* ms);
/* Device self-test of an arithmetic shortcut of the kernels (diagnostic; no reference counterpart).
 * which = 1: the three-operation RN(x / 3) the kernels use for the m/z of triply charged fragments against the IEEE division, over
 * n doubles (pseudo-random significands and exponents from `seed`, magnitudes 2^-8 .. 2^16, plus their neighbours in ulps).
 * *mismatches = how many differed (must be 0). */
int32_t pq_selftest(pq_ctx* ctx, int32_t which, uint64_t n, uint64_t seed, uint64_t* mismatches);

/* Host-side text front end: single-pass MGF parser into CSR buffers.  Two-call pattern:
 * first call with NULL outputs returns the required sizes. */
int32_t pq_parse_mgf(const char* text, uint64_t n_bytes, int64_t* n_spectra, uint64_t* n_peaks,
                     double* precursor_mz, int32_t* charge, uint64_t* peak_offset, double* mz, double* intensity);

/* On-disk MS/MS library: a directory with one MGF per 0.1-Da bin of the neutral precursor mass,
 * <round(10 * mass)>.mgf, as index/BuildMSlibrary.java:173-352 + index/IndexWorker.java:422-486 write it and
 * msio/MsLibrarySearchWorker.java:42-153 reads it.  Returns the spectra with mass_lo <= mass <= mass_hi (the range of one GPU's
 * shard of a sweep) of the bins round(10 mass_lo) .. round(10 mass_hi), bin by bin, as CSR arrays for pq_load_spectra, plus their
 * TITLE texts (the identity of a library spectrum).  Spectra without a charge are ignored like the reference does.  Two-call
 * pattern like pq_parse_mgf; plain-text bins only (the reference also accepts .gz / .xz, decompressed by the Java host). */
int32_t pq_read_ms_library(const char* dir, double mass_lo, double mass_hi, int64_t* n_spectra, uint64_t* n_peaks, uint64_t* title_bytes,
                           double* precursor_mz, int32_t* charge, uint64_t* peak_offset, double* mz, double* intensity,
                           uint64_t* title_offset /* n+1 */, char* titles);

#ifdef __cplusplus
}
#endif
#endif /* PEPQUERY_B200_H */

// pq_host.h — host-side peptide logic of libpepquery_b200.so: modification codes, form expansion, digestion,
// shuffle group order.  Written independently of oracle/ (the oracle is only ever a test-side checker).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "pq_internal.h"

namespace pq {

struct HostMod {
    int32_t id, position;
    uint32_t residue_mask;   // bit (letter - 'A')
    double delta, loss;
};

// Compact form record (host): parent sequence, mass and the (modification, site) list in reference list order.
struct FormRec {
    int32_t seq;
    double mass;
    uint8_t n;
    uint8_t site[PQ_MAX_FORM_MODS];
    int8_t mod[PQ_MAX_FORM_MODS];      // index into Codes::mods (var first, then fixed)
};

struct Codes {
    std::vector<HostMod> mods;          // var mods then fixed mods
    int n_var = 0, n_fixed = 0;
    int max_var = 3, max_per_aa = 1;
    ModTable table;                     // device copy is made from this
    ShuffleMods shuffle;                // K2 view
    int n_codes = 26;
    double residue_mass[26];
    bool init(const pq_params& p, std::string& err);
    // sites a modification may occupy on a sequence, ascending (compomics ModificationUtils.getPossibleModificationSites, A9)
    void possible_sites(const std::string& seq, int mod, std::vector<int>& out) const;
    // PeptideInput.calcPeptideIsoforms + addFixedModification (pg/PeptideInput.java:114-237)
    void expand_forms(const std::string& seq, int32_t seq_index, std::vector<FormRec>& out) const;
    double form_mass(const std::string& seq, const FormRec& f) const;
    bool encode_row(const std::string& seq, const FormRec& f, uint8_t* row) const;   // 64-byte device row
    void to_public(const std::string& seq, const FormRec& f, pq_form* o) const;
    bool from_public(const pq_form& f, std::string& seq, FormRec& out) const;
    // inverse of encode_row: the modification list comes back in reference list order (variable mods by (mod, site), then fixed)
    void decode_row(const uint8_t* row, std::string& seq, FormRec& out) const;
};

// Cleavage rule of one enzyme as residue bit masks (bit = letter - 'A'): a cut between residues x | y happens when
// (x in cut_after and y not in block_next) or (y in cut_before).  pg/DatabaseInput.java:145-267 lists 17 enzymes
// (index = the -e option); compomics Enzyme.isCleavageSite applies the rule (A10).
struct EnzymeRule { uint32_t cut_after, cut_before, block_next; };
bool enzyme_rule(int32_t index, EnzymeRule& out);          // false: unknown index; 0 = NoEnzyme (unspecific digestion: every residue ends a piece)

// The reference table as K6 builds it on the device (k6_reference.cu).  rows/mass point at the table's DevBufs.
struct RefDeviceTable {
    DevBuf residues, starts, cand_pos, cand_len, uniq, seq_id;   // normalised proteins, segment starts, candidate (pos,len), unique candidate ids (lexicographic), parent id per form
    DevBuf* rows = nullptr; DevBuf* mass = nullptr;
    DevBuf w[26];                       // work buffers of the build, kept so that a rebuild allocates nothing
    uint32_t n_unique = 0, n_forms = 0, n_raw = 0, n_tkeys = 0; int32_t n_proteins = 0;
};
// the three parts of build_reference_on_device (k6_reference.cu): upload, digest (independent of the spectra), expansion
bool upload_target_keys_for_reference(cudaStream_t st, const std::vector<std::string>& targets, RefDeviceTable& T, std::string& err);
bool upload_proteins_for_reference(cudaStream_t st, int32_t n_proteins, const uint64_t* off, const char* res, RefDeviceTable& T, std::string& err);
bool digest_reference_on_device(cudaStream_t st, const pq_params& P, const Codes& codes, const std::vector<std::string>& targets, RefDeviceTable& T,
                                uint64_t* launches, std::string& err);
bool expand_reference_on_device(cudaStream_t st, const pq_params& P, const Codes& codes, double lo, double hi, RefDeviceTable& T, uint64_t* launches,
                                std::string& err);
bool build_reference_on_device(cudaStream_t st, const pq_params& P, const Codes& codes, int32_t n_proteins, const uint64_t* off, const char* res,
                               const std::vector<std::string>& targets, double lo, double hi, RefDeviceTable& T, uint64_t* launches, std::string& err);

// The target table as K6 builds it on the device; the pointers name the context's buffers.
struct TargetDeviceTable {
    DevBuf* rows = nullptr; DevBuf* mass = nullptr; DevBuf* form_of_row = nullptr; DevBuf* seq_of_form = nullptr; DevBuf* form_mods = nullptr;
    DevBuf* seq_off = nullptr; DevBuf* seq_letters = nullptr;
    DevBuf w[13];
    uint32_t n_forms = 0;
};
bool build_target_table_on_device(cudaStream_t st, const Codes& codes, uint32_t n_seq, const uint32_t* off, const char* flat, TargetDeviceTable& T,
                                  uint64_t* launches, std::string& err);

// MGF text parsed on the device (pq_mgf.cu): work buffers kept in the context so that a second file allocates nothing
struct MgfDeviceWork { DevBuf text, line_pos, type, delta, incl, pflag, pidx, hard, tmp, scal; std::vector<cudaEvent_t> ev; };
struct ByteToU64 { __host__ __device__ uint64_t operator()(uint8_t b) const { return b; } };
bool parse_mgf_on_device(cudaStream_t st, cudaStream_t sc, const char* text, uint64_t n_bytes, MgfDeviceWork& W, DevBuf& d_prec, DevBuf& d_charge,
                         DevBuf& d_off, DevBuf& d_mz, DevBuf& d_in, int64_t* n_spectra, uint64_t* n_peaks, uint64_t* launches, std::string& err);

// commons-math 2 MathUtils tables (A12): log10(n!) for n <= 64 and ln(n!) for n < count
void factorial_tables(double* log10_fact65, std::vector<double>& ln_fact, int count);

}  // namespace pq

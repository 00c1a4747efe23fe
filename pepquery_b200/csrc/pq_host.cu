// pq_host.cu — host-side peptide logic (see pq_host.h).  Product code; shares nothing with oracle/.
#include "pq_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_set>

namespace pq {

static void set_residue_masses(double* m) {
    for (int i = 0; i < 26; ++i) m[i] = 0.0;
    auto S = [&](char c, double v) { m[c - 'A'] = v; };
    // compomics AminoAcid monoisotopic masses (atoms pinned by resources/top_modifications.tsv)
    S('G', 57.02146372057);  S('A', 71.03711378471);  S('S', 87.03202840427);  S('P', 97.05276384885);
    S('V', 99.06841391299);  S('T', 101.04767846841); S('C', 103.00918478471); S('L', 113.08406397713);
    S('I', 113.08406397713); S('N', 114.04292744114); S('D', 115.02694302383); S('Q', 128.05857750528);
    S('K', 128.094963014);   S('E', 129.04259308797); S('M', 131.04048491299); S('H', 137.05891185845);
    S('F', 147.06841391299); S('R', 156.1011110236);  S('Y', 163.06332853255); S('W', 186.07931294986);
}

bool Codes::init(const pq_params& p, std::string& err) {
    set_residue_masses(residue_mass);
    n_var = p.n_var; n_fixed = p.n_fixed; max_var = p.max_var_mods; max_per_aa = p.max_mods_per_aa;
    if (n_var < 0 || n_var > PQ_MAX_MODS || n_fixed < 0 || n_fixed > PQ_MAX_MODS) { err = "too many modifications"; return false; }
    if (max_per_aa != 1) { err = "max_mods_per_aa != 1 is not supported on the device path"; return false; }
    memset(&table, 0, sizeof table);
    memset(&shuffle, 0, sizeof shuffle);
    for (int c = 0; c < 26; ++c) table.aa[c] = residue_mass[c];
    n_codes = 26;
    int n_term = 1;
    mods.clear();
    for (int i = 0; i < n_var + n_fixed; ++i) {
        const pq_mod& m = i < n_var ? p.var[i] : p.fixed[i - n_var];
        HostMod h; h.id = m.id; h.position = m.position; h.delta = m.delta; h.loss = m.neutral_loss; h.residue_mask = 0;
        for (int k = 0; k < 24 && m.residues[k]; ++k) {
            char c = m.residues[k];
            if (c < 'A' || c > 'Z') { err = "modification residue must be A-Z"; return false; }
            h.residue_mask |= 1u << (c - 'A');
        }
        bool residue_specific = m.position == PQ_MOD_ANYWHERE || m.position == PQ_MOD_PEP_NTERM_AA || m.position == PQ_MOD_PEP_CTERM_AA ||
                                m.position == PQ_MOD_PROT_NTERM_AA || m.position == PQ_MOD_PROT_CTERM_AA;
        shuffle.position[i] = m.position; shuffle.residue_mask[i] = h.residue_mask;
        if (residue_specific) {
            if (!h.residue_mask) { err = "residue-specific modification without residues"; return false; }
            for (int r = 0; r < 26; ++r) if ((h.residue_mask >> r) & 1u) {
                if (n_codes >= kMaxCodes) { err = "too many (residue, modification) codes"; return false; }
                table.aa[n_codes] = residue_mass[r]; table.delta[n_codes] = m.delta; table.loss[n_codes] = m.neutral_loss;
                shuffle.code_of[i][r] = (uint8_t)n_codes;
                ++n_codes;
            }
        } else {
            if (n_term >= kMaxTermCodes) { err = "too many terminal modifications"; return false; }
            table.nterm[n_term] = m.delta; table.cterm[n_term] = m.delta;
            table.nterm_loss[n_term] = m.neutral_loss; table.cterm_loss[n_term] = m.neutral_loss;
            shuffle.term_code[i] = (uint8_t)n_term;
            ++n_term;
        }
        mods.push_back(h);
    }
    shuffle.n_mods = n_var + n_fixed; shuffle.n_var = n_var;
    shuffle.shuffle_seed = p.shuffle_seed; shuffle.mod_seed = p.mod_seed;
    return true;
}

void Codes::possible_sites(const std::string& seq, int mod, std::vector<int>& out) const {
    out.clear();
    const HostMod& m = mods[mod];
    const int L = (int)seq.size();
    auto has = [&](char c) { return c >= 'A' && c <= 'Z' && ((m.residue_mask >> (c - 'A')) & 1u); };
    switch (m.position) {
        case PQ_MOD_ANYWHERE: for (int i = 0; i < L; ++i) if (has(seq[i])) out.push_back(i + 1); break;
        case PQ_MOD_PEP_NTERM: case PQ_MOD_PROT_NTERM: out.push_back(0); break;
        case PQ_MOD_PEP_CTERM: case PQ_MOD_PROT_CTERM: out.push_back(L + 1); break;
        case PQ_MOD_PEP_NTERM_AA: case PQ_MOD_PROT_NTERM_AA: if (L && has(seq[0])) out.push_back(1); break;
        case PQ_MOD_PEP_CTERM_AA: case PQ_MOD_PROT_CTERM_AA: if (L && has(seq[L - 1])) out.push_back(L); break;
        default: break;
    }
}

double Codes::form_mass(const std::string& seq, const FormRec& f) const {
    // Peptide.getMass via JPeptide.getMass (PSMMatch/JPeptide.java:73-78): H2O, residues in order, modifications in list order
    double m = kWater;
    for (char c : seq) m += residue_mass[c - 'A'];
    for (int i = 0; i < f.n; ++i) m += mods[f.mod[i]].delta;
    return m;
}

void Codes::expand_forms(const std::string& seq, int32_t seq_index, std::vector<FormRec>& out) const {
    struct Slot { int mod, site; };
    std::vector<Slot> slots;
    std::vector<int> tmp;
    for (int v = 0; v < n_var; ++v) { possible_sites(seq, v, tmp); for (int s : tmp) slots.push_back({v, s}); }
    // fixed-modification sites, in fixed-mod order
    std::vector<Slot> fixed_slots;
    for (int f = n_var; f < n_var + n_fixed; ++f) { possible_sites(seq, f, tmp); for (int s : tmp) fixed_slots.push_back({f, s}); }
    auto emit = [&](const int* pick, int k) {
        FormRec r; memset(&r, 0, sizeof r);
        r.seq = seq_index;
        uint64_t taken = 0ull;
        for (int i = 0; i < k; ++i) {
            const Slot& s = slots[pick[i]];
            if ((taken >> s.site) & 1ull) return;                // more than max_mods_per_aa (=1) on a site
            taken |= 1ull << s.site;
            if (r.n >= PQ_MAX_FORM_MODS) return;
            r.site[r.n] = (uint8_t)s.site; r.mod[r.n] = (int8_t)s.mod; ++r.n;
        }
        for (const Slot& s : fixed_slots) {
            if ((taken >> s.site) & 1ull) continue;               // only the variable sites block a fixed mod (:218-233)
            if (r.n >= PQ_MAX_FORM_MODS) break;
            r.site[r.n] = (uint8_t)s.site; r.mod[r.n] = (int8_t)s.mod; ++r.n;
        }
        r.mass = form_mass(seq, r);
        out.push_back(r);
    };
    const int n = (int)slots.size();
    const int kmax = std::min(n, max_var);
    int pick[PQ_MAX_FORM_MODS];
    for (int k = 1; k <= kmax && k <= PQ_MAX_FORM_MODS; ++k) {
        for (int i = 0; i < k; ++i) pick[i] = i;                  // lexicographic k-subsets (combinatoricslib3 simple(k))
        for (;;) {
            emit(pick, k);
            int i = k - 1;
            while (i >= 0 && pick[i] == n - k + i) --i;
            if (i < 0) break;
            ++pick[i];
            for (int j = i + 1; j < k; ++j) pick[j] = pick[j - 1] + 1;
        }
    }
    emit(pick, 0);                                                // the fixed-only form comes last (:177-179)
}

bool Codes::encode_row(const std::string& seq, const FormRec& f, uint8_t* row) const {
    const int L = (int)seq.size();
    if (L < 1 || L > kMaxLen) return false;
    memset(row, 0, kRowBytes);
    row[0] = (uint8_t)L;
    for (int i = 0; i < L; ++i) { char c = seq[i]; if (c < 'A' || c > 'Z') return false; row[kRowHeader + i] = (uint8_t)(c - 'A'); }
    for (int i = 0; i < f.n; ++i) {
        int m = f.mod[i], s = f.site[i];
        if (s == 0) { if (row[1]) return false; row[1] = shuffle.term_code[m]; }
        else if (s == L + 1) { if (row[2]) return false; row[2] = shuffle.term_code[m]; }
        else {
            uint8_t& b = row[kRowHeader + s - 1];
            if (b >= 26) return false;                            // two modifications on one residue
            uint8_t code = shuffle.code_of[m][b];
            if (!code) return false;
            b = code;
        }
    }
    return true;
}

void Codes::to_public(const std::string& seq, const FormRec& f, pq_form* o) const {
    memset(o, 0, sizeof *o);
    o->seq_index = f.seq; o->length = (int32_t)seq.size(); o->mass = f.mass; o->n_mods = f.n;
    for (int i = 0; i < f.n; ++i) { o->mod_site[i] = f.site[i]; o->mod_index[i] = (int8_t)(f.mod[i] < n_var ? f.mod[i] : -(f.mod[i] - n_var) - 1); }
    strncpy(o->seq, seq.c_str(), PQ_MAX_PEP_LEN);
}
bool Codes::from_public(const pq_form& p, std::string& seq, FormRec& out) const {
    seq.assign(p.seq, strnlen(p.seq, PQ_MAX_PEP_LEN));
    memset(&out, 0, sizeof out);
    out.seq = p.seq_index;
    if (p.n_mods < 0 || p.n_mods > PQ_MAX_FORM_MODS) return false;
    out.n = (uint8_t)p.n_mods;
    for (int i = 0; i < p.n_mods; ++i) {
        int mi = p.mod_index[i] >= 0 ? p.mod_index[i] : n_var + (-(int)p.mod_index[i] - 1);
        if (mi < 0 || mi >= n_var + n_fixed) return false;
        out.site[i] = p.mod_site[i]; out.mod[i] = (int8_t)mi;
    }
    for (char c : seq) if (c < 'A' || c > 'Z') return false;
    out.mass = form_mass(seq, out);
    return true;
}

void Codes::decode_row(const uint8_t* row, std::string& seq, FormRec& out) const {
    const int L = row[0];
    seq.resize((size_t)L);
    memset(&out, 0, sizeof out);
    int8_t mod_at[kMaxLen + 2];
    for (int i = 0; i < kMaxLen + 2; ++i) mod_at[i] = -1;
    const int nm = n_var + n_fixed;
    for (int i = 0; i < L; ++i) {
        uint8_t c = row[kRowHeader + i];
        if (c < 26) { seq[(size_t)i] = (char)('A' + c); continue; }
        for (int m = 0; m < nm && mod_at[i + 1] < 0; ++m)
            for (int r = 0; r < 26; ++r) if (shuffle.code_of[m][r] == c) { seq[(size_t)i] = (char)('A' + r); mod_at[i + 1] = (int8_t)m; break; }
    }
    for (int m = 0; m < nm; ++m) {
        if (!shuffle.term_code[m]) continue;
        if (row[1] == shuffle.term_code[m] && (shuffle.position[m] == PQ_MOD_PEP_NTERM || shuffle.position[m] == PQ_MOD_PROT_NTERM)) mod_at[0] = (int8_t)m;
        if (row[2] == shuffle.term_code[m] && (shuffle.position[m] == PQ_MOD_PEP_CTERM || shuffle.position[m] == PQ_MOD_PROT_CTERM)) mod_at[L + 1] = (int8_t)m;
    }
    for (int m = 0; m < nm; ++m)
        for (int s = 0; s <= L + 1; ++s)
            if (mod_at[s] == m && out.n < PQ_MAX_FORM_MODS) { out.site[out.n] = (uint8_t)s; out.mod[out.n] = (int8_t)m; ++out.n; }
    out.mass = form_mass(seq, out);
}

bool enzyme_rule(int32_t index, EnzymeRule& out) {
    struct Def { const char* after; const char* before; const char* block; };
    static const Def defs[18] = {
        {"ABCDEFGHIJKLMNOPQRSTUVWXYZ", "", ""},      // 0 NoEnzyme: unspecific digestion, every residue ends a piece (K6 then joins up to max_len of them)
        {"RK", "", "P"}, {"RK", "", ""}, {"R", "", "P"}, {"R", "", ""}, {"", "R", ""}, {"E", "", ""}, {"K", "", "P"}, {"K", "", ""},
        {"", "K", ""}, {"", "D", ""}, {"", "DE", ""}, {"FYWL", "", "P"}, {"FYWL", "", ""}, {"FL", "", ""}, {"M", "", ""},
        {"", "AFILMV", ""}, {"", "RK", ""}};
    if (index < 0 || index > 17) return false;
    auto mask = [](const char* s) { uint32_t m = 0; for (; *s; ++s) m |= 1u << (*s - 'A'); return m; };
    out.cut_after = mask(defs[index].after); out.cut_before = mask(defs[index].before); out.block_next = mask(defs[index].block);
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
void factorial_tables(double* log10_fact65, std::vector<double>& ln_fact, int count) {
    if (count < 65) count = 65;
    ln_fact.assign((size_t)count, 0.0);
    uint64_t f = 1;
    for (int i = 0; i < 21 && i < count; ++i) { if (i > 0) f *= (uint64_t)i; ln_fact[(size_t)i] = jlog((double)f); }
    double s = 0.0;
    for (int i = 2; i < count; ++i) { s += jlog((double)i); if (i >= 21) ln_fact[(size_t)i] = s; }
    log10_fact65[0] = 0.0;
    f = 1;
    for (int i = 1; i <= 64; ++i) {
        double fd;
        if (i < 21) { f *= (uint64_t)i; fd = (double)f; }
        else fd = std::floor(std::exp(ln_fact[(size_t)i]) + 0.5);       // MathUtils.factorialDouble (A12)
        log10_fact65[i] = jlog10(1.0 * fd);
    }
}

}  // namespace pq

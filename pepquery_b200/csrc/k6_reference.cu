// k6_reference.cu — K6: the reference peptide table built ON THE DEVICE (north_star item 2).
//
// Replaces DatabaseInput.protein_digest (pg/DatabaseInput.java:402-478) + DigestProteinWorker.run
// (OpenModificationSearch/DigestProteinWorker.java:64-137) + generate_reference_peptide_index (pg/DatabaseInput.java:490-556)
// + GeneratePeptideformWorker.run (OpenModificationSearch/GeneratePeptideformWorker.java:25-52) +
// PeptideInput.calcPeptideIsoforms / addFixedModification (pg/PeptideInput.java:114-237):
//
//   residues (upper-cased, '*' stripped, I->L on the host: one linear pass)            -> H2D
//   cut flags      : a segment starts at a protein start or after K/R not followed by P (A10)
//   candidates     : segment a joined with 0..max_missed following segments of the same protein, length / residue /
//                    mass (H2O + residues, summed in order) filters                       (DigestProteinWorker.java:79-104)
//   unique         : 5-bit-packed 300-bit keys (60 residues), LSD radix sort (cub) over the words in use = lexicographic order, run heads, minus the targets
//   forms          : per unique sequence, all k-subsets (k <= max_var) of the variable (mod, site) slots in lexicographic
//                    order, fixed modifications on the remaining sites, fixed-only form last; kept when
//                    lo < mass < hi (the matched-spectra window +-255 Da)
//   table          : stable radix sort by mass -> 64-byte rows + masses + parent-sequence ids, resident in HBM
//
// The host never sees the sequences unless pq_get_reference_forms asks for them.
#include <algorithm>
#include <cstring>

#include <cub/cub.cuh>

#include "pq_host.h"

namespace pq {

namespace {

constexpr int kKeyWords = 5;            // 60 residues x 5 bits (kMaxLen): twelve residues per 64-bit word
constexpr int kMaxSlots = 64;

struct ByteToU32 { __host__ __device__ uint32_t operator()(uint8_t b) const { return b; } };

// DigestProteinWorker.java:66-70 on the device: upper-case, I -> L, '*' dropped (keep flag for the compaction)
__global__ void k6_normalise(const uint8_t* __restrict__ raw, uint32_t n, uint8_t* __restrict__ norm, uint8_t* __restrict__ keep) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t c = raw[i];
    if (c >= 'a' && c <= 'z') c = (uint8_t)(c - 32);
    if (c == 'I') c = 'L';
    norm[i] = c; keep[i] = c != '*';
}
__global__ void k6_compact(const uint8_t* __restrict__ norm, const uint8_t* __restrict__ keep, const uint32_t* __restrict__ pos, uint32_t n,
                           uint8_t* __restrict__ out, uint8_t* __restrict__ first) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (keep[i]) { out[pos[i]] = norm[i]; first[pos[i]] = 0; }
}
// a protein starts at the first kept residue at or after its raw offset
__global__ void k6_protein_starts(const uint64_t* __restrict__ off, int32_t n_prot, const uint32_t* __restrict__ pos, uint32_t n_raw, uint32_t n_kept,
                                  uint8_t* __restrict__ first) {
    int32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_prot) return;
    uint64_t a = off[p], b = off[p + 1];
    uint32_t pa = a < n_raw ? pos[a] : n_kept, pb = b < n_raw ? pos[b] : n_kept;
    if (pa < pb) first[pa] = 1;
}

__global__ void k6_cut_flags(const uint8_t* __restrict__ res, const uint8_t* __restrict__ first, uint32_t n, EnzymeRule rule, uint8_t* __restrict__ is_start) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t f = first[i];
    if (!f && i > 0) {                                      // compomics Enzyme.isCleavageSite(x, y) (A10)
        const uint8_t x = res[i - 1], y = res[i];
        if (x >= 'A' && x <= 'Z' && y >= 'A' && y <= 'Z')
            f = ((((rule.cut_after >> (x - 'A')) & 1u) && !((rule.block_next >> (y - 'A')) & 1u)) || ((rule.cut_before >> (y - 'A')) & 1u)) ? 1 : 0;
    }
    is_start[i] = f;
}

struct DigestCfg { int32_t max_missed, min_len, max_len; double min_mass, max_mass; double residue_mass[26]; };

__global__ void k6_candidate_flags(const uint32_t* __restrict__ starts, uint32_t n_seg, const uint8_t* __restrict__ first,
                                   const uint8_t* __restrict__ res, DigestCfg cfg, uint8_t* __restrict__ flags) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t per = (uint32_t)cfg.max_missed + 1;
    if (t >= n_seg * per) return;
    uint32_t a = t / per, mc = t % per;
    uint8_t ok = 0;
    if (a + mc + 1 <= n_seg) {
        bool same = true;
        for (uint32_t j = 1; j <= mc; ++j) if (first[starts[a + j]]) { same = false; break; }
        uint32_t b = starts[a], e = starts[a + mc + 1];
        int len = (int)(e - b);
        if (same && len >= cfg.min_len && len <= cfg.max_len) {
            double mass = kWater; bool good = true;
            for (uint32_t i = b; i < e; ++i) {
                uint8_t c = res[i];
                if (c < 'A' || c > 'Z') { good = false; break; }
                double m = cfg.residue_mass[c - 'A'];
                if (m == 0.0) { good = false; break; }          // X and other ambiguity codes (DigestProteinWorker.java:103)
                mass = __dadd_rn(mass, m);
            }
            ok = (good && mass >= cfg.min_mass && mass <= cfg.max_mass) ? 1 : 0;
        }
    }
    flags[t] = ok;
}

__global__ void k6_pack_keys(const uint32_t* __restrict__ cand, uint32_t n_cand, const uint32_t* __restrict__ starts, uint32_t per,
                             const uint8_t* __restrict__ res, uint64_t* __restrict__ keys /* [4][n_cand] */, uint32_t* __restrict__ pos,
                             uint32_t* __restrict__ len_out, uint32_t* __restrict__ perm) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cand) return;
    uint32_t t = cand[i], a = t / per, mc = t % per;
    uint32_t b = starts[a], e = starts[a + mc + 1];
    uint64_t k[kKeyWords] = {0, 0, 0, 0, 0};
    for (uint32_t j = b; j < e; ++j) {
        uint32_t r = j - b;
        k[r / 12] |= (uint64_t)(res[j] - 'A' + 1) << (59 - 5 * (r % 12));
    }
    for (int w = 0; w < kKeyWords; ++w) keys[(size_t)w * n_cand + i] = k[w];
    pos[i] = b; len_out[i] = e - b; perm[i] = i;
}

__global__ void k6_gather_u64(const uint64_t* __restrict__ src, const uint32_t* __restrict__ perm, uint32_t n, uint64_t* __restrict__ dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}

// After ONE radix sort on the first key word (the first twelve residues): candidates that share it — mostly the same start
// with more missed cleavages behind a long first piece — form short runs; the thread of a run's head puts the run in order of the
// remaining words with a stable insertion sort, which leaves exactly the order the word-by-word LSD sort gives (equal keys keep
// candidate order).  A run longer than kMaxRun raises *overflow: the caller then sorts word by word.
constexpr uint32_t kMaxRun = 48;
__device__ __forceinline__ bool tail_less(const uint64_t* __restrict__ keys, uint32_t n, int words, uint32_t a, uint32_t b) {      // words 1.. of candidate a < those of b
    for (int w = 1; w < words; ++w) {
        const uint64_t x = keys[(size_t)w * n + a], y = keys[(size_t)w * n + b];
        if (x != y) return x < y;
    }
    return false;
}
__global__ void k6_order_runs(const uint64_t* __restrict__ word0_sorted, uint32_t* __restrict__ perm, const uint64_t* __restrict__ keys, uint32_t n, int words,
                              uint32_t* __restrict__ overflow) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k0 = word0_sorted[i];
    if (i > 0 && word0_sorted[i - 1] == k0) return;                 // not a head
    if (i + 1 >= n || word0_sorted[i + 1] != k0) return;            // a run of one
    uint32_t len = 2;
    while (i + len < n && word0_sorted[i + len] == k0) { if (++len > kMaxRun) { *overflow = 1u; return; } }
    for (uint32_t a = 1; a < len; ++a) {                            // stable insertion sort of perm[i .. i+len)
        const uint32_t c = perm[i + a];
        uint32_t b = a;
        while (b > 0 && tail_less(keys, n, words, c, perm[i + b - 1])) { perm[i + b] = perm[i + b - 1]; --b; }
        perm[i + b] = c;
    }
}

__device__ __forceinline__ int key_cmp(const uint64_t* keys, uint32_t n, uint32_t i, const uint64_t* tk) {
    for (int w = 0; w < kKeyWords; ++w) {
        uint64_t a = keys[(size_t)w * n + i], b = tk[w];
        if (a != b) return a < b ? -1 : 1;
    }
    return 0;
}

__global__ void k6_unique_flags(const uint32_t* __restrict__ perm, const uint64_t* __restrict__ keys, uint32_t n,
                                const uint64_t* __restrict__ tkeys /* [n_t][4] sorted */, uint32_t n_t, uint8_t* __restrict__ flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = perm[i];
    bool head = true;
    if (i > 0) {
        uint32_t p = perm[i - 1];
        head = false;
        for (int w = 0; w < kKeyWords; ++w) if (keys[(size_t)w * n + c] != keys[(size_t)w * n + p]) { head = true; break; }
    }
    if (head && n_t) {      // DatabaseInput.protein_digest(db, target_peptides) drops the target sequences (:480-488)
        uint32_t lo = 0, hi = n_t;
        while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (key_cmp(keys, n, c, tkeys + (size_t)mid * kKeyWords) > 0) lo = mid + 1; else hi = mid; }
        if (lo < n_t && key_cmp(keys, n, c, tkeys + (size_t)lo * kKeyWords) == 0) head = false;
    }
    flags[i] = head ? 1 : 0;
}

struct ExpandCfg {
    int32_t n_var, n_fixed, max_var;
    int32_t position[2 * PQ_MAX_MODS];
    uint32_t residue_mask[2 * PQ_MAX_MODS];
    double delta[2 * PQ_MAX_MODS];
    uint8_t code_of[2 * PQ_MAX_MODS][26];
    uint8_t term_code[2 * PQ_MAX_MODS];
    double residue_mass[26];
    double lo, hi;
};

__device__ __forceinline__ uint64_t sites_of(const ExpandCfg& C, int m, const uint8_t* s, int L) {
    uint64_t r = 0ull; uint32_t mask = C.residue_mask[m];
    switch (C.position[m]) {
        case PQ_MOD_ANYWHERE: for (int i = 0; i < L; ++i) if ((mask >> (s[i] - 'A')) & 1u) r |= 1ull << (i + 1); break;
        case PQ_MOD_PEP_NTERM: case PQ_MOD_PROT_NTERM: r = 1ull; break;
        case PQ_MOD_PEP_CTERM: case PQ_MOD_PROT_CTERM: r = 1ull << (L + 1); break;
        case PQ_MOD_PEP_NTERM_AA: case PQ_MOD_PROT_NTERM_AA: if ((mask >> (s[0] - 'A')) & 1u) r = 1ull << 1; break;
        case PQ_MOD_PEP_CTERM_AA: case PQ_MOD_PROT_CTERM_AA: if ((mask >> (s[L - 1] - 'A')) & 1u) r = 1ull << L; break;
        default: break;
    }
    return r;
}

// FILL = false: count the forms of each unique sequence inside the mass window; FILL = true: write them.
__global__ void k6_set_u32(uint32_t* p, uint32_t v) { *p = v; }

template <bool FILL>
__global__ void k6_expand(const uint32_t* __restrict__ uniq, uint32_t n_uniq, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ len,
                          const uint8_t* __restrict__ res, ExpandCfg C, uint32_t* __restrict__ count, const uint32_t* __restrict__ form_off,
                          uint8_t* __restrict__ rows, double* __restrict__ mass_out, uint32_t* __restrict__ seq_id, FormMods* __restrict__ fm_out = nullptr) {
    uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_uniq) return;
    const uint32_t c = uniq[u];
    const int L = (int)len[c];
    const uint8_t* s = res + pos[c];
    uint8_t slot_mod[kMaxSlots], slot_site[kMaxSlots];
    int n = 0;
    for (int v = 0; v < C.n_var; ++v) {
        uint64_t m = sites_of(C, v, s, L);
        while (m && n < kMaxSlots) { int site = __ffsll((long long)m) - 1; m &= m - 1; slot_mod[n] = (uint8_t)v; slot_site[n] = (uint8_t)site; ++n; }
    }
    double base = kWater;
    for (int i = 0; i < L; ++i) base = __dadd_rn(base, C.residue_mass[s[i] - 'A']);
    uint32_t emitted = 0;
    const uint32_t out0 = FILL ? form_off[u] : 0u;
    int pick[PQ_MAX_FORM_MODS];
    const int kmax = min(min(n, C.max_var), PQ_MAX_FORM_MODS);
    for (int k = 0; k <= kmax; ++k) {
        // order of the reference: k = 1..kmax subsets first, the fixed-only form (k = 0) last
        const int kk = (k < kmax) ? k + 1 : 0;
        for (int i = 0; i < kk; ++i) pick[i] = i;
        for (;;) {
            uint64_t taken = 0ull; bool clash = false;
            double mass = base;
            for (int i = 0; i < kk; ++i) {
                uint64_t bit = 1ull << slot_site[pick[i]];
                if (taken & bit) { clash = true; break; }
                taken |= bit;
                mass = __dadd_rn(mass, C.delta[slot_mod[pick[i]]]);
            }
            if (!clash) {
                uint64_t fixed_sites[PQ_MAX_MODS];
                for (int f = 0; f < C.n_fixed; ++f) {
                    uint64_t m = sites_of(C, C.n_var + f, s, L) & ~taken;       // only variable sites block a fixed mod
                    fixed_sites[f] = m;
                    for (int q = __popcll(m); q > 0; --q) mass = __dadd_rn(mass, C.delta[C.n_var + f]);
                }
                if (mass > C.lo && mass < C.hi) {
                    if (FILL) {
                        uint32_t w[kRowBytes / 4];
                        uint8_t* row = reinterpret_cast<uint8_t*>(w);
                        for (int i = 0; i < kRowBytes / 4; ++i) w[i] = 0u;
                        row[0] = (uint8_t)L;
                        for (int i = 0; i < L; ++i) row[kRowHeader + i] = (uint8_t)(s[i] - 'A');
                        auto apply = [&](int m, int site) {
                            if (site == 0) row[1] = C.term_code[m];
                            else if (site == L + 1) row[2] = C.term_code[m];
                            else row[kRowHeader + site - 1] = C.code_of[m][s[site - 1] - 'A'];
                        };
                        for (int i = 0; i < kk; ++i) apply(slot_mod[pick[i]], slot_site[pick[i]]);
                        for (int f = 0; f < C.n_fixed; ++f) { uint64_t m = fixed_sites[f]; while (m) { int site = __ffsll((long long)m) - 1; m &= m - 1; apply(C.n_var + f, site); } }
                        uint4* dst = reinterpret_cast<uint4*>(rows + (size_t)(out0 + emitted) * kRowBytes);
                        dst[0] = make_uint4(w[0], w[1], w[2], w[3]); dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                        dst[2] = make_uint4(w[8], w[9], w[10], w[11]); dst[3] = make_uint4(w[12], w[13], w[14], w[15]);
                        mass_out[out0 + emitted] = mass; seq_id[out0 + emitted] = u;
                        if (fm_out) {        // the form's modification list in the reference's order: the variable picks, then the fixed mods (PeptideInput.java:141-179, 213-237)
                            FormMods fm; fm.n_mods = 0;
                            for (int i = 0; i < PQ_MAX_FORM_MODS; ++i) { fm.mod[i] = 0; fm.site[i] = 0; }
                            for (int i = 0; i < kk; ++i) { fm.mod[fm.n_mods] = (int8_t)slot_mod[pick[i]]; fm.site[fm.n_mods] = slot_site[pick[i]]; ++fm.n_mods; }
                            for (int f = 0; f < C.n_fixed; ++f) {
                                uint64_t m = fixed_sites[f];
                                while (m && fm.n_mods < PQ_MAX_FORM_MODS) { int site = __ffsll((long long)m) - 1; m &= m - 1; fm.mod[fm.n_mods] = (int8_t)(C.n_var + f); fm.site[fm.n_mods] = (uint8_t)site; ++fm.n_mods; }
                            }
                            fm_out[out0 + emitted] = fm;
                        }
                    }
                    ++emitted;
                }
            }
            if (kk == 0) break;
            int i = kk - 1;
            while (i >= 0 && pick[i] == n - kk + i) --i;
            if (i < 0) break;
            ++pick[i];
            for (int j = i + 1; j < kk; ++j) pick[j] = pick[j - 1] + 1;
        }
    }
    if (!FILL) count[u] = emitted;
}

__global__ void k6_mass_keys(const double* __restrict__ mass, uint32_t n, uint64_t* __restrict__ key, uint32_t* __restrict__ idx) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = (uint64_t)__double_as_longlong(mass[i]);      // positive doubles order like their bit patterns
    idx[i] = i;
}
__global__ void k6_gather_rows(const uint32_t* __restrict__ order, uint32_t n, const uint8_t* __restrict__ rows_in, const double* __restrict__ mass_in,
                               const uint32_t* __restrict__ seq_in, uint8_t* __restrict__ rows, double* __restrict__ mass, uint32_t* __restrict__ seq) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t i = t >> 2, q = t & 3;
    if (i >= n) return;
    uint32_t s = order[i];
    reinterpret_cast<uint4*>(rows)[(size_t)i * 4 + q] = reinterpret_cast<const uint4*>(rows_in)[(size_t)s * 4 + q];
    if (q == 0) { mass[i] = mass_in[s]; seq[i] = seq_in[s]; }
}

#define RB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return false; } } while (0)

template <class F> bool select_flagged(const uint8_t* flags, uint32_t n, uint32_t* out, uint32_t* d_count, DevBuf& tmp, cudaStream_t st, std::string& err, F) {
    size_t bytes = 0;
    cub::CountingInputIterator<uint32_t> iota(0);
    cub::DeviceSelect::Flagged(nullptr, bytes, iota, flags, out, d_count, (int)n, st);
    RB(tmp.reserve(bytes + 64));
    RB(cub::DeviceSelect::Flagged(tmp.p, bytes, iota, flags, out, d_count, (int)n, st));
    return true;
}

}  // namespace

// Work buffers of the build (RefDeviceTable::w), by role
#define K6_BUFFERS(T) \
    DevBuf& d_first = T.w[0]; DevBuf& d_start = T.w[1]; DevBuf& d_flags = T.w[2]; DevBuf& d_cnt = T.w[3]; DevBuf& d_cand = T.w[4]; DevBuf& d_keys = T.w[5]; \
    DevBuf& d_tmpkey = T.w[6]; DevBuf& d_tmpkey2 = T.w[7]; DevBuf& d_perm = T.w[8]; DevBuf& d_perm2 = T.w[9]; DevBuf& d_tk = T.w[10]; DevBuf& d_count = T.w[11]; \
    DevBuf& d_foff = T.w[12]; DevBuf& d_rows = T.w[13]; DevBuf& d_mass = T.w[14]; DevBuf& d_seq = T.w[15]; DevBuf& d_mkey = T.w[16]; DevBuf& d_mkey2 = T.w[17]; \
    DevBuf& d_idx = T.w[18]; DevBuf& d_idx2 = T.w[19]; DevBuf& tmp = T.w[20]; DevBuf& d_raw = T.w[21]; DevBuf& d_norm = T.w[22]; DevBuf& d_keep = T.w[23]; \
    DevBuf& d_pos = T.w[24]; DevBuf& d_off = T.w[25]; \
    (void)d_first; (void)d_start; (void)d_flags; (void)d_cnt; (void)d_cand; (void)d_keys; (void)d_tmpkey; (void)d_tmpkey2; (void)d_perm; (void)d_perm2; (void)d_tk; \
    (void)d_count; (void)d_foff; (void)d_rows; (void)d_mass; (void)d_seq; (void)d_mkey; (void)d_mkey2; (void)d_idx; (void)d_idx2; (void)tmp; (void)d_raw; (void)d_norm; \
    (void)d_keep; (void)d_pos; (void)d_off;

// Part 0: the raw protein bytes and offsets to the device (they are normalised there, DigestProteinWorker.java:66-70).
// Sorted, unique 300-bit keys of the target sequences (the digest removes them, DatabaseInput.java:480-488).  A host-to-device
// copy, so it is issued before the spectrum upload is queued, not from inside the digest.
bool upload_target_keys_for_reference(cudaStream_t st, const std::vector<std::string>& targets, RefDeviceTable& T, std::string& err) {
    // packed keys (left-aligned 5-bit letters, so tuple order = string order), I -> L folded like the digest
    // (DatabaseInput.add_target_peptides, pg/DatabaseInput.java:47-51), sorted, unique
    struct Key { uint64_t w[kKeyWords]; };
    std::vector<Key> ks; ks.reserve(targets.size());
    for (const std::string& s : targets) {
        if (s.size() > (size_t)kMaxLen) continue;
        Key k; for (int w = 0; w < kKeyWords; ++w) k.w[w] = 0;
        for (size_t r = 0; r < s.size(); ++r) { const char c = s[r] == 'I' ? 'L' : s[r]; k.w[r / 12] |= (uint64_t)(c - 'A' + 1) << (59 - 5 * (r % 12)); }
        ks.push_back(k);
    }
    auto less = [](const Key& a, const Key& b) { for (int w = 0; w < kKeyWords; ++w) if (a.w[w] != b.w[w]) return a.w[w] < b.w[w]; return false; };
    auto same = [](const Key& a, const Key& b) { for (int w = 0; w < kKeyWords; ++w) if (a.w[w] != b.w[w]) return false; return true; };
    std::sort(ks.begin(), ks.end(), less);
    ks.erase(std::unique(ks.begin(), ks.end(), same), ks.end());
    std::vector<uint64_t> tk; tk.reserve(ks.size() * kKeyWords);
    for (const Key& k : ks) for (int w = 0; w < kKeyWords; ++w) tk.push_back(k.w[w]);
    T.n_tkeys = (uint32_t)(tk.size() / kKeyWords);
    DevBuf& d_tk = T.w[10];
    RB(d_tk.reserve(tk.size() * 8 + 64));
    if (T.n_tkeys) RB(cudaMemcpyAsync(d_tk.p, tk.data(), tk.size() * 8, cudaMemcpyHostToDevice, st));
    RB(cudaStreamSynchronize(st));
    return true;
}

bool upload_proteins_for_reference(cudaStream_t st, int32_t n_proteins, const uint64_t* off, const char* res, RefDeviceTable& T, std::string& err) {
    uint64_t total = n_proteins ? off[n_proteins] : 0;
    if (total >= 0x7fffff00ull) { err = "protein database too large for one table"; return false; }
    T.n_forms = 0; T.n_unique = 0; T.n_raw = (uint32_t)total; T.n_proteins = n_proteins;
    if (total == 0) return true;
    K6_BUFFERS(T)
    RB(d_raw.reserve(T.n_raw)); RB(d_off.reserve(((size_t)n_proteins + 1) * 8));
    RB(cudaMemcpyAsync(d_raw.p, res, T.n_raw, cudaMemcpyHostToDevice, st));
    RB(cudaMemcpyAsync(d_off.p, off, ((size_t)n_proteins + 1) * 8, cudaMemcpyHostToDevice, st));
    RB(cudaStreamSynchronize(st));            // the caller's buffers are free again
    return true;
}

// Part 1: digest, unique sequences, minus the targets -> T.uniq / T.n_unique.  Independent of the spectra, so it can run while
// they are still being uploaded (any stream; synchronised at return).
bool digest_reference_on_device(cudaStream_t st, const pq_params& P, const Codes& codes, const std::vector<std::string>& targets, RefDeviceTable& T,
                                uint64_t* launches, std::string& err) {
    T.n_forms = 0; T.n_unique = 0;
    if (T.n_raw == 0) return true;
    const uint32_t n_raw = T.n_raw; const int32_t n_proteins = T.n_proteins;
    K6_BUFFERS(T)
    RB(d_norm.reserve(n_raw)); RB(d_keep.reserve(n_raw)); RB(d_pos.reserve(((size_t)n_raw + 1) * 4));
    RB(T.residues.reserve(n_raw)); RB(d_first.reserve(n_raw)); RB(d_start.reserve(n_raw)); RB(d_cnt.reserve(64));
    k6_normalise<<<(n_raw + 255) / 256, 256, 0, st>>>(d_raw.as<uint8_t>(), n_raw, d_norm.as<uint8_t>(), d_keep.as<uint8_t>());
    {
        size_t bytes = 0;
        cub::TransformInputIterator<uint32_t, ByteToU32, const uint8_t*> keep32(d_keep.as<uint8_t>(), ByteToU32());      // accumulate in 32 bits
        cub::DeviceScan::ExclusiveSum(nullptr, bytes, keep32, d_pos.as<uint32_t>(), (int)n_raw, st);
        RB(tmp.reserve(bytes + 64));
        RB(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, keep32, d_pos.as<uint32_t>(), (int)n_raw, st));
    }
    uint32_t last_pos = 0; uint8_t last_keep = 0;
    RB(cudaMemcpyAsync(&last_pos, d_pos.as<uint32_t>() + (n_raw - 1), 4, cudaMemcpyDeviceToHost, st));
    RB(cudaMemcpyAsync(&last_keep, d_keep.as<uint8_t>() + (n_raw - 1), 1, cudaMemcpyDeviceToHost, st));
    RB(cudaStreamSynchronize(st));
    const uint32_t n = last_pos + (last_keep ? 1u : 0u);
    if (n == 0) return true;
    k6_compact<<<(n_raw + 255) / 256, 256, 0, st>>>(d_norm.as<uint8_t>(), d_keep.as<uint8_t>(), d_pos.as<uint32_t>(), n_raw, T.residues.as<uint8_t>(), d_first.as<uint8_t>());
    k6_protein_starts<<<((uint32_t)n_proteins + 255) / 256, 256, 0, st>>>(d_off.as<uint64_t>(), n_proteins, d_pos.as<uint32_t>(), n_raw, n, d_first.as<uint8_t>());
    *launches += 5;
    const uint8_t* d_res = T.residues.as<uint8_t>();
    EnzymeRule rule; if (!enzyme_rule(P.enzyme, rule)) { err = "unsupported enzyme index"; return false; }
    k6_cut_flags<<<(n + 255) / 256, 256, 0, st>>>(d_res, d_first.as<uint8_t>(), n, rule, d_start.as<uint8_t>());
    // segment starts (+ sentinel)
    RB(T.starts.reserve(((size_t)n + 2) * 4));
    if (!select_flagged(d_start.as<uint8_t>(), n, T.starts.as<uint32_t>(), d_cnt.as<uint32_t>(), tmp, st, err, 0)) return false;
    uint32_t n_seg = 0;
    RB(cudaMemcpyAsync(&n_seg, d_cnt.p, 4, cudaMemcpyDeviceToHost, st));
    RB(cudaStreamSynchronize(st));
    k6_set_u32<<<1, 1, 0, st>>>(T.starts.as<uint32_t>() + n_seg, n);          // (no host-to-device copy here: the copy engine may be busy with the spectra)
    // candidate peptides
    // NoEnzyme (enzyme 0, pg/DatabaseInput.java:117-121): unspecific digestion — every residue starts a piece and a peptide is any
    // run of up to max_len pieces, i.e. every stretch of the protein that passes the length and mass filters
    const uint32_t per = P.enzyme == 0 ? (uint32_t)std::max(P.max_len, 1) : (uint32_t)P.max_missed + 1;
    const uint64_t n_slots64 = (uint64_t)n_seg * per;
    if (n_slots64 >= 0x7fffff00ull) { err = "too many digestion candidates"; return false; }
    const uint32_t n_slots = (uint32_t)n_slots64;
    DigestCfg dc; dc.max_missed = (int32_t)per - 1; dc.min_len = P.min_len; dc.max_len = P.max_len; dc.min_mass = P.min_mass; dc.max_mass = P.max_mass;
    for (int i = 0; i < 26; ++i) dc.residue_mass[i] = codes.residue_mass[i];
    RB(d_flags.reserve((size_t)n_slots + 64)); RB(d_cand.reserve(((size_t)n_slots + 2) * 4));
    k6_candidate_flags<<<(n_slots + 255) / 256, 256, 0, st>>>(T.starts.as<uint32_t>(), n_seg, d_first.as<uint8_t>(), d_res, dc, d_flags.as<uint8_t>());
    if (!select_flagged(d_flags.as<uint8_t>(), n_slots, d_cand.as<uint32_t>(), d_cnt.as<uint32_t>(), tmp, st, err, 0)) return false;
    uint32_t n_cand = 0;
    RB(cudaMemcpyAsync(&n_cand, d_cnt.p, 4, cudaMemcpyDeviceToHost, st));
    RB(cudaStreamSynchronize(st));
    *launches += 6;
    if (n_cand == 0) return true;
    // keys + lexicographic sort (LSD over the 60-bit words the longest allowed peptide needs)
    RB(d_keys.reserve((size_t)n_cand * 8 * kKeyWords)); RB(T.cand_pos.reserve((size_t)n_cand * 4)); RB(T.cand_len.reserve((size_t)n_cand * 4));
    RB(d_perm.reserve((size_t)n_cand * 4)); RB(d_perm2.reserve((size_t)n_cand * 4)); RB(d_tmpkey.reserve((size_t)n_cand * 8)); RB(d_tmpkey2.reserve((size_t)n_cand * 8));
    k6_pack_keys<<<(n_cand + 255) / 256, 256, 0, st>>>(d_cand.as<uint32_t>(), n_cand, T.starts.as<uint32_t>(), per, d_res, d_keys.as<uint64_t>(),
                                                        T.cand_pos.as<uint32_t>(), T.cand_len.as<uint32_t>(), d_perm.as<uint32_t>());
    {
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_tmpkey.as<uint64_t>(), d_tmpkey2.as<uint64_t>(), d_perm.as<uint32_t>(), d_perm2.as<uint32_t>(), (int)n_cand, 0, 64, st);
        RB(tmp.reserve(bytes + 64));
        uint32_t* pa = d_perm.as<uint32_t>(); uint32_t* pb = d_perm2.as<uint32_t>();
        const int words_used = (P.max_len + 11) / 12;
        // first word only, then the short runs of equal first words put in order by their heads (k6_order_runs)
        uint32_t overflow = words_used > 1 && !getenv("PQ_DIGEST_LSD") ? 0u : 1u;
        if (!overflow) {
            RB(cudaMemsetAsync(d_cnt.as<uint32_t>() + 4, 0, 4, st));
            RB(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_keys.as<uint64_t>(), d_tmpkey2.as<uint64_t>(), pa, pb, (int)n_cand, 4, 64, st));      // pa = identity: word 0 needs no gather
            k6_order_runs<<<(n_cand + 255) / 256, 256, 0, st>>>(d_tmpkey2.as<uint64_t>(), pb, d_keys.as<uint64_t>(), n_cand, words_used, d_cnt.as<uint32_t>() + 4);
            RB(cudaMemcpyAsync(&overflow, d_cnt.as<uint32_t>() + 4, 4, cudaMemcpyDeviceToHost, st));
            RB(cudaStreamSynchronize(st));
            *launches += 5;
            if (!overflow) RB(cudaMemcpyAsync(d_perm.p, pb, (size_t)n_cand * 4, cudaMemcpyDeviceToDevice, st));
        }
        if (overflow) {      // long runs (a repetitive database) or one-word keys: least-significant word first, one stable radix sort per word
            pa = d_perm.as<uint32_t>(); pb = d_perm2.as<uint32_t>();
            if (words_used > 1) k6_pack_keys<<<(n_cand + 255) / 256, 256, 0, st>>>(d_cand.as<uint32_t>(), n_cand, T.starts.as<uint32_t>(), per, d_res, d_keys.as<uint64_t>(),
                                                                                    T.cand_pos.as<uint32_t>(), T.cand_len.as<uint32_t>(), d_perm.as<uint32_t>());      // identity permutation again
            for (int w = words_used - 1; w >= 0; --w) {
                k6_gather_u64<<<(n_cand + 255) / 256, 256, 0, st>>>(d_keys.as<uint64_t>() + (size_t)w * n_cand, pa, n_cand, d_tmpkey.as<uint64_t>());
                RB(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_tmpkey.as<uint64_t>(), d_tmpkey2.as<uint64_t>(), pa, pb, (int)n_cand, 4, 64, st));
                std::swap(pa, pb);
                *launches += 5;
            }
            if (pa != d_perm.as<uint32_t>()) RB(cudaMemcpyAsync(d_perm.p, pa, (size_t)n_cand * 4, cudaMemcpyDeviceToDevice, st));
        }
    }
    const uint32_t n_t = T.n_tkeys;          // uploaded by upload_target_keys_for_reference
    (void)targets;
    RB(d_flags.reserve((size_t)n_cand + 64)); RB(T.uniq.reserve(((size_t)n_cand + 2) * 4));
    k6_unique_flags<<<(n_cand + 255) / 256, 256, 0, st>>>(d_perm.as<uint32_t>(), d_keys.as<uint64_t>(), n_cand, d_tk.as<uint64_t>(), n_t, d_flags.as<uint8_t>());
    // uniq[] must hold candidate ids (perm values), not positions: select over the perm array
    {
        size_t bytes = 0;
        cub::DeviceSelect::Flagged(nullptr, bytes, d_perm.as<uint32_t>(), d_flags.as<uint8_t>(), T.uniq.as<uint32_t>(), d_cnt.as<uint32_t>(), (int)n_cand, st);
        RB(tmp.reserve(bytes + 64));
        RB(cub::DeviceSelect::Flagged(tmp.p, bytes, d_perm.as<uint32_t>(), d_flags.as<uint8_t>(), T.uniq.as<uint32_t>(), d_cnt.as<uint32_t>(), (int)n_cand, st));
    }
    uint32_t n_uniq = 0;
    RB(cudaMemcpyAsync(&n_uniq, d_cnt.p, 4, cudaMemcpyDeviceToHost, st));
    RB(cudaStreamSynchronize(st));
    *launches += 4;
    T.n_unique = n_uniq;
    return true;
}

// Part 2: isoforms of the unique sequences inside the mass window of the matched spectra, mass-sorted rows.
bool expand_reference_on_device(cudaStream_t st, const pq_params& P, const Codes& codes, double lo, double hi, RefDeviceTable& T, uint64_t* launches,
                                std::string& err) {
    (void)P;
    T.n_forms = 0;
    const uint32_t n_uniq = T.n_unique;
    if (n_uniq == 0) return true;
    K6_BUFFERS(T)
    const uint8_t* d_res = T.residues.as<uint8_t>();
    // forms
    ExpandCfg ec; memset(&ec, 0, sizeof ec);
    ec.n_var = codes.n_var; ec.n_fixed = codes.n_fixed; ec.max_var = codes.max_var; ec.lo = lo; ec.hi = hi;
    for (int m = 0; m < codes.n_var + codes.n_fixed; ++m) {
        ec.position[m] = codes.mods[(size_t)m].position; ec.residue_mask[m] = codes.mods[(size_t)m].residue_mask; ec.delta[m] = codes.mods[(size_t)m].delta;
        memcpy(ec.code_of[m], codes.shuffle.code_of[m], 26); ec.term_code[m] = codes.shuffle.term_code[m];
    }
    for (int i = 0; i < 26; ++i) ec.residue_mass[i] = codes.residue_mass[i];
    RB(d_count.reserve(((size_t)n_uniq + 2) * 4)); RB(d_foff.reserve(((size_t)n_uniq + 2) * 4));
    RB(cudaMemsetAsync(d_count.as<uint32_t>() + n_uniq, 0, 4, st));
    k6_expand<false><<<(n_uniq + 127) / 128, 128, 0, st>>>(T.uniq.as<uint32_t>(), n_uniq, T.cand_pos.as<uint32_t>(), T.cand_len.as<uint32_t>(), d_res, ec,
                                                           d_count.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr);
    {
        size_t bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_count.as<uint32_t>(), d_foff.as<uint32_t>(), (int)n_uniq + 1, st);
        RB(tmp.reserve(bytes + 64));
        RB(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, d_count.as<uint32_t>(), d_foff.as<uint32_t>(), (int)n_uniq + 1, st));
    }
    uint32_t n_forms = 0;
    RB(cudaMemcpyAsync(&n_forms, d_foff.as<uint32_t>() + n_uniq, 4, cudaMemcpyDeviceToHost, st));
    RB(cudaStreamSynchronize(st));
    *launches += 3;
    T.n_forms = n_forms;
    if (n_forms == 0) return true;
    RB(d_rows.reserve((size_t)n_forms * kRowBytes)); RB(d_mass.reserve((size_t)n_forms * 8)); RB(d_seq.reserve((size_t)n_forms * 4));
    k6_expand<true><<<(n_uniq + 127) / 128, 128, 0, st>>>(T.uniq.as<uint32_t>(), n_uniq, T.cand_pos.as<uint32_t>(), T.cand_len.as<uint32_t>(), d_res, ec,
                                                          nullptr, d_foff.as<uint32_t>(), d_rows.as<uint8_t>(), d_mass.as<double>(), d_seq.as<uint32_t>());
    // stable sort by mass: equal masses keep (sequence, form) generation order
    RB(d_mkey.reserve((size_t)n_forms * 8)); RB(d_mkey2.reserve((size_t)n_forms * 8)); RB(d_idx.reserve((size_t)n_forms * 4)); RB(d_idx2.reserve((size_t)n_forms * 4));
    k6_mass_keys<<<(n_forms + 255) / 256, 256, 0, st>>>(d_mass.as<double>(), n_forms, d_mkey.as<uint64_t>(), d_idx.as<uint32_t>());
    {
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_mkey.as<uint64_t>(), d_mkey2.as<uint64_t>(), d_idx.as<uint32_t>(), d_idx2.as<uint32_t>(), (int)n_forms, 0, 64, st);
        RB(tmp.reserve(bytes + 64));
        RB(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_mkey.as<uint64_t>(), d_mkey2.as<uint64_t>(), d_idx.as<uint32_t>(), d_idx2.as<uint32_t>(), (int)n_forms, 0, 64, st));
    }
    RB(T.rows->reserve((size_t)n_forms * kRowBytes)); RB(T.mass->reserve((size_t)n_forms * 8)); RB(T.seq_id.reserve((size_t)n_forms * 4));
    k6_gather_rows<<<(n_forms * 4 + 255) / 256, 256, 0, st>>>(d_idx2.as<uint32_t>(), n_forms, d_rows.as<uint8_t>(), d_mass.as<double>(), d_seq.as<uint32_t>(),
                                                              T.rows->as<uint8_t>(), T.mass->as<double>(), T.seq_id.as<uint32_t>());
    RB(cudaStreamSynchronize(st));
    RB(cudaGetLastError());
    *launches += 8;
    return true;
}

__global__ void k6_u32_to_i32(const uint32_t* __restrict__ in, uint32_t n, int32_t* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)in[i];
}
__global__ void k6_gather_target_rows(const uint32_t* __restrict__ order, uint32_t n, const uint8_t* __restrict__ rows_in, const double* __restrict__ mass_in,
                                      uint8_t* __restrict__ rows, double* __restrict__ mass) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t i = t >> 2, q = t & 3;
    if (i >= n) return;
    uint32_t s = order[i];
    reinterpret_cast<uint4*>(rows)[(size_t)i * 4 + q] = reinterpret_cast<const uint4*>(rows_in)[(size_t)s * 4 + q];
    if (q == 0) mass[i] = mass_in[s];
}

// The TARGET peptide table on the device: CreatePeptideInputWorker -> PeptideInput.calcPeptideIsoforms (pg/PeptideInput.java:114-237)
// with the same expansion kernel as the reference table, every form kept (no mass window), then a stable mass sort.
// Public form order = generation order (sequence order, then calcPeptideIsoforms order); device row order = mass order;
// form_of_row translates.  flat = the sequences back to back (letters 'A'..'Z'), off[n_seq + 1].
bool build_target_table_on_device(cudaStream_t st, const Codes& codes, uint32_t n_seq, const uint32_t* off, const char* flat, TargetDeviceTable& T,
                                  uint64_t* launches, std::string& err) {
    T.n_forms = 0;
    if (n_seq == 0) return true;
    const uint32_t total = off[n_seq];
    DevBuf& d_res = T.w[0]; DevBuf& d_pos = T.w[1]; DevBuf& d_len = T.w[2]; DevBuf& d_ident = T.w[3]; DevBuf& d_count = T.w[4]; DevBuf& d_foff = T.w[5];
    DevBuf& d_rows = T.w[6]; DevBuf& d_mass = T.w[7]; DevBuf& d_seq = T.w[8]; DevBuf& d_mkey = T.w[9]; DevBuf& d_mkey2 = T.w[10]; DevBuf& d_idx = T.w[11]; DevBuf& tmp = T.w[12];
    std::vector<uint32_t> h((size_t)n_seq * 3);                  // pos | len | identity
    std::vector<uint8_t> letters(total);
    for (uint32_t i = 0; i < n_seq; ++i) { h[i] = off[i]; h[n_seq + i] = off[i + 1] - off[i]; h[2 * (size_t)n_seq + i] = i; }
    for (uint32_t i = 0; i < total; ++i) letters[i] = (uint8_t)(flat[i] - 'A');
    RB(d_res.reserve(total + 64)); RB(d_pos.reserve((size_t)n_seq * 12)); RB(T.seq_off->reserve(((size_t)n_seq + 1) * 4)); RB(T.seq_letters->reserve(total + 64));
    RB(cudaMemcpyAsync(d_res.p, flat, total, cudaMemcpyHostToDevice, st));
    RB(cudaMemcpyAsync(d_pos.p, h.data(), h.size() * 4, cudaMemcpyHostToDevice, st));
    RB(cudaMemcpyAsync(T.seq_off->p, off, ((size_t)n_seq + 1) * 4, cudaMemcpyHostToDevice, st));
    RB(cudaMemcpyAsync(T.seq_letters->p, letters.data(), total, cudaMemcpyHostToDevice, st));
    const uint32_t* pos = d_pos.as<uint32_t>(); const uint32_t* len = pos + n_seq; const uint32_t* ident = pos + 2 * (size_t)n_seq;
    (void)d_len; (void)d_ident;
    ExpandCfg ec; memset(&ec, 0, sizeof ec);
    ec.n_var = codes.n_var; ec.n_fixed = codes.n_fixed; ec.max_var = codes.max_var; ec.lo = -1.0; ec.hi = 1.7976931348623157e308;
    for (int m = 0; m < codes.n_var + codes.n_fixed; ++m) {
        ec.position[m] = codes.mods[(size_t)m].position; ec.residue_mask[m] = codes.mods[(size_t)m].residue_mask; ec.delta[m] = codes.mods[(size_t)m].delta;
        memcpy(ec.code_of[m], codes.shuffle.code_of[m], 26); ec.term_code[m] = codes.shuffle.term_code[m];
    }
    for (int i = 0; i < 26; ++i) ec.residue_mass[i] = codes.residue_mass[i];
    RB(d_count.reserve(((size_t)n_seq + 2) * 4)); RB(d_foff.reserve(((size_t)n_seq + 2) * 4));
    RB(cudaMemsetAsync(d_count.as<uint32_t>() + n_seq, 0, 4, st));
    k6_expand<false><<<(n_seq + 127) / 128, 128, 0, st>>>(ident, n_seq, pos, len, d_res.as<uint8_t>(), ec, d_count.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr);
    {
        size_t bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_count.as<uint32_t>(), d_foff.as<uint32_t>(), (int)n_seq + 1, st);
        RB(tmp.reserve(bytes + 64));
        RB(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, d_count.as<uint32_t>(), d_foff.as<uint32_t>(), (int)n_seq + 1, st));
    }
    uint32_t n_forms = 0;
    RB(cudaMemcpyAsync(&n_forms, d_foff.as<uint32_t>() + n_seq, 4, cudaMemcpyDeviceToHost, st));
    RB(cudaStreamSynchronize(st));           // also: the host staging vectors above are free again
    *launches += 3;
    T.n_forms = n_forms;
    if (n_forms == 0) return true;
    RB(d_rows.reserve((size_t)n_forms * kRowBytes)); RB(d_mass.reserve((size_t)n_forms * 8)); RB(d_seq.reserve((size_t)n_forms * 4));
    RB(T.form_mods->reserve((size_t)n_forms * sizeof(FormMods))); RB(T.seq_of_form->reserve((size_t)n_forms * 4)); RB(T.form_of_row->reserve((size_t)n_forms * 4));
    RB(T.rows->reserve((size_t)n_forms * kRowBytes)); RB(T.mass->reserve((size_t)n_forms * 8));
    k6_expand<true><<<(n_seq + 127) / 128, 128, 0, st>>>(ident, n_seq, pos, len, d_res.as<uint8_t>(), ec, nullptr, d_foff.as<uint32_t>(), d_rows.as<uint8_t>(),
                                                         d_mass.as<double>(), d_seq.as<uint32_t>(), T.form_mods->as<FormMods>());
    k6_u32_to_i32<<<(n_forms + 255) / 256, 256, 0, st>>>(d_seq.as<uint32_t>(), n_forms, T.seq_of_form->as<int32_t>());
    // stable sort by mass: equal masses keep generation order
    RB(d_mkey.reserve((size_t)n_forms * 8)); RB(d_mkey2.reserve((size_t)n_forms * 8)); RB(d_idx.reserve((size_t)n_forms * 4));
    k6_mass_keys<<<(n_forms + 255) / 256, 256, 0, st>>>(d_mass.as<double>(), n_forms, d_mkey.as<uint64_t>(), d_idx.as<uint32_t>());
    {
        size_t bytes = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, bytes, d_mkey.as<uint64_t>(), d_mkey2.as<uint64_t>(), d_idx.as<uint32_t>(), T.form_of_row->as<uint32_t>(), (int)n_forms, 0, 64, st);
        RB(tmp.reserve(bytes + 64));
        RB(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, d_mkey.as<uint64_t>(), d_mkey2.as<uint64_t>(), d_idx.as<uint32_t>(), T.form_of_row->as<uint32_t>(), (int)n_forms, 0, 64, st));
    }
    k6_gather_target_rows<<<(n_forms * 4 + 255) / 256, 256, 0, st>>>(T.form_of_row->as<uint32_t>(), n_forms, d_rows.as<uint8_t>(), d_mass.as<double>(),
                                                                     T.rows->as<uint8_t>(), T.mass->as<double>());
    RB(cudaGetLastError());
    *launches += 8;
    return true;
}

bool build_reference_on_device(cudaStream_t st, const pq_params& P, const Codes& codes, int32_t n_proteins, const uint64_t* off, const char* res,
                               const std::vector<std::string>& targets, double lo, double hi, RefDeviceTable& T, uint64_t* launches, std::string& err) {
    return upload_proteins_for_reference(st, n_proteins, off, res, T, err) && upload_target_keys_for_reference(st, targets, T, err) &&
           digest_reference_on_device(st, P, codes, targets, T, launches, err) &&
           expand_reference_on_device(st, P, codes, lo, hi, T, launches, err);
}

}  // namespace pq

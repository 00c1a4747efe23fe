// k2_shuffle.cu — K2: the shuffled-peptide generator (Step 4's candidates), bit-compatible with the reference's
// seeded RNG streams.
//
//   k2_generate   PeptideInput.generateRandomPeptidesFast (pg/PeptideInput.java:376-476): ONE java.util.Random
//                 (seed 12457) per target sequence drives all rounds, so a sequence is a strictly serial stream:
//                 one thread per target sequence replays it (thousands of sequences run side by side).
//   k2_dedup      the HashSet filter (:464-469): a shuffle equal to the target or to an earlier shuffle is dropped
//                 after its draws were consumed.  One CTA per sequence, 64-bit hashes in shared memory, exact
//                 byte compare on hash equality, order-preserving compaction.
//   k2_place_mods PeptideInput.getModificationPeptide (:485-557): a fresh Random(2022) per (target form, shuffle)
//                 re-places the form's modifications; then addFixedModification (:213-237).  One thread per row.
#include "pq_internal.h"
#include "k_cells.cuh"

namespace pq {

namespace {

struct JRandom {                      // java.util.Random (JDK spec): 48-bit LCG
    uint64_t seed;
    __device__ explicit JRandom(int64_t s) : seed(((uint64_t)s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1)) {}
    __device__ int32_t next31() {
        seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (int32_t)(seed >> 17);
    }
    __device__ int32_t nextInt(int32_t bound) {
        int32_t r = next31();
        int32_t m = bound - 1;
        if ((bound & m) == 0) return (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        for (int32_t u = r;; u = next31()) {
            r = u % bound;
            if ((int64_t)u - (int64_t)r + (int64_t)m <= 0x7fffffffLL) break;   // Java int overflow test
        }
        return r;
    }
};

// position of the r-th (0-based) set bit: a binary search on population counts (__fns is a loop over the bits in software: it was
// 57 % of k2_generate's instructions)
__device__ __forceinline__ int select64(uint64_t x, int r) {
    uint32_t w = (uint32_t)x;
    int pos = 0, c = __popc(w);
    if (r >= c) { r -= c; pos = 32; w = (uint32_t)(x >> 32); }
    c = __popc(w & 0xFFFFu); if (r >= c) { r -= c; pos += 16; w >>= 16; }
    c = __popc(w & 0xFFu);   if (r >= c) { r -= c; pos += 8; w >>= 8; }
    c = __popc(w & 0xFu);    if (r >= c) { r -= c; pos += 4; w >>= 4; }
    c = __popc(w & 0x3u);    if (r >= c) { r -= c; pos += 2; w >>= 2; }
    if (r >= (int)(w & 1u)) pos += 1;
    return pos;
}

// Affine map of n LCG steps: state -> A*state + C (mod 2^48).  The stream of one sequence is strictly serial, but the
// number of draws per round is known (L-1, plus a retry of nextInt's rejection loop about once per 1e8 draws), so a lane
// can jump straight to the start of its round.
struct Affine { uint64_t a, c; };
__device__ __forceinline__ Affine affine_pow(uint64_t n) {
    const uint64_t M = (1ULL << 48) - 1;
    Affine r{1ULL, 0ULL}, b{0x5DEECE66DULL, 0xBULL};
    while (n) {
        if (n & 1ULL) { r.a = (r.a * b.a) & M; r.c = (r.c * b.a + b.c) & M; }      // r = b o r
        b.c = (b.c * b.a + b.c) & M; b.a = (b.a * b.a) & M;                          // b = b o b
        n >>= 1;
    }
    return r;
}

constexpr int kGenWarps = 4;      // sequences per CTA (one warp each)

// One warp per target sequence.  Lane l generates rounds base + l, base + l + 32, ...; every round starts from the LCG
// state reached by skipping (round - base) * (L-1) draws from the state at `base`.  If some round had to redraw
// (java.util.Random.nextInt's rejection loop), every later round starts one draw later: the warp restarts after the
// first such round from its true end state.  Rounds before it are already exact.
__global__ void __launch_bounds__(32 * kGenWarps)
k2_generate(const ShuffleSeq* __restrict__ seqs, uint32_t n_seqs, int64_t seed, uint8_t* __restrict__ raw) {
    __shared__ __align__(16) uint8_t s_rows[kGenWarps][32][kRowBytes + 16];        // +16: rows of neighbouring lanes start in different banks
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t t = blockIdx.x * kGenWarps + warp;
    if (t >= n_seqs) return;
    const ShuffleSeq& S = seqs[t];
    const int len = (int)S.len - 1;                     // C-terminal residue stays (fixCtermAA)
    const uint32_t num = S.num;
    const uint64_t M = (1ULL << 48) - 1;
    const uint64_t all = len >= 63 ? ~1ull : (((1ull << len) - 1ull) << 1);   // positions 1..len
    uint8_t* row = s_rows[warp][lane];
    const Affine per_round = affine_pow((uint64_t)len);
    const Affine per_lane = affine_pow((uint64_t)len * lane);
    const Affine per_sweep = affine_pow((uint64_t)len * 32u);
    uint64_t base_state = ((uint64_t)seed ^ 0x5DEECE66DULL) & M;
    uint32_t base_round = 0;
    (void)per_round;
    while (base_round < num) {
        uint64_t st = (per_lane.a * base_state + per_lane.c) & M;              // state at the start of round base + lane
        uint32_t first_retry = 0xFFFFFFFFu; uint64_t retry_state = 0;
        for (uint32_t i = base_round + lane; i < num; i += 32) {
            JRandom rnd(0); rnd.seed = st;
            uint64_t free_ = all; int draws = 0;
            for (int k = 0; k < kRowBytes / 4; ++k) reinterpret_cast<uint32_t*>(row)[k] = 0u;
            for (uint32_t g = 0; g < S.n_groups; ++g) {
                const uint8_t code = S.group_code[g];
                for (uint32_t r_ = 0; r_ < S.group_count[g]; ++r_) {
                    const int bound = __popcll(free_);
                    // nextInt(bound) with the draw count exposed
                    int32_t r = rnd.next31(); ++draws;
                    const int32_t m = bound - 1;
                    if ((bound & m) == 0) r = (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
                    else {
                        for (int32_t u = r;; u = rnd.next31(), ++draws) {
                            r = u % bound;
                            if ((int64_t)u - (int64_t)r + (int64_t)m <= 0x7fffffffLL) break;
                        }
                    }
                    const int pp = select64(free_, r);
                    row[kRowHeader + pp - 1] = code;
                    free_ &= ~(1ull << pp);
                }
            }
            row[kRowHeader + len] = S.seq[len];
            row[0] = (uint8_t)S.len;
            uint4* dst = reinterpret_cast<uint4*>(raw + (size_t)(S.out_base + i) * kRowBytes);
            const uint4* srcv = reinterpret_cast<const uint4*>(row);
            dst[0] = srcv[0]; dst[1] = srcv[1]; dst[2] = srcv[2]; dst[3] = srcv[3];
            if (draws != len && first_retry == 0xFFFFFFFFu) { first_retry = i; retry_state = rnd.seed; }
            if (first_retry != 0xFFFFFFFFu) break;                               // later rounds of this lane start from a wrong state anyway
            st = (per_sweep.a * st + per_sweep.c) & M;
        }
        // the earliest round that consumed extra draws decides where the stream really continues
        uint32_t fr = first_retry;
        for (int o = 16; o > 0; o >>= 1) fr = min(fr, __shfl_xor_sync(0xffffffffu, fr, o));
        if (fr == 0xFFFFFFFFu) break;
        const int owner = __ffs(__ballot_sync(0xffffffffu, first_retry == fr)) - 1;
        base_state = __shfl_sync(0xffffffffu, retry_state, owner);
        base_round = fr + 1;
    }
}

// 64-bit hash of `len` residue codes at a 4-byte aligned address, four codes per step (any hash serves: equal hashes are followed
// by an exact comparison; the byte-wise FNV this replaces was a third of the dedup kernel)
__device__ __forceinline__ uint64_t codes_hash(const uint8_t* p, int len) {
    uint64_t h = 0x9E3779B97F4A7C15ull ^ (uint64_t)len;
    for (int k = 0; k < len; k += 4) {
        uint32_t w = *reinterpret_cast<const uint32_t*>(p + k);
        const int rem = len - k;
        if (rem < 4) w &= (1u << (8 * rem)) - 1u;
        h = (h ^ w) * 0xD6E8FEB86659FD93ull;
        h ^= h >> 32;
    }
    return h;
}
__device__ __forceinline__ uint64_t row_hash(const uint8_t* row, int len) { return codes_hash(row + kRowHeader, len); }
__device__ __forceinline__ bool row_equal(const uint8_t* a, const uint8_t* b, int len) {
    for (int k = 0; k < len; ++k) if (a[kRowHeader + k] != b[kRowHeader + k]) return false;
    return true;
}

__global__ void k2_dedup_quadratic(const ShuffleSeq* __restrict__ seqs, const uint8_t* __restrict__ raw, uint32_t* __restrict__ kept_index,
                         uint32_t* __restrict__ kept_count, uint32_t max_num) {
    extern __shared__ __align__(16) uint8_t sm[];
    uint64_t* hs = reinterpret_cast<uint64_t*>(sm);                  // [max_num]
    uint8_t* flag = reinterpret_cast<uint8_t*>(hs + max_num);        // [max_num]
    __shared__ uint32_t s_base;
    const ShuffleSeq& S = seqs[blockIdx.x];
    const int L = (int)S.len;
    const uint8_t* rows = raw + (size_t)S.out_base * kRowBytes;
    const uint64_t ht = codes_hash(S.seq, L);
    for (uint32_t i = threadIdx.x; i < S.num; i += blockDim.x) hs[i] = row_hash(rows + (size_t)i * kRowBytes, L);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < S.num; i += blockDim.x) {
        const uint8_t* ri = rows + (size_t)i * kRowBytes;
        uint64_t h = hs[i];
        bool keep = true;
        if (h == ht) { bool eq = true; for (int k = 0; k < L; ++k) if (ri[kRowHeader + k] != S.seq[k]) { eq = false; break; } if (eq) keep = false; }
        for (uint32_t j = 0; keep && j < i; ++j)
            if (hs[j] == h && row_equal(ri, rows + (size_t)j * kRowBytes, L)) keep = false;
        flag[i] = keep ? 1 : 0;
    }
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    // order-preserving compaction, one block-wide chunk at a time
    for (uint32_t c0 = 0; c0 < S.num; c0 += blockDim.x) {
        uint32_t i = c0 + threadIdx.x;
        uint32_t f = (i < S.num) ? flag[i] : 0u;
        uint32_t bal = __ballot_sync(0xffffffffu, f);
        __shared__ uint32_t s_w[32];
        uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) s_w[wid] = __popc(bal);
        __syncthreads();
        uint32_t before = 0;
        for (uint32_t w = 0; w < wid; ++w) before += s_w[w];
        uint32_t pos = s_base + before + __popc(bal & ((1u << lane) - 1u));
        if (f) kept_index[S.out_base + pos] = i;
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t tot = 0; for (uint32_t w = 0; w < blockDim.x / 32; ++w) tot += s_w[w]; s_base += tot; }
        __syncthreads();
    }
    if (threadIdx.x == 0) kept_count[blockIdx.x] = s_base;
}

// Same filter with a shared-memory hash table (open addressing on the 64-bit row hash, smallest row index per hash): O(n)
// instead of O(n^2) hash compares.  A row is a duplicate iff an earlier row is byte-equal; the table gives the first row
// with the same hash, an exact compare confirms; if two different rows ever share a 64-bit hash the row falls back to the
// exact scan over all earlier rows, so the result is identical to the quadratic kernel.
__global__ void k2_dedup_hashed(const ShuffleSeq* __restrict__ seqs, const uint8_t* __restrict__ raw, uint32_t* __restrict__ kept_index,
                                uint32_t* __restrict__ kept_count, uint32_t max_num, uint32_t table) {
    extern __shared__ __align__(16) uint8_t sm[];
    unsigned long long* hk = reinterpret_cast<unsigned long long*>(sm);             // [table] hash keys, 0 = empty
    unsigned long long* hs = hk + table;                                            // [max_num] row hashes
    uint32_t* first = reinterpret_cast<uint32_t*>(hs + max_num);                    // [table] smallest row index of the key
    uint8_t* flag = reinterpret_cast<uint8_t*>(first + table);                      // [max_num]
    __shared__ uint32_t s_base;
    const ShuffleSeq& S = seqs[blockIdx.x];
    const int L = (int)S.len;
    const uint8_t* rows = raw + (size_t)S.out_base * kRowBytes;
    unsigned long long ht = codes_hash(S.seq, L);
    if (ht == 0ull) ht = 1ull;
    for (uint32_t i = threadIdx.x; i < table; i += blockDim.x) { hk[i] = 0ull; first[i] = 0xFFFFFFFFu; }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < S.num; i += blockDim.x) {
        unsigned long long h = row_hash(rows + (size_t)i * kRowBytes, L);
        if (h == 0ull) h = 1ull;
        hs[i] = h;
        for (uint32_t slot = (uint32_t)h & (table - 1);; slot = (slot + 1) & (table - 1)) {
            unsigned long long old = atomicCAS(&hk[slot], 0ull, h);
            if (old == 0ull || old == h) { atomicMin(&first[slot], i); break; }
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < S.num; i += blockDim.x) {
        const uint8_t* ri = rows + (size_t)i * kRowBytes;
        const unsigned long long h = hs[i];
        bool keep = true;
        if (h == ht) { bool eq = true; for (int k = 0; k < L; ++k) if (ri[kRowHeader + k] != S.seq[k]) { eq = false; break; } if (eq) keep = false; }
        if (keep) {
            uint32_t slot = (uint32_t)h & (table - 1);
            while (hk[slot] != h) slot = (slot + 1) & (table - 1);
            const uint32_t f = first[slot];
            if (f != i) {
                if (row_equal(ri, rows + (size_t)f * kRowBytes, L)) keep = false;
                else for (uint32_t j = 0; keep && j < i; ++j)                       // two rows share a 64-bit hash: exact scan
                    if (hs[j] == h && row_equal(ri, rows + (size_t)j * kRowBytes, L)) keep = false;
            }
        }
        flag[i] = keep ? 1 : 0;
    }
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    // order-preserving compaction, one block-wide chunk at a time
    for (uint32_t c0 = 0; c0 < S.num; c0 += blockDim.x) {
        uint32_t i = c0 + threadIdx.x;
        uint32_t f = (i < S.num) ? flag[i] : 0u;
        uint32_t bal = __ballot_sync(0xffffffffu, f);
        __shared__ uint32_t s_w[32];
        uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        if (lane == 0) s_w[wid] = __popc(bal);
        __syncthreads();
        uint32_t before = 0;
        for (uint32_t w = 0; w < wid; ++w) before += s_w[w];
        uint32_t pos = s_base + before + __popc(bal & ((1u << lane) - 1u));
        if (f) kept_index[S.out_base + pos] = i;
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t tot = 0; for (uint32_t w = 0; w < blockDim.x / 32; ++w) tot += s_w[w]; s_base += tot; }
        __syncthreads();
    }
    if (threadIdx.x == 0) kept_count[blockIdx.x] = s_base;
}

__global__ void k2_place_mods(const ShuffleSeq* __restrict__ seqs, const ShuffleForm* __restrict__ forms, uint32_t n_forms,
                              const ShuffleMods* __restrict__ M, const uint8_t* __restrict__ raw,
                              const uint32_t* __restrict__ kept_index, const uint32_t* __restrict__ kept_count,
                              uint8_t* __restrict__ rows, const ModTable* __restrict__ table, uint4* __restrict__ lists) {
    __shared__ double2 s_tab[kMaxCodes];
    __shared__ double s_nt[kMaxTermCodes], s_ct[kMaxTermCodes];
    for (int i = threadIdx.x; i < kMaxCodes; i += blockDim.x) s_tab[i] = make_double2(table->aa[i], table->delta[i]);
    for (int i = threadIdx.x; i < kMaxTermCodes; i += blockDim.x) { s_nt[i] = table->nterm[i]; s_ct[i] = table->cterm[i]; }
    __syncthreads();
    const uint32_t f = blockIdx.x;
    if (f >= n_forms) return;
    const ShuffleForm& F = forms[f];
    const ShuffleSeq& S = seqs[F.seq_slot];
    const uint32_t nk = kept_count[F.seq_slot];
    const int L = (int)S.len;
    for (uint32_t k = blockIdx.y * blockDim.x + threadIdx.x; k < nk; k += gridDim.y * blockDim.x) {
        const uint8_t* src = raw + (size_t)(S.out_base + kept_index[S.out_base + k]) * kRowBytes;
        uint32_t w[16];
        const uint4* sp4 = reinterpret_cast<const uint4*>(src);
        uint4 q0 = sp4[0], q1 = sp4[1], q2 = sp4[2], q3 = sp4[3];
        w[0] = q0.x; w[1] = q0.y; w[2] = q0.z; w[3] = q0.w; w[4] = q1.x; w[5] = q1.y; w[6] = q1.z; w[7] = q1.w;
        w[8] = q2.x; w[9] = q2.y; w[10] = q2.z; w[11] = q2.w; w[12] = q3.x; w[13] = q3.y; w[14] = q3.z; w[15] = q3.w;
        uint8_t* b = reinterpret_cast<uint8_t*>(w);
        uint8_t letters[kMaxLen];
        for (int i = 0; i < L; ++i) letters[i] = b[kRowHeader + i];
        uint64_t used = 0ull;                       // bit = site (0 .. L+1)
        JRandom rnd(M->mod_seed);
        auto sites_of = [&](int m) -> uint64_t {
            uint64_t s = 0ull;
            uint32_t mask = M->residue_mask[m];
            switch (M->position[m]) {
                case PQ_MOD_ANYWHERE: for (int i = 0; i < L; ++i) if ((mask >> letters[i]) & 1u) s |= 1ull << (i + 1); break;
                case PQ_MOD_PEP_NTERM: case PQ_MOD_PROT_NTERM: s = 1ull; break;
                case PQ_MOD_PEP_CTERM: case PQ_MOD_PROT_CTERM: s = 1ull << (L + 1); break;
                case PQ_MOD_PEP_NTERM_AA: case PQ_MOD_PROT_NTERM_AA: if ((mask >> letters[0]) & 1u) s = 1ull << 1; break;
                case PQ_MOD_PEP_CTERM_AA: case PQ_MOD_PROT_CTERM_AA: if ((mask >> letters[L - 1]) & 1u) s = 1ull << L; break;
                default: break;
            }
            return s;
        };
        auto apply = [&](int m, int site) {
            if (site == 0) b[1] = M->term_code[m];
            else if (site == L + 1) b[2] = M->term_code[m];
            else b[kRowHeader + site - 1] = M->code_of[m][letters[site - 1]];
        };
        for (uint32_t q = 0; q < F.n_mods; ++q) {                     // target form's modifications, list order
            int m = F.mod[q];
            uint64_t s = sites_of(m) & ~used;
            int n = __popcll(s);
            if (n >= 1) {
                int idx = rnd.nextInt(n);
                int site = select64(s, idx);
                apply(m, site);
                used |= 1ull << site;
            }
        }
        const uint64_t var_sites = used;                               // addFixedModification: sites taken before the call
        for (int m = M->n_var; m < M->n_mods; ++m) {
            uint64_t s = sites_of(m) & ~var_sites;
            while (s) { int site = __ffsll((long long)s) - 1; s &= s - 1; apply(m, site); }
        }
        uint4* dst = reinterpret_cast<uint4*>(rows + (size_t)(F.row_base + k) * kRowBytes);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]); dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        dst[2] = make_uint4(w[8], w[9], w[10], w[11]); dst[3] = make_uint4(w[12], w[13], w[14], w[15]);
        // the spectrum-independent half of K1's Step-4 count pass: the cell of every (ion, fragment charge) of this row (k_cells.cuh)
        if (lists) write_cell_lists(b, s_tab, s_nt, s_ct, (int)F.list_nch, lists + F.list_base + k, F.list_chunks, (size_t)F.list_rows);
    }
}

}  // namespace

int launch_shuffles(const ShuffleSeq* seqs, uint32_t n_seqs, const ShuffleForm* forms, uint32_t n_forms,
                    const ShuffleMods* mods, uint8_t* raw, uint32_t* kept_index, uint32_t* kept_count, uint8_t* rows,
                    uint32_t max_num, int64_t shuffle_seed, const ModTable* mod_table, uint4* lists, cudaStream_t st) {
    if (n_seqs == 0) return 0;
    k2_generate<<<(n_seqs + kGenWarps - 1) / kGenWarps, 32 * kGenWarps, 0, st>>>(seqs, n_seqs, shuffle_seed, raw);
    uint32_t table = 64; while (table < 2 * max_num) table <<= 1;
    const size_t smem_hashed = (size_t)table * 12 + (size_t)max_num * 9 + 16;
    if (smem_hashed <= 160 * 1024) {
        cudaFuncSetAttribute(k2_dedup_hashed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_hashed);
        k2_dedup_hashed<<<n_seqs, 256, smem_hashed, st>>>(seqs, raw, kept_index, kept_count, max_num, table);
    } else {
        size_t smem = (size_t)max_num * 9 + 16;
        cudaFuncSetAttribute(k2_dedup_quadratic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k2_dedup_quadratic<<<n_seqs, 256, smem, st>>>(seqs, raw, kept_index, kept_count, max_num);
    }
    if (n_forms > 0) {
        dim3 grid(n_forms, (max_num + 127) / 128);
        k2_place_mods<<<grid, 128, 0, st>>>(seqs, forms, n_forms, mods, raw, kept_index, kept_count, rows, mod_table, lists);
    }
    return 3;
}

}  // namespace pq

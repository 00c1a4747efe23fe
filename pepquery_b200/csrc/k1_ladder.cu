// k1_ladder.cu — Step 3 (DBSearchWorker.run, pg/DBSearchWorker.java:25-113) with the fragment ladders of the reference rows
// precomputed.
//
// A reference peptide form is scored against every spectrum whose precursor window holds it (~15 spectra at C2 size), and its
// theoretical fragment m/z — PeptideSpectrumAnnotator's b/y ladder (JMatch.java:388-466), walked residue by residue with
// explicit round-to-nearest adds — depend on the row only.  So the ladder is walked ONCE per row, when the table is built:
//   pm[side][k]      the mass of ion k before the charge is applied (b: running sum; y: running sum + H2O), the very doubles the
//                    ladder-walking scorer of k1_device.cuh holds in `acc`
//   cl[side][c][k]   for fragment charge c = 1..3: 2 * cell_filter(m/z of ion k at charge c), the byte offset of its cell in the
//                    prepared spectrum's table (pq_common.h); slots that are no ion (k < c, A6; padding) point at the
//                    always-clean sentinel entry behind the table
// The per-pair scorer then is: scan the cell lists (one 16-byte load per eight ions, one shared-memory flag test per ion and
// charge) into bit masks of the ions that land in a cell where something can be claimed; for those few (~1 in 10) load the
// ion mass, apply the charge and resolve the match exactly as the ladder-walking scorer does (same m/z bits, same claim order:
// ion number ascending, charge ascending, b before y), so counts and scores are bit-identical.  ~3x fewer instructions per pair,
// no per-thread row columns or residue table in shared memory.  Charges above 3 (precursor charge >= 5) are derived from pm[]
// on the fly.
#include <algorithm>
#include <string>

#include "k1_device.cuh"

namespace pq {

namespace {

using namespace dev;

__host__ __device__ __forceinline__ uint32_t lad_pm_units(int Lm1) { return (uint32_t)(Lm1 + 1) >> 1; }      // 16-byte units of one side's ion masses
__host__ __device__ __forceinline__ uint32_t lad_cl_units(int Lm1) { return (uint32_t)(Lm1 + 7) >> 3; }      // units of one (side, charge) cell list

// one thread per row; the row's slot (stride units of 16 bytes, sized for the longest peptide of the search) holds, packed at its
// front, [pm b][pm y][cl b c1..c3][cl y c1..c3]
constexpr int kFillBlock = 128;
__global__ void __launch_bounds__(kFillBlock)
k_ladder_fill(const uint8_t* __restrict__ rows, uint32_t n, uint32_t stride, const ModTable* __restrict__ M, uint4* __restrict__ out, uint8_t* __restrict__ len_out) {
    __shared__ uint32_t s_codes[16][kFillBlock];
    __shared__ double2 s_tab[kMaxCodes];
    __shared__ double s_nt[kMaxTermCodes], s_ct[kMaxTermCodes];
    const int tid = threadIdx.x;
    for (int i = tid; i < kMaxCodes; i += kFillBlock) s_tab[i] = make_double2(M->aa[i], M->delta[i]);
    for (int i = tid; i < kMaxTermCodes; i += kFillBlock) { s_nt[i] = M->nterm[i]; s_ct[i] = M->cterm[i]; }
    const uint32_t i = blockIdx.x * kFillBlock + tid;
    if (i < n) {
        const uint4* rp = reinterpret_cast<const uint4*>(rows + (size_t)i * kRowBytes);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const uint4 v = __ldg(rp + q); s_codes[4 * q][tid] = v.x; s_codes[4 * q + 1][tid] = v.y; s_codes[4 * q + 2][tid] = v.z; s_codes[4 * q + 3][tid] = v.w; }
    }
    __syncthreads();
    if (i >= n) return;
    const uint32_t head = s_codes[0][tid];
    const int L = (int)(head & 0xFFu), Lm1 = L - 1;
    len_out[i] = (uint8_t)L;                 // the lengths once more, one byte per row: a warp of K1 reads 32 of them from one sector
    if (Lm1 <= 0) return;
    auto residue = [&](int idx) -> uint32_t { return (s_codes[1 + (idx >> 2)][tid] >> (8 * (idx & 3))) & 0xFFu; };      // 0-based residue
    const uint32_t PM = lad_pm_units(Lm1), U = lad_cl_units(Lm1);
    uint4* base = out + (size_t)i * stride;
    constexpr uint32_t kSent2 = 0x00010001u * kCellSentinel;
    for (int side = 0; side < 2; ++side) {
        double2* pm = reinterpret_cast<double2*>(base + (size_t)side * PM);
        uint4* cl = base + 2 * (size_t)PM + (size_t)side * kLadderCharges * U;
        double acc = side == 0 ? s_nt[(head >> 8) & 0xFFu] : s_ct[(head >> 16) & 0xFFu];
        double prev = 0.0;
        for (uint32_t u = 0; u < U; ++u) {      // a unit (eight ions) at a time: every slot of the packed words is a compile-time position
            uint32_t w[kLadderCharges][4];
#pragma unroll
            for (int c = 0; c < kLadderCharges; ++c) { w[c][0] = kSent2; w[c][1] = kSent2; w[c][2] = kSent2; w[c][3] = kSent2; }
            auto ion = [&](int k, auto slot) {
                constexpr int J = decltype(slot)::value;
                const double2 t = s_tab[residue(side == 0 ? k - 1 : L - k)];
                acc = __dadd_rn(__dadd_rn(acc, t.x), t.y);
                const double m = side == 0 ? acc : __dadd_rn(acc, kWater);
                if (!(J & 1)) { prev = m; if (k == Lm1) pm[(k - 1) >> 1] = make_double2(m, 0.0); }      // ion k sits in slot (k - 1) & 7: odd k = even slot
                else pm[(k - 1) >> 1] = make_double2(prev, m);
                unit_put<J>(w[0], 2u * cell_filter(__dadd_rn(m, kProton)));
                if (k >= 2) unit_put<J>(w[1], 2u * cell_filter(__dmul_rn(__dadd_rn(m, 2.0 * kProton), 0.5)));
                if (k >= 3) unit_put<J>(w[2], 2u * cell_filter(div3_rn(__dadd_rn(m, __dmul_rn(3.0, kProton)))));
            };
            const int k0 = 8 * (int)u + 1;
            if (k0 + 0 <= Lm1) ion(k0 + 0, std::integral_constant<int, 0>());
            if (k0 + 1 <= Lm1) ion(k0 + 1, std::integral_constant<int, 1>());
            if (k0 + 2 <= Lm1) ion(k0 + 2, std::integral_constant<int, 2>());
            if (k0 + 3 <= Lm1) ion(k0 + 3, std::integral_constant<int, 3>());
            if (k0 + 4 <= Lm1) ion(k0 + 4, std::integral_constant<int, 4>());
            if (k0 + 5 <= Lm1) ion(k0 + 5, std::integral_constant<int, 5>());
            if (k0 + 6 <= Lm1) ion(k0 + 6, std::integral_constant<int, 6>());
            if (k0 + 7 <= Lm1) ion(k0 + 7, std::integral_constant<int, 7>());
#pragma unroll
            for (int c = 0; c < kLadderCharges; ++c) cl[(size_t)c * U + u] = make_uint4(w[c][0], w[c][1], w[c][2], w[c][3]);
        }
    }
}

// bit j of the result: entry j of the unit lands in a cell where something can be claimed
__device__ __forceinline__ uint32_t dirty8(const uint4& q, uint32_t cells_s) {
    auto clean = [&](uint32_t e, int j) -> uint32_t { return (lds_u16(cells_s + e) >> (15 - j)) & (1u << j); };
    const uint32_t c = clean(q.x & 0xFFFFu, 0) | clean(q.x >> 16, 1) | clean(q.y & 0xFFFFu, 2) | clean(q.y >> 16, 3) |
                       clean(q.z & 0xFFFFu, 4) | clean(q.z >> 16, 5) | clean(q.w & 0xFFFFu, 6) | clean(q.w >> 16, 7);
    return c ^ 0xFFu;
}

struct LadderMasks { uint64_t b[kLadderCharges], y[kLadderCharges]; };      // bit k-1 of [c-1]: ion k at charge c may match

// The cell lists of one row against the staged spectrum: which (ion, charge) land in a cell where something can be claimed.
// Both sides' units of a trip are in flight together (the lists come from L2).
__device__ __forceinline__ void scan_ladder(const uint4* __restrict__ lad, int Lm1, int nch, uint32_t cells_s, LadderMasks& M) {
    const uint32_t PM = lad_pm_units(Lm1), U = lad_cl_units(Lm1);
    const uint4* clb = lad + 2 * (size_t)PM;
    const uint4* cly = clb + (size_t)kLadderCharges * U;
#pragma unroll
    for (int c = 0; c < kLadderCharges; ++c) { M.b[c] = 0ull; M.y[c] = 0ull; }
#pragma unroll 1
    for (uint32_t u = 0; u < U; ++u) {
        const uint4 b1 = __ldg(clb + u), y1 = __ldg(cly + u);
        uint4 b2 = b1, y2 = y1, b3 = b1, y3 = y1;
        if (nch >= 2) { b2 = __ldg(clb + U + u); y2 = __ldg(cly + U + u); }
        if (nch >= 3) { b3 = __ldg(clb + 2 * U + u); y3 = __ldg(cly + 2 * U + u); }
        M.b[0] |= (uint64_t)dirty8(b1, cells_s) << (8 * u); M.y[0] |= (uint64_t)dirty8(y1, cells_s) << (8 * u);
        if (nch >= 2) { M.b[1] |= (uint64_t)dirty8(b2, cells_s) << (8 * u); M.y[1] |= (uint64_t)dirty8(y2, cells_s) << (8 * u); }
        if (nch >= 3) { M.b[2] |= (uint64_t)dirty8(b3, cells_s) << (8 * u); M.y[2] |= (uint64_t)dirty8(y3, cells_s) << (8 * u); }
    }
}

// One pair from the row's tables: the ions that may match are resolved in claim order (b before y, ion ascending, charge
// ascending: A5), the mass of the next one already on its way while the current one is matched.
template <int BLOCK, int METHOD>
__device__ __forceinline__ PairScore score_ladder(const uint4* __restrict__ lad, int Lm1, const BlobView& B, int z, double itol,
                                               const double* log10_fact, const double* ln_fact, int32_t predicted, uint32_t* usedw) {
    const int nch = z <= 1 ? 1 : z - 1;                        // fragment charges 1 .. z-1 (JMatch.java:398-407)
    const uint32_t PM = lad_pm_units(Lm1);
    LadderMasks M;
    scan_ladder(lad, Lm1, min(nch, kLadderCharges), smem_addr(B.cells), M);
    uint64_t used = 0ull;
    int ca = 0, cb = 0, key0 = 0, key1 = 0, key2 = 0;
    double sum = 0.0;
    if (METHOD == 2) for (int w = 0; w < 10; ++w) usedw[w * BLOCK] = 0u;
    // a surviving peak counts once, for the first ion that reaches it, in annotator order (A5): same rule as score_pair_n
    auto claim = [&](uint32_t code, bool is_b) {
        if (code == kAnsNone) return;
        if (METHOD == 1) {
            const uint64_t bit = 1ull << code;
            if (used & bit) return;
            used |= bit;
            if (is_b) ++ca; else ++cb;
            sum = __dadd_rn(sum, B.val[code]);
        } else {
            const uint32_t sid = code & 0xFFFu;
            const uint32_t w = usedw[(sid >> 5) * BLOCK], bit = 1u << (sid & 31);
            if (w & bit) return;
            usedw[(sid >> 5) * BLOCK] = w | bit;
            const uint32_t c = (code >> 12) & 3u;
            key0 += c == 0; key1 += c == 1; key2 += c == 2;
        }
    };
    const uint64_t high = nch > kLadderCharges && Lm1 > 3 ? ((~0ull >> (64 - Lm1)) & ~7ull) : 0ull;      // charges >= 4 are not in the lists: every ion k >= 4 is looked at
    // ONE loop over both sides (all b ions, then all y ions): a warp runs as long as its busiest lane, and the busiest lane of the
    // combined work is less busy than the busiest b lane plus the busiest y lane
    const double* pmb = reinterpret_cast<const double*>(lad);
    const double* pmy = reinterpret_cast<const double*>(lad + (size_t)PM);
    uint64_t tb = M.b[0] | M.b[1] | M.b[2] | high, ty = M.y[0] | M.y[1] | M.y[2] | high;
    auto next = [&](int& side, int& b, double& m) {           // the next ion to look at and its mass (the load is issued here)
        side = tb ? 0 : 1;
        const uint64_t t = tb ? tb : ty;
        b = __ffsll((long long)t) - 1;
        m = __ldg((side ? pmy : pmb) + b);
        if (tb) tb &= tb - 1; else ty &= ty - 1;
    };
    int side_n = 0, b_n = 0; double m_n = 0.0;
    bool more = (tb | ty) != 0ull;
    if (more) next(side_n, b_n, m_n);
#pragma unroll 1
    while (more) {
        const int side = side_n, b = b_n; const double m = m_n;
        more = (tb | ty) != 0ull;
        if (more) next(side_n, b_n, m_n);
        const int k = b + 1;
        const uint64_t d1 = side ? M.y[0] : M.b[0], d2 = side ? M.y[1] : M.b[1], d3 = side ? M.y[2] : M.b[2];
        if ((d1 >> b) & 1ull) claim(match_ion<METHOD>(B, __dadd_rn(m, kProton), itol), side == 0);
        if ((d2 >> b) & 1ull) claim(match_ion<METHOD>(B, __dmul_rn(__dadd_rn(m, 2.0 * kProton), 0.5), itol), side == 0);
        if ((d3 >> b) & 1ull) claim(match_ion<METHOD>(B, div3_rn(__dadd_rn(m, __dmul_rn(3.0, kProton))), itol), side == 0);
        if (nch > kLadderCharges && k >= 4) {
            double t = __dmul_rn(__dadd_rn(m, 4.0 * kProton), 0.25);
            if (ion_may_match(B, t)) claim(match_ion<METHOD>(B, t, itol), side == 0);
            for (int c = 5; c <= nch && c <= k; ++c) {
                t = __ddiv_rn(__dadd_rn(m, __dmul_rn((double)c, kProton)), (double)c);
                if (ion_may_match(B, t)) claim(match_ion<METHOD>(B, t, itol), side == 0);
            }
        }
    }
    return finish_pair<METHOD>(B, ca, cb, sum, key0, key1, key2, predicted, log10_fact, ln_fact);
}

#ifndef K1L_MIN_BLOCKS
#define K1L_MIN_BLOCKS 12
#endif
// One CTA per job (= one matched spectrum x one isotope error x the run of reference rows around its precursor mass), jobs handed
// out by an atomic counter, the prepared spectrum staged by one TMA bulk copy, one thread per (row, spectrum) pair: the job
// protocol and the reduction of k1_score_kernel (k1_score.cu); only the pair scorer differs.
// (Tried on top of this and dropped, both parity-green: skipping the candidates whose score bound — from the scan alone — is below
// the target's score and the best score seen so far, either with the resolve work pooled per warp through a shared-memory queue
// or with the survivors compacted per job into a dense list.  84 % of the candidates can be skipped, but the launch is bound by
// the dependent chain of each job on its few resident warps, and every variant added a phase to that chain: 6.3 ms vs 5.4 ms.)
template <int BLOCK, int METHOD>
__global__ void __launch_bounds__(BLOCK, BLOCK == 64 ? K1L_MIN_BLOCKS : 6)
k1_ladder_kernel(ScoreLaunch L, ScoreConfig cfg, uint32_t stage_bytes, int nstage) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ double s_red_d[BLOCK / 32];
    __shared__ uint32_t s_red_u[3][BLOCK / 32];
    __shared__ uint32_t s_job[2];
    double* s_l10 = reinterpret_cast<double*>(smem);                                  // [130]
    uint32_t* usedw_all = reinterpret_cast<uint32_t*>(s_l10 + 130);                   // MVH only: [10][BLOCK]
    uint8_t* stage0 = smem + (((size_t)130 * 8 + (METHOD == 2 ? (size_t)10 * BLOCK * 4 : 0) + 127) & ~(size_t)127);

    const int tid = threadIdx.x;
    for (int i = tid; i < 130; i += BLOCK) s_l10[i] = L.log10_fact[i];
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const double itol = cfg.itol;
    const bool prefetch = cfg.prefetch_next != 0;
    uint32_t* usedw = usedw_all + tid;

    auto load = [&](uint32_t slot, uint32_t nj) {      // thread 0: start the TMA of job nj into `slot`
        s_job[slot] = nj;
        if (nj < L.n_jobs) {
            const Job& Jn = L.jobs[nj];
            uint32_t bytes = L.blob_len[Jn.blob];
            mbar_expect_tx(&s_bar[slot], bytes);
            tma_load_1d(stage0 + (size_t)stage_bytes * slot, L.blobs + L.blob_off[Jn.blob], bytes, &s_bar[slot]);
        }
    };
    auto fetch = [&](uint32_t slot) { load(slot, atomicAdd(L.job_counter, 1u)); };      // take the next job and start its TMA
    // one stage: the next job is taken while the current one is scored and its blob pulled towards L2, so the copy that starts when
    // the stage is free again comes from L2 instead of HBM
    uint32_t next_job = 0xFFFFFFFFu;                   // (thread 0 only)
    auto claim = [&]() {
        next_job = atomicAdd(L.job_counter, 1u);
        if (next_job < L.n_jobs) { const Job& Jn = L.jobs[next_job]; prefetch_l2_bulk(L.blobs + L.blob_off[Jn.blob], L.blob_len[Jn.blob]); }
    };
    if (tid == 0) fetch(0);
    __syncthreads();
    uint32_t job = s_job[0];
    for (uint32_t it = 0; job < L.n_jobs; ++it) {
        const uint32_t cur = nstage == 2 ? (it & 1u) : 0u;
        const uint32_t parity = nstage == 2 ? ((it >> 1) & 1u) : (it & 1u);
        if (tid == 0) { if (nstage == 2) fetch(cur ^ 1u); else if (prefetch) claim(); }
        const Job J = L.jobs[job];
        mbar_wait(&s_bar[cur], parity);
        const BlobView B = view_blob(stage0 + (size_t)stage_bytes * cur);
        const long long va = jround(__dmul_rn(J.mass, 10.0));

        uint32_t n_in = 0, n_better = 0;
        double best = -1.0; uint32_t best_cand = 0xFFFFFFFFu;
        for (uint32_t c0 = 0; c0 < J.cand_count; c0 += BLOCK) {
            const uint32_t c = c0 + tid;
            if (c >= J.cand_count) continue;
            const uint32_t row = J.cand_begin + c;
            if (!in_window(L.row_mass[row], J.mass, va, cfg)) continue;                // DBSearchWorker.java:38-52
            const int Lm1 = (int)L.lad_len[row] - 1;
            int32_t pred = 0;
            if (METHOD == 2) pred = L.row_pred[2 * (size_t)row + (J.charge == 2 ? 0 : 1)];
            const PairScore ps = score_ladder<BLOCK, METHOD>(L.lad + (size_t)row * L.lad_stride, Lm1, B, J.charge, itol, s_l10, L.ln_fact, pred, usedw);
            ++n_in;
            if (ps.score > best) { best = ps.score; best_cand = row; }                 // rows ascending: the first of equals stays (:63-65)
            if (J.ref_score <= ps.score) {                                              // DBSearchWorker.java:68-76
                ++n_better;
                if (ps.score >= 20.0) {
                    uint32_t slot = atomicAdd(L.hit_count, 1u);
                    if (slot < L.hit_capacity) {
                        Hit h; h.spectrum = J.spectrum; h.cand = row; h.score = ps.score; h.mass = J.mass;
                        h.count_a = ps.a; h.count_b = ps.b; h.iso = J.iso; h.aux = J.aux; h.blob = J.blob; h.pad = 0;
                        L.hits[slot] = h;
                    }
                }
            }
        }
        // ---- CTA reduction of the job ----
        for (int o = 16; o > 0; o >>= 1) {
            n_in += __shfl_xor_sync(0xffffffffu, n_in, o);
            n_better += __shfl_xor_sync(0xffffffffu, n_better, o);
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            uint32_t oc = __shfl_xor_sync(0xffffffffu, best_cand, o);
            if (ob > best || (ob == best && oc < best_cand)) { best = ob; best_cand = oc; }
        }
        if ((tid & 31) == 0) { s_red_u[0][tid >> 5] = n_in; s_red_u[1][tid >> 5] = n_better; s_red_u[2][tid >> 5] = best_cand; s_red_d[tid >> 5] = best; }
        __syncthreads();     // also: every thread is done with this stage's blob
        if (tid == 0) {
            uint32_t a = 0, b = 0; double bs = -1.0; uint32_t bc = 0xFFFFFFFFu;
            for (int w = 0; w < BLOCK / 32; ++w) {
                a += s_red_u[0][w]; b += s_red_u[1][w];
                if (s_red_d[w] > bs || (s_red_d[w] == bs && s_red_u[2][w] < bc)) { bs = s_red_d[w]; bc = s_red_u[2][w]; }
            }
            JobResult R; R.n_in_window = a; R.n_better = b; R.best_score = bs; R.best_cand = bc; R.pad = 0;
            L.results[job] = R;
            if (nstage == 1) { if (prefetch) load(0, next_job); else fetch(0); }
        }
        __syncthreads();
        job = s_job[nstage == 2 ? (cur ^ 1u) : 0u];
    }
}

template <int BLOCK, int METHOD>
int launch_ladder_kernel(const ScoreLaunch& L, const ScoreConfig& cfg, cudaStream_t st, int nstage) {
    const uint32_t stage = (L.max_blob_bytes + 127u) & ~127u;
    const size_t fixed = ((size_t)130 * 8 + (METHOD == 2 ? (size_t)10 * BLOCK * 4 : 0) + 127) & ~(size_t)127;
    const size_t smem = (size_t)stage * nstage + fixed;
    auto kernel = k1_ladder_kernel<BLOCK, METHOD>;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0; cudaGetDevice(&dev);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BLOCK, smem);
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (uint32_t)(sms * per_sm);
    if (grid > L.n_jobs) grid = L.n_jobs;
    cudaMemsetAsync(L.job_counter, 0, 4, st);
    kernel<<<grid, BLOCK, smem, st>>>(L, cfg, stage, nstage);
    return 1;
}

}  // namespace

int launch_score_ladder(const ScoreLaunch& L, const ScoreConfig& cfg, cudaStream_t st) {
    if (L.n_jobs == 0) return 0;
    int nstage = 1, block = 64;
    if (const char* e = getenv("PQ_K1_STAGES")) nstage = atoi(e) == 2 ? 2 : 1;
    if (const char* e = getenv("PQ_STEP3_BLOCK")) block = atoi(e) == 128 ? 128 : 64;
    if (block == 128) return cfg.method == 2 ? launch_ladder_kernel<128, 2>(L, cfg, st, nstage) : launch_ladder_kernel<128, 1>(L, cfg, st, nstage);
    return cfg.method == 2 ? launch_ladder_kernel<64, 2>(L, cfg, st, nstage) : launch_ladder_kernel<64, 1>(L, cfg, st, nstage);
}

// Builds the ladder tables of a row table (n rows of 64 bytes, none longer than max_len residues): one slot of `stride` 16-byte
// units per row.  Returns the stride through *stride_units.
bool build_row_ladders(cudaStream_t st, const uint8_t* rows, uint32_t n, int max_len, const ModTable* mods, DevBuf& lad, DevBuf& lad_len, uint32_t* stride_units,
                       uint64_t* launches, std::string& err) {
    const int Lm1 = std::max(std::min(max_len, (int)kMaxLen) - 1, 1);
    const uint32_t stride = 2u * lad_pm_units(Lm1) + 2u * (uint32_t)kLadderCharges * lad_cl_units(Lm1);
    *stride_units = stride;
    if (n == 0) return true;
    cudaError_t e = lad.reserve((size_t)n * stride * 16);
    if (e == cudaSuccess) e = lad_len.reserve((size_t)n + 64);
    if (e != cudaSuccess) { err = std::string("ladder tables: ") + cudaGetErrorString(e); return false; }
    k_ladder_fill<<<(n + kFillBlock - 1) / kFillBlock, kFillBlock, 0, st>>>(rows, n, stride, mods, lad.as<uint4>(), lad_len.as<uint8_t>());
    e = cudaGetLastError();
    if (e != cudaSuccess) { err = std::string("ladder tables: ") + cudaGetErrorString(e); return false; }
    if (launches) *launches += 1;
    return true;
}

}  // namespace pq

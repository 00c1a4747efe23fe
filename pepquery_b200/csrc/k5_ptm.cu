// k5_ptm.cu — K5: unrestricted-modification competitor search (Step 5).
//
// Replaces ModificationDB.getCandidateRefPeptidesFast (OpenModificationSearch/ModificationDB.java:993-1062) ->
// ModPepQueryAndScoringWorker.do_score(Modification) (ModPepQueryAndScoringWorker.java:149-276) ->
// ModificationDB.getCandidatePTMs(Peptide, Modification) (ModificationDB.java:263-326): for one validated spectrum, every
// PTM specificity t (not protein-terminal, -250 <= delta <= 250), every reference form whose mass + delta_t sits in the
// precursor window (bucket round(10*(precursor - delta_t)) +-1 and |ppm| <= tol), and every site of t on that form that
// carries no modification yet, gives one candidate = the form plus t at that site; each is scored like any other pair.
//
// One CTA per validated spectrum (blob staged once by TMA).  Phase 1: one thread per PTM binary-searches the mass-sorted
// reference table for its shifted window; a block scan turns the counts into a flat item list.  Phase 2: one thread per
// (PTM, reference row) item checks the exact predicate and a residue-presence mask, then walks the free sites, scoring
// each candidate with the shared pair scorer (k1_device.cuh, Extra{site, delta}).  Hyperscore only (MVH would need a
// predicted-peak count per candidate).
#include "k4_device.cuh"

namespace pq {

namespace {

using namespace dev;

constexpr int kPtmBlock = 256;
constexpr int kMaxPtm = 2048;

__device__ __forceinline__ uint32_t lower_bound_g(const double* a, uint32_t n, double v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (__ldg(a + mid) < v) lo = mid + 1; else hi = mid; }
    return lo;
}

template <int METHOD>
__global__ void __launch_bounds__(kPtmBlock)
k5_ptm_kernel(PtmLaunch L, ScoreConfig cfg, uint32_t stage_bytes) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_scan[kPtmBlock / 32];
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_red_u[3][kPtmBlock / 32];
    __shared__ int s_stop;
    __shared__ uint32_t s_fill;
    __shared__ uint64_t s_cand[2 * kPtmBlock];          // compacted (row, PTM, site) candidates waiting to be scored
    constexpr int BLOCK = kPtmBlock;

    uint8_t* stage0 = smem;
    uint32_t* codes_all = reinterpret_cast<uint32_t*>(smem + stage_bytes);            // [16][BLOCK]
    double2* s_tab = reinterpret_cast<double2*>(codes_all + 16 * BLOCK);                               // {residue mass, modification mass} per code
    double* s_nt = reinterpret_cast<double*>(s_tab + kMaxCodes);
    double* s_ct = s_nt + kMaxTermCodes;
    double* s_l10 = s_ct + kMaxTermCodes;                                               // [65]
    uint32_t* s_lo = reinterpret_cast<uint32_t*>(s_l10 + 130);                           // [kMaxPtm]
    uint32_t* s_off = s_lo + kMaxPtm;                                                   // [kMaxPtm + 1]
    uint32_t* usedw_all = s_off + kMaxPtm + 2;                                          // MVH only: [10][BLOCK]

    const int tid = threadIdx.x;
    for (int i = tid; i < kMaxCodes; i += BLOCK) s_tab[i] = make_double2(L.mods->aa[i], L.mods->delta[i]);
    for (int i = tid; i < kMaxTermCodes; i += BLOCK) { s_nt[i] = L.mods->nterm[i]; s_ct[i] = L.mods->cterm[i]; }
    for (int i = tid; i < 130; i += BLOCK) s_l10[i] = L.log10_fact[i];
    if (tid == 0) { mbar_init(&s_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const double itol = cfg.itol;
    uint32_t* codes = codes_all + tid;
    uint32_t* usedw = usedw_all + tid;
    const bool plain_iso = L.n_isotope == 1 && L.isotope[0] == 0;

    uint32_t it = 0;
    for (uint32_t job = blockIdx.x; job < L.n_jobs; job += gridDim.x, ++it) {
        const Job J = L.jobs[job];
        if (tid == 0) {
            uint32_t bytes = L.blob_len[J.blob];
            mbar_expect_tx(&s_bar, bytes);
            tma_load_1d(stage0, L.blobs + L.blob_off[J.blob], bytes, &s_bar);
            s_stop = 0; s_fill = 0;
        }
        mbar_wait(&s_bar, it & 1u);
        __syncthreads();
        const BlobView B = view_blob(stage0);
        uint32_t n_pairs = 0, n_better = 0, n_bound = 0;

        double mono = J.mass; int cur_iso = 0;
        // one thread per candidate: the reference form plus one PTM at one site, scored like any other pair
        auto score_batch = [&](uint32_t base, uint32_t count) {
            if (tid < (int)count) {
                const uint64_t cand = s_cand[base + tid];
                const uint32_t row = (uint32_t)cand, t = (uint32_t)(cand >> 32) & 0xFFFFu; const int site = (int)(cand >> 48);
                const pq_ptm P = L.ptms[t];
                const uint4* rp = reinterpret_cast<const uint4*>(L.rows + (size_t)row * kRowBytes);
                uint4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2), q3 = __ldg(rp + 3);
                codes[0 * BLOCK] = q0.x; codes[1 * BLOCK] = q0.y; codes[2 * BLOCK] = q0.z; codes[3 * BLOCK] = q0.w;
                codes[4 * BLOCK] = q1.x; codes[5 * BLOCK] = q1.y; codes[6 * BLOCK] = q1.z; codes[7 * BLOCK] = q1.w;
                codes[8 * BLOCK] = q2.x; codes[9 * BLOCK] = q2.y; codes[10 * BLOCK] = q2.z; codes[11 * BLOCK] = q2.w;
                codes[12 * BLOCK] = q3.x; codes[13 * BLOCK] = q3.y; codes[14 * BLOCK] = q3.z; codes[15 * BLOCK] = q3.w;
                Extra ex; ex.site = site; ex.delta = P.delta;
                int32_t predicted = 0;
                if (METHOD == 2) predicted = predict_peaks(L.rows + (size_t)row * kRowBytes, L.mods, itol, J.charge == 2 ? 1 : 2, ex);   // MVHMatch.java:165-172
                // only candidates that reach the target's score matter (rows of ptm.txt, n_ptm): the score bound decides most others
                ++n_pairs;
                // (hyperscore: the count pass of Step 4 with the extra modification; survivors are ~0.1 %, so no compaction here)
                if (METHOD == 1 && !cfg.score_all &&
                    below_score_bound<BLOCK, 1, 1, true>(codes, B, s_tab, s_nt, s_ct, J.charge, s_l10, J.ref_score, nullptr, 0, ex)) { ++n_bound; return; }
                PairScore ps = score_pair<BLOCK, METHOD>(codes, B, s_tab, s_nt, s_ct, J.charge, itol, ex, s_l10, L.ln_fact, predicted, usedw,
                                                         (cfg.score_all || METHOD == 1) ? 0.0 : J.ref_score);
                if (ps.a < 0) { ++n_bound; return; }                                           // decided by the bound: below the target
                bool better = cfg.hc ? ps.score >= J.ref_score : ps.score > J.ref_score;      // :239-259, PSMT:920-943
                if (better) { ++n_better; if (cfg.fast) s_stop = 1; }
                if (ps.score >= cfg.min_score && ps.score >= J.ref_score) {                    // rows of ptm.txt (ModificationDB.java:1660-1663)
                    uint32_t slot = atomicAdd(L.hit_count, 1u);
                    if (slot < L.hit_capacity) {
                        Hit h; h.spectrum = J.spectrum; h.cand = row; h.score = ps.score; h.mass = __dadd_rn(__ldg(L.row_mass + row), P.delta);
                        h.count_a = ps.a; h.count_b = ps.b; h.iso = cur_iso; h.aux = (int32_t)((t << 8) | (uint32_t)site);
                        h.blob = J.blob; h.pad = 0;
                        L.hits[slot] = h;
                    }
                }
            }
        };
        for (int ii = 0; ii < L.n_isotope; ++ii) {
            mono = J.mass; cur_iso = plain_iso ? 0 : L.isotope[ii];                                                       // CParameter.java:423-448
            if (!plain_iso) mono = __dsub_rn(mono, __dmul_rn((double)L.isotope[ii], kIsotope));
            // ---- phase 1: shifted window of every PTM ----
            uint32_t local = 0;
            for (uint32_t t = tid; t < (uint32_t)L.n_ptm; t += BLOCK) {
                const pq_ptm P = L.ptms[t];
                uint32_t lo = 0, cnt = 0;
                if (P.position < PQ_MOD_PROT_NTERM && P.delta >= -250.0 && P.delta <= 250.0) {        // ModificationDB.java:1037-1042
                    double c = __dsub_rn(mono, P.delta);
                    double w = cfg.tol_ppm ? cfg.tol * 1e-6 * (fabs(mono) + 1.0) * 1.01 + 1e-6 : cfg.tol * 1.01 + 1e-6;
                    if (w > 0.2) w = 0.2;
                    lo = lower_bound_g(L.row_mass, L.n_rows, c - w);
                    uint32_t hi = lower_bound_g(L.row_mass, L.n_rows, c + w);
                    cnt = hi - lo;
                }
                s_lo[t] = lo; s_off[t] = cnt;
            }
            __syncthreads();
            // ---- block exclusive scan over n_ptm counts (8 per thread) ----
            {
                const uint32_t per = (uint32_t)(L.n_ptm + BLOCK - 1) / BLOCK;
                uint32_t b = tid * per, e = min(b + per, (uint32_t)L.n_ptm);
                for (uint32_t t = b; t < e; ++t) local += s_off[t];
                uint32_t incl = local;
                for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
                if ((tid & 31) == 31) s_scan[tid >> 5] = incl;
                __syncthreads();
                uint32_t base = 0;
                for (int w = 0; w < (tid >> 5); ++w) base += s_scan[w];
                uint32_t run = base + incl - local;
                for (uint32_t t = b; t < e; ++t) { uint32_t c = s_off[t]; s_off[t] = run; run += c; }
                if (tid == BLOCK - 1) { s_off[L.n_ptm] = run; s_total = run; }
                __syncthreads();
            }
            const uint32_t total = s_total;
            // ---- phase 2: enumerate the candidates (PTM, reference row, free site) of one chunk of items, compact them
            // into a shared-memory list, and score the list densely, one thread per candidate, BLOCK at a time ----
            for (uint32_t i0 = 0; i0 < total; i0 += BLOCK) {
                if (cfg.fast && s_stop) break;
                const uint32_t i = i0 + tid;
                uint64_t sites = 0ull;                 // bit s: site s of this item takes the PTM
                uint32_t row = 0, t = 0;
                if (i < total) {
                    uint32_t lo = 0, hi = (uint32_t)L.n_ptm;                                 // last t with off[t] <= i
                    while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (s_off[mid] <= i) lo = mid; else hi = mid; }
                    t = lo;
                    const pq_ptm P = L.ptms[t];
                    row = s_lo[t] + (i - s_off[t]);
                    // exact window (ModPepQueryAndScoringWorker.java:165-203)
                    const double refm = __ldg(L.row_mass + row);
                    const long long ind = jround(__dmul_rn(10.0, __dsub_rn(mono, P.delta)));
                    const long long key = jround(__dmul_rn(refm, 10.0));
                    bool ok = key >= ind - 1 && key <= ind + 1;
                    if (ok) {
                        const double pep_mass = __dadd_rn(refm, P.delta);
                        double del = cfg.tol_ppm ? __ddiv_rn(__dmul_rn(1.0e6, __dsub_rn(pep_mass, mono)), pep_mass) : __dsub_rn(pep_mass, mono);
                        ok = fabs(del) <= cfg.tol;
                    }
                    const uint32_t rcode = P.residue ? (uint32_t)(P.residue - 'A') : 0xFFu;
                    if (ok && P.residue && !((__ldg(L.row_mask + row) >> rcode) & 1u)) ok = false;   // :207-209 residue absent
                    if (ok) {
                        // sites of the PTM on this form that carry nothing yet (ModificationDB.java:268-292)
                        const uint32_t* rw = reinterpret_cast<const uint32_t*>(L.rows + (size_t)row * kRowBytes);
                        const uint32_t head = __ldg(rw);
                        const int Lp = (int)(head & 0xFF);
                        const uint32_t ntc = (head >> 8) & 0xFF, ctc = (head >> 16) & 0xFF;
                        auto residue_ok = [&](int site) {
                            uint32_t w = __ldg(rw + ((kRowHeader + site - 1) >> 2));
                            uint32_t code = (w >> (((kRowHeader + site - 1) & 3) * 8)) & 0xFFu;
                            return P.residue ? code == rcode : code < 26u;
                        };
                        switch (P.position) {
                            case PQ_MOD_ANYWHERE: for (int site = 1; site <= Lp; ++site) if (residue_ok(site)) sites |= 1ull << site; break;
                            case PQ_MOD_PEP_NTERM: if (ntc == 0) sites = 1ull; break;
                            case PQ_MOD_PEP_CTERM: if (ctc == 0) sites = 1ull << (Lp + 1); break;
                            case PQ_MOD_PEP_NTERM_AA: if (residue_ok(1)) sites = 1ull << 1; break;
                            default: if (residue_ok(Lp)) sites = 1ull << Lp; break;                 // PQ_MOD_PEP_CTERM_AA
                        }
                    }
                }
                // append the chunk's candidates round by round (round r = every item's r-th site): at most BLOCK per round
                for (;;) {
                    const bool have = sites != 0ull;
                    const uint32_t bal = __ballot_sync(0xffffffffu, have);
                    if ((tid & 31) == 0) s_scan[tid >> 5] = __popc(bal);
                    __syncthreads();
                    uint32_t before = 0, all = 0;
                    for (int w = 0; w < BLOCK / 32; ++w) { uint32_t c = s_scan[w]; if (w < (tid >> 5)) before += c; all += c; }
                    if (all == 0) { __syncthreads(); break; }
                    if (have) {
                        const int site = __ffsll((long long)sites) - 1;
                        sites &= sites - 1;
                        s_cand[s_fill + before + __popc(bal & ((1u << (tid & 31)) - 1u))] = (uint64_t)row | ((uint64_t)t << 32) | ((uint64_t)site << 48);
                    }
                    __syncthreads();
                    if (tid == 0) s_fill += all;
                    __syncthreads();
                    // score full batches (always leaves room for the next round)
                    while (s_fill >= (uint32_t)BLOCK) {
                        const uint32_t base = s_fill - BLOCK;
                        score_batch(base, BLOCK);
                        __syncthreads();
                        if (tid == 0) s_fill = base;
                        __syncthreads();
                        if (cfg.fast && s_stop) break;
                    }
                    if (cfg.fast && s_stop) break;
                }
            }
            if (!(cfg.fast && s_stop) && s_fill > 0) score_batch(0, s_fill);
            __syncthreads();
            if (tid == 0) s_fill = 0;
            __syncthreads();
        }
        for (int o = 16; o > 0; o >>= 1) { n_pairs += __shfl_xor_sync(0xffffffffu, n_pairs, o); n_better += __shfl_xor_sync(0xffffffffu, n_better, o); n_bound += __shfl_xor_sync(0xffffffffu, n_bound, o); }
        if ((tid & 31) == 0) { s_red_u[0][tid >> 5] = n_pairs; s_red_u[1][tid >> 5] = n_better; s_red_u[2][tid >> 5] = n_bound; }
        __syncthreads();
        if (tid == 0) {
            uint32_t a = 0, b = 0, nb = 0;
            for (int w = 0; w < BLOCK / 32; ++w) { a += s_red_u[0][w]; b += s_red_u[1][w]; nb += s_red_u[2][w]; }
            JobResult R; R.n_in_window = a; R.n_better = b; R.best_score = -1.0; R.best_cand = 0xFFFFFFFFu; R.pad = nb;
            L.results[job] = R;
        }
        __syncthreads();
    }
}

__global__ void k_row_masks(const uint8_t* __restrict__ rows, uint32_t n_rows, uint32_t* __restrict__ mask) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const uint8_t* row = rows + (size_t)r * kRowBytes;
    uint32_t m = 0; int L = row[0];
    for (int k = 0; k < L; ++k) { uint32_t c = row[kRowHeader + k]; if (c < 26u) m |= 1u << c; }     // unmodified residues only
    mask[r] = m;
}

}  // namespace

int launch_row_masks(const uint8_t* rows, uint32_t n_rows, uint32_t* mask, cudaStream_t st) {
    if (n_rows == 0) return 0;
    k_row_masks<<<(n_rows + 255) / 256, 256, 0, st>>>(rows, n_rows, mask);
    return 1;
}

template <int METHOD>
static int launch_ptm_m(const PtmLaunch& L, const ScoreConfig& cfg, cudaStream_t st) {
    uint32_t stage = (L.max_blob_bytes + 127u) & ~127u;
    size_t smem = (size_t)stage + (size_t)16 * kPtmBlock * 4 + (size_t)(2 * kMaxCodes + 2 * kMaxTermCodes + 130) * 8 + (size_t)(2 * kMaxPtm + 4) * 4 +
                  (METHOD == 2 ? (size_t)10 * kPtmBlock * 4 : 0);
    auto kernel = k5_ptm_kernel<METHOD>;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kPtmBlock, smem);
    if (per_sm < 1) per_sm = 1;
    int dev = 0; cudaGetDevice(&dev);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    uint32_t grid = (uint32_t)(sms * per_sm);
    if (grid > L.n_jobs) grid = L.n_jobs;
    kernel<<<grid, kPtmBlock, smem, st>>>(L, cfg, stage);
    return 1;
}

int launch_ptm(const PtmLaunch& L, const ScoreConfig& cfg, cudaStream_t st) {
    if (L.n_jobs == 0 || L.n_ptm == 0) return 0;
    return cfg.method == 2 ? launch_ptm_m<2>(L, cfg, st) : launch_ptm_m<1>(L, cfg, st);
}

}  // namespace pq

// k1_score.cu — K1: the pair scorer.  One CTA owns one job (= one prepared spectrum x a run of candidate peptide
// rows); the spectrum blob is staged into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier; the next
// job is taken early and its blob pulled towards L2 meanwhile, or — 256-thread configuration — loaded into a second
// stage); one THREAD scores one (peptide, spectrum) pair.
//
// Replaces the body of PeptideSearchMT.scorePeptide2Spectrum (pg/PeptideSearchMT.java:1144-1251):
//   annotator   JMatch.getSpectrumAnnotation (PSMMatch/JMatch.java:388-466) -> compomics PeptideSpectrumAnnotator
//   hyperscore  HyperscoreMatch.calcHyperScore (PSMMatch/HyperscoreMatch.java:61-187)
//   MVH         MVHMatch.calcMVH (PSMMatch/MVHMatch.java:121-302)
// and the per-spectrum loops around it (DBSearchWorker.java:25-113, StatisticScoreWorker.java:36-63,
// FindCandidateSpectraAndScoreWorker.java:33-100), whose counts are reduced inside the CTA.
//
// Why a thread per pair and not a warp per pair: the shuffles of one PSM (and the reference forms of one precursor
// window) have the same or similar lengths, so 32 pairs walk their fragment ladders in lock-step with no divergence,
// the b-before-y claim order and the sequential FP64 intensity sum of the reference fall out of program order, and no
// shuffle/ballot traffic is spent on 1-2 ions per lane.  See DESIGN.md §K1.
#include <cstdlib>

#include "k1_device.cuh"

#ifndef K1_MIN_BLOCKS
#define K1_MIN_BLOCKS 4
#endif

namespace pq {

namespace {

using namespace dev;

// bytes of the fixed part of K1's dynamic shared memory (row columns + tables [+ MVH used-bitmap columns]), 128-byte aligned
template <int BLOCK, int METHOD>
__host__ __device__ constexpr size_t kFixedSmem() {      // without the residue table, which follows it (its size is a launch parameter)
    return ((size_t)16 * BLOCK * 4 + (size_t)(2 * kMaxTermCodes + 130) * 8 + (METHOD == 2 ? (size_t)10 * BLOCK * 4 : 0) + 127) & ~(size_t)127;
}
// Copies of every residue-table entry: the Step-4 configuration (BLOCK 128 / 256) is bound by shared-memory bandwidth (ncu: 0.77
// wavefronts per cycle and SM, 41 % of them bank conflicts of the random 16-byte table look-ups); with 8 copies, copy
// lane & 7, a quarter-warp's look-ups are conflict-free.
template <int BLOCK> __host__ __device__ constexpr int kTabCopies() { return BLOCK >= 128 ? 8 : 1; }

template <int BLOCK, int METHOD>
__global__ void __launch_bounds__(BLOCK, BLOCK == 256 ? K1_MIN_BLOCKS : (BLOCK == 128 ? 2 * K1_MIN_BLOCKS : (BLOCK == 64 ? 11 : 1)))
k1_score_kernel(ScoreLaunch L, ScoreConfig cfg, uint32_t stage_bytes, int nstage, uint32_t tab_bytes, uint32_t n_codes) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ double s_red_d[BLOCK / 32];
    __shared__ uint32_t s_red_u[4][BLOCK / 32];

    // The per-thread row columns and the small tables sit at compile-time offsets from the start of dynamic shared memory (so
    // their addresses fold into the load instructions); the blob stages, whose size is a launch parameter, come after them.
    constexpr int TS = kTabCopies<BLOCK>();
    uint32_t* codes_all = reinterpret_cast<uint32_t*>(smem);                                            // [16][BLOCK]
    double* s_nt = reinterpret_cast<double*>(codes_all + 16 * BLOCK);
    double* s_ct = s_nt + kMaxTermCodes;
    double* s_l10 = s_ct + kMaxTermCodes;                                                              // [130]
    uint32_t* usedw_all = reinterpret_cast<uint32_t*>(s_l10 + 130);                                     // MVH only: [10][BLOCK]
    double2* s_tab_all = reinterpret_cast<double2*>(smem + kFixedSmem<BLOCK, METHOD>());               // {residue mass, modification mass} per code, TS copies each
    uint8_t* stage0 = smem + kFixedSmem<BLOCK, METHOD>() + tab_bytes;

    const int tid = threadIdx.x;
    for (int i = tid; i < (int)n_codes * TS; i += BLOCK) s_tab_all[i] = make_double2(L.mods->aa[i / TS], L.mods->delta[i / TS]);
    const double2* s_tab = s_tab_all + (TS > 1 ? (tid & (TS - 1)) : 0);                                // this lane's copy
    for (int i = tid; i < kMaxTermCodes; i += BLOCK) { s_nt[i] = L.mods->nterm[i]; s_ct[i] = L.mods->cterm[i]; }
    for (int i = tid; i < 130; i += BLOCK) s_l10[i] = L.log10_fact[i];     // [0..64] log10(n!), [65..129] log10(100 n)
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();

    const double itol = cfg.itol;
    const bool prefetch = cfg.prefetch_next != 0;
    uint32_t* codes = codes_all + tid;
    uint32_t* usedw = usedw_all + tid;

    // Jobs are handed out by an atomic counter (in list order, so neighbouring CTAs share table rows in L2); the next
    // job's blob is already in flight (second stage) while the current one is scored.
    __shared__ uint32_t s_job[2];
    constexpr uint32_t kSurvCap = BLOCK >= 128 ? 1024 : 2 * BLOCK;      // only the Step-4 configurations (BLOCK 128 / 256) pay for a long list
    __shared__ uint16_t s_surv[kSurvCap];         // bounded Step 4: candidates the score bound could not decide (index within the job)
    __shared__ __align__(16) uint8_t s_cflag[BLOCK >= 128 ? kCells + 16 : 16];      // Step-4 count pass: one byte per cell of the job's spectrum, 1 = nothing can be claimed (+ the list sentinel)
    __shared__ uint32_t s_nsurv;
    const bool bounded = L.kind == kJobShuffles && !cfg.score_all && !cfg.check_window;
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

        uint32_t n_in = 0, n_better = 0, n_bound = 0;
        double best = -1.0; uint32_t best_cand = 0xFFFFFFFFu;
        auto load_row = [&](uint32_t row) {
            const uint4* rp = reinterpret_cast<const uint4*>(L.rows + (size_t)row * kRowBytes);
            uint4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2), q3 = __ldg(rp + 3);
            codes[0 * BLOCK] = q0.x; codes[1 * BLOCK] = q0.y; codes[2 * BLOCK] = q0.z; codes[3 * BLOCK] = q0.w;
            codes[4 * BLOCK] = q1.x; codes[5 * BLOCK] = q1.y; codes[6 * BLOCK] = q1.z; codes[7 * BLOCK] = q1.w;
            codes[8 * BLOCK] = q2.x; codes[9 * BLOCK] = q2.y; codes[10 * BLOCK] = q2.z; codes[11 * BLOCK] = q2.w;
            codes[12 * BLOCK] = q3.x; codes[13 * BLOCK] = q3.y; codes[14 * BLOCK] = q3.z; codes[15 * BLOCK] = q3.w;
        };
        if (bounded) {
            // Step 4 only counts shuffles that reach the target's score (StatisticScoreWorker.java:51).  Count pass: every
            // shuffle's possible matches and the score bound they allow; the few shuffles the bound cannot decide (a few per
            // cent) are collected per CTA and scored in full, densely — the full scorer never runs on a warp with one live lane.
            Extra ex; ex.site = -1; ex.delta = 0.0;
            auto score_survivors = [&](uint32_t base, uint32_t first, uint32_t count) {
                if ((uint32_t)tid < count) {
                    const uint32_t row = J.cand_begin + base + s_surv[first + tid];
                    load_row(row);
                    int32_t pred = 0;
                    if (METHOD == 2) pred = L.row_pred[2 * (size_t)row + (J.charge == 2 ? 0 : 1)];
                    PairScore ps = score_pair<BLOCK, METHOD, TS>(codes, B, s_tab, s_nt, s_ct, J.charge, itol, ex, s_l10, L.ln_fact, pred, usedw, 0.0);
                    if (ps.score >= J.ref_score) ++n_better;
                }
            };
            // cell lists of the job's shuffle rows (written by K2 with the rows): the count pass needs no ladder walk and no row
            const uint4* job_lists = nullptr; size_t list_stride = 0; uint32_t list_chunks = 0, list_units = 0, list_n = 0;
            if (L.lists) {
                const ShuffleForm& F = L.sh_forms[L.form_slot[J.aux]];
                const int nch = J.charge <= 1 ? 1 : J.charge - 1;
                if (nch <= (int)F.list_nch) {
                    list_chunks = F.list_chunks; list_stride = F.list_rows; list_n = list_entries((int)F.len, nch);
                    list_units = (uint32_t)nch * list_stream_units((int)F.len);      // the first nch charge streams of each side
                    job_lists = L.lists + F.list_base;                     // unit u of row c of the job (= shuffle c of the form) at [u * list_stride + c]
                }
            }
            if (tid == 0) s_nsurv = 0;
            if (BLOCK >= 128 && job_lists) {
                // the clean flags of the staged cell table, one byte per cell: a list entry then costs a byte load and an add
                const uint4* src = reinterpret_cast<const uint4*>(B.cells);
                for (uint32_t q = tid; q < (uint32_t)kCells / 8; q += BLOCK) {
                    const uint4 v = src[q];
                    const uint32_t lo = ((v.x >> 15) & 1u) | ((v.x >> 23) & 0x100u) | (((v.y >> 15) & 1u) << 16) | (((v.y >> 31) & 1u) << 24);
                    const uint32_t hi = ((v.z >> 15) & 1u) | ((v.z >> 23) & 0x100u) | (((v.w >> 15) & 1u) << 16) | (((v.w >> 31) & 1u) << 24);
                    reinterpret_cast<uint2*>(s_cflag)[q] = make_uint2(lo, hi);
                }
                if (tid < 4) reinterpret_cast<uint32_t*>(s_cflag + kCells)[tid] = 0x01010101u;
            }
            __syncthreads();
            // jobs whose whole candidate list fits the survivor list (the usual case: -n 1000) run the count pass without a
            // barrier; longer jobs drain the list in rounds of BLOCK, 2048 candidates at a time
            for (uint32_t s0 = 0; s0 < J.cand_count; s0 += kSurvCap) {
                const uint32_t s1 = min(s0 + kSurvCap, J.cand_count);
                for (uint32_t c = s0 + tid; c < s1; c += BLOCK) {
                    ++n_in;
                    int32_t pred = 0;
                    if (METHOD == 2) pred = L.row_pred[2 * ((size_t)J.cand_begin + c) + (J.charge == 2 ? 0 : 1)];
                    bool below;
                    if (job_lists) below = below_score_bound_lists<METHOD>(job_lists + c, list_stride, list_chunks, list_units, list_n, B, smem_addr(s_cflag), s_l10, J.ref_score, L.ln_fact, pred);
                    else { load_row(J.cand_begin + c); below = below_score_bound<BLOCK, TS, METHOD>(codes, B, s_tab, s_nt, s_ct, J.charge, s_l10, J.ref_score, L.ln_fact, pred); }
                    if (below) ++n_bound;
                    else s_surv[atomicAdd(&s_nsurv, 1u)] = (uint16_t)(c - s0);
                }
                __syncthreads();
                const uint32_t ns = s_nsurv;
                for (uint32_t f = 0; f < ns; f += BLOCK) score_survivors(s0, f, min((uint32_t)BLOCK, ns - f));
                __syncthreads();
                if (tid == 0) s_nsurv = 0;
            }
        } else
        for (uint32_t c0 = 0; c0 < J.cand_count; c0 += BLOCK) {
            uint32_t c = c0 + tid;
            bool active = c < J.cand_count;
            uint32_t row = J.cand_begin + (active ? c : 0u);
            Extra ex; ex.site = -1; ex.delta = 0.0;
            if (active) {
                load_row(row);
                if (cfg.check_window) active = in_window(L.row_mass[row], J.mass, va, cfg);
                if (active && L.tgt.n && L.kind == kJobTargets) active = targeted_allows(L.tgt, row, J.spectrum);
            }
            if (active) {
                int32_t pred = 0;
                if (METHOD == 2) pred = L.row_pred[2 * (size_t)row + (J.charge == 2 ? 0 : 1)];
                PairScore ps = score_pair<BLOCK, METHOD, TS>(codes, B, s_tab, s_nt, s_ct, J.charge, itol, ex,
                                                         s_l10, L.ln_fact, pred, usedw, 0.0);
                ++n_in;
                if (ps.score > best) { best = ps.score; best_cand = row; }
                bool emit = false;
                switch (L.kind) {
                    case kJobTargets: emit = ps.score > cfg.min_score; break;                      // FindCandidate...java:78
                    case kJobCompetitors:
                        if (J.ref_score <= ps.score) { ++n_better; emit = ps.score >= 20.0; }      // DBSearchWorker.java:68-76
                        break;
                    case kJobShuffles: if (ps.score >= J.ref_score) ++n_better; break;             // StatisticScoreWorker.java:51
                    case kJobPairs:
                        L.pair_score[J.out + c] = ps.score; L.pair_a[J.out + c] = ps.a; L.pair_b[J.out + c] = ps.b;
                        break;
                    default: break;
                }
                if (emit) {
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
            n_bound += __shfl_xor_sync(0xffffffffu, n_bound, o);
            double ob = __shfl_xor_sync(0xffffffffu, best, o);
            uint32_t oc = __shfl_xor_sync(0xffffffffu, best_cand, o);
            if (ob > best || (ob == best && oc < best_cand)) { best = ob; best_cand = oc; }
        }
        if ((tid & 31) == 0) { s_red_u[0][tid >> 5] = n_in; s_red_u[1][tid >> 5] = n_better; s_red_u[2][tid >> 5] = best_cand; s_red_u[3][tid >> 5] = n_bound; s_red_d[tid >> 5] = best; }
        __syncthreads();     // also: every thread is done with this stage's blob
        if (tid == 0) {
            uint32_t a = 0, b = 0, nb = 0; double bs = -1.0; uint32_t bc = 0xFFFFFFFFu;
            for (int w = 0; w < BLOCK / 32; ++w) {
                a += s_red_u[0][w]; b += s_red_u[1][w]; nb += s_red_u[3][w];
                if (s_red_d[w] > bs || (s_red_d[w] == bs && s_red_u[2][w] < bc)) { bs = s_red_d[w]; bc = s_red_u[2][w]; }
            }
            JobResult R; R.n_in_window = a; R.n_better = b; R.best_score = bs; R.best_cand = bc; R.pad = nb;
            L.results[job] = R;
            if (nstage == 1) { if (prefetch) load(0, next_job); else fetch(0); }
        }
        __syncthreads();
        job = s_job[nstage == 2 ? (cur ^ 1u) : 0u];
    }
}

// Short jobs (Step 2: one or two target forms per spectrum) leave a CTA-per-job kernel with one busy lane per warp.
// Here every THREAD takes its own job and reads that spectrum's blob straight from global memory (L2): no staging, all
// lanes busy, latency hidden by occupancy.  Same pair scorer, same results.
template <int BLOCK, int METHOD>
__global__ void __launch_bounds__(BLOCK)      // (72 registers, 7 resident CTAs; forcing 10 CTAs at 48 registers measured 1.72 vs 1.57 ms)
k1_scatter_kernel(ScoreLaunch L, ScoreConfig cfg) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t* codes_all = reinterpret_cast<uint32_t*>(smem);                                            // [16][BLOCK]
    double2* s_tab = reinterpret_cast<double2*>(codes_all + 16 * BLOCK);
    double* s_nt = reinterpret_cast<double*>(s_tab + kMaxCodes);
    double* s_ct = s_nt + kMaxTermCodes;
    double* s_l10 = s_ct + kMaxTermCodes;                                                              // [130]
    uint32_t* usedw_all = reinterpret_cast<uint32_t*>(s_l10 + 130);                                     // MVH only: [10][BLOCK]
    const int tid = threadIdx.x;
    for (int i = tid; i < kMaxCodes; i += BLOCK) s_tab[i] = make_double2(L.mods->aa[i], L.mods->delta[i]);
    for (int i = tid; i < kMaxTermCodes; i += BLOCK) { s_nt[i] = L.mods->nterm[i]; s_ct[i] = L.mods->cterm[i]; }
    for (int i = tid; i < 130; i += BLOCK) s_l10[i] = L.log10_fact[i];
    __syncthreads();
    const double itol = cfg.itol;
    uint32_t* codes = codes_all + tid;
    uint32_t* usedw = usedw_all + tid;
    for (uint32_t job = blockIdx.x * BLOCK + tid; job < L.n_jobs; job += gridDim.x * BLOCK) {
        const Job J = L.jobs[job];
        const BlobView B = view_blob(L.blobs + L.blob_off[J.blob]);
        const long long va = jround(__dmul_rn(J.mass, 10.0));
        uint32_t n_in = 0, n_better = 0;
        double best = -1.0; uint32_t best_cand = 0xFFFFFFFFu;
        for (uint32_t c = 0; c < J.cand_count; ++c) {
            const uint32_t row = J.cand_begin + c;
            if (cfg.check_window && !in_window(L.row_mass[row], J.mass, va, cfg)) continue;
            if (L.tgt.n && L.kind == kJobTargets && !targeted_allows(L.tgt, row, J.spectrum)) continue;
            const uint4* rp = reinterpret_cast<const uint4*>(L.rows + (size_t)row * kRowBytes);
            uint4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2), q3 = __ldg(rp + 3);
            codes[0 * BLOCK] = q0.x; codes[1 * BLOCK] = q0.y; codes[2 * BLOCK] = q0.z; codes[3 * BLOCK] = q0.w;
            codes[4 * BLOCK] = q1.x; codes[5 * BLOCK] = q1.y; codes[6 * BLOCK] = q1.z; codes[7 * BLOCK] = q1.w;
            codes[8 * BLOCK] = q2.x; codes[9 * BLOCK] = q2.y; codes[10 * BLOCK] = q2.z; codes[11 * BLOCK] = q2.w;
            codes[12 * BLOCK] = q3.x; codes[13 * BLOCK] = q3.y; codes[14 * BLOCK] = q3.z; codes[15 * BLOCK] = q3.w;
            Extra ex; ex.site = -1; ex.delta = 0.0;
            int32_t pred = 0;
            if (METHOD == 2) pred = L.row_pred[2 * (size_t)row + (J.charge == 2 ? 0 : 1)];
            PairScore ps = score_pair<BLOCK, METHOD, 1, true>(codes, B, s_tab, s_nt, s_ct, J.charge, itol, ex, s_l10, L.ln_fact, pred, usedw, 0.0);      // ion filter from the blob's bit map
            ++n_in;
            if (ps.score > best) { best = ps.score; best_cand = row; }
            bool emit = false;
            switch (L.kind) {
                case kJobTargets: emit = ps.score > cfg.min_score; break;                      // FindCandidate...java:78
                case kJobCompetitors:
                    if (J.ref_score <= ps.score) { ++n_better; emit = ps.score >= 20.0; }      // DBSearchWorker.java:68-76
                    break;
                case kJobShuffles: if (ps.score >= J.ref_score) ++n_better; break;             // StatisticScoreWorker.java:51
                case kJobPairs:
                    L.pair_score[J.out + c] = ps.score; L.pair_a[J.out + c] = ps.a; L.pair_b[J.out + c] = ps.b;
                    break;
                default: break;
            }
            if (emit) {
                uint32_t slot = atomicAdd(L.hit_count, 1u);
                if (slot < L.hit_capacity) {
                    Hit h; h.spectrum = J.spectrum; h.cand = row; h.score = ps.score; h.mass = J.mass;
                    h.count_a = ps.a; h.count_b = ps.b; h.iso = J.iso; h.aux = J.aux; h.blob = J.blob; h.pad = 0;
                    L.hits[slot] = h;
                }
            }
        }
        JobResult R; R.n_in_window = n_in; R.n_better = n_better; R.best_score = best; R.best_cand = best_cand; R.pad = 0;
        L.results[job] = R;
    }
}

template <int METHOD>
int launch_scatter(const ScoreLaunch& L, const ScoreConfig& cfg, cudaStream_t st) {
    constexpr int BLOCK = 128;
    const size_t smem = (size_t)16 * BLOCK * 4 + (size_t)(2 * kMaxCodes + 2 * kMaxTermCodes + 130) * 8 + (METHOD == 2 ? (size_t)10 * BLOCK * 4 : 0);
    auto kernel = k1_scatter_kernel<BLOCK, METHOD>;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0; cudaGetDevice(&dev);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BLOCK, smem);
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (uint32_t)(sms * per_sm);
    uint32_t need = (L.n_jobs + BLOCK - 1) / BLOCK;
    if (grid > need) grid = need;
    kernel<<<grid, BLOCK, smem, st>>>(L, cfg);
    return 1;
}

template <int BLOCK, int METHOD>
int launch_kernel(const ScoreLaunch& L, const ScoreConfig& cfg, cudaStream_t st) {
    const uint32_t stage = (L.max_blob_bytes + 127u) & ~127u;
    const uint32_t n_codes = kTabCopies<BLOCK>() > 1 ? std::min<uint32_t>(std::max<uint32_t>(L.n_codes, 26u), (uint32_t)kMaxCodes) : (uint32_t)kMaxCodes;
    const uint32_t tab_bytes = (n_codes * (uint32_t)kTabCopies<BLOCK>() * 16u + 127u) & ~127u;
    const size_t tail = kFixedSmem<BLOCK, METHOD>() + tab_bytes;
    // 256-thread shuffle jobs (~1000 pairs each) hide the blob load behind a second stage; everywhere else one stage and more
    // resident CTAs win (measured: Step 2 9.1 -> 5.4 ms, Step 3 18.7 -> 13.4 ms, Step 4 with 128 threads 9.1 -> 7.85 ms).
    int nstage = BLOCK == 256 ? 2 : 1;
    if ((size_t)stage * 2 + tail > 100 * 1024) nstage = 1;
    if (const char* e = getenv("PQ_K1_STAGES")) nstage = atoi(e) == 1 ? 1 : (atoi(e) == 2 ? 2 : nstage);
    const size_t smem = (size_t)stage * nstage + tail;
    auto kernel = k1_score_kernel<BLOCK, METHOD>;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0; cudaGetDevice(&dev);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BLOCK, smem);      // resident CTAs: registers, shared memory, warps
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (uint32_t)(sms * per_sm);
    if (grid > L.n_jobs) grid = L.n_jobs;
    cudaMemsetAsync(L.job_counter, 0, 4, st);
    kernel<<<grid, BLOCK, smem, st>>>(L, cfg, stage, nstage, tab_bytes, n_codes);
    return 1;
}
template <int BLOCK>
int launch_block(const ScoreLaunch& L, const ScoreConfig& cfg, cudaStream_t st) {
    return cfg.method == 2 ? launch_kernel<BLOCK, 2>(L, cfg, st) : launch_kernel<BLOCK, 1>(L, cfg, st);
}

}  // namespace

int launch_score(const ScoreLaunch& L, const ScoreConfig& cfg, cudaStream_t st) {
    if (L.n_jobs == 0) return 0;
    switch (L.kind) {
        case kJobShuffles: {
            // CTAs of 128 threads, one blob stage: the ~40 undecided shuffles of a job keep two of FOUR warps busy instead of two of eight,
            // and seven CTAs per SM (shared memory) hide the blob load of the next job (C2 size: 96 threads 8.85 ms, 128 7.85, 160 8.1, 192 8.3, 256 with
            // two stages 8.6; PQ_STEP4_BLOCK=256 selects the latter)
            const char* e = getenv("PQ_STEP4_BLOCK");
            if (cfg.score_all && !e) return launch_block<256>(L, cfg, st);      // every shuffle scored in full: all warps busy anyway, two stages win (32.3 vs 35.0 ms)
            return e && atoi(e) == 256 ? launch_block<256>(L, cfg, st) : launch_block<128>(L, cfg, st);
        }
        case kJobCompetitors:
            if (L.lad && !getenv("PQ_STEP3_WALK")) return launch_score_ladder(L, cfg, st);
            return launch_block<64>(L, cfg, st);
        case kJobTargets:
            if (!getenv("PQ_STEP2_STAGED")) return cfg.method == 2 ? launch_scatter<2>(L, cfg, st) : launch_scatter<1>(L, cfg, st);
            return launch_block<32>(L, cfg, st);
        default: return launch_block<32>(L, cfg, st);
    }
}

}  // namespace pq

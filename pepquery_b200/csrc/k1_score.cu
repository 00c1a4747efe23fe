// k1_score.cu — K1: the pair scorer.  One CTA owns one job (= one prepared spectrum x a run of candidate peptide
// rows); the spectrum blob is staged into shared memory with one TMA bulk copy (cp.async.bulk + mbarrier,
// double-buffered across jobs); one THREAD scores one (peptide, spectrum) pair.
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
#include "pq_internal.h"

namespace pq {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct BlobView {
    const double* pmz; const uint32_t* pinfo; const double* val; const uint16_t* cells;
    uint32_t E; uint32_t nsurv; uint32_t flags; int32_t cls[4];
};
__device__ __forceinline__ BlobView view_blob(const uint8_t* b) {
    const BlobHeader* h = reinterpret_cast<const BlobHeader*>(b);
    BlobView v;
    v.E = h->n_peaks; v.nsurv = h->n_survivors; v.flags = h->flags;
    for (int c = 0; c < 4; ++c) v.cls[c] = h->cls_count[c];
    uint32_t ep = h->n_pad;
    v.pmz = reinterpret_cast<const double*>(b + sizeof(BlobHeader));
    v.pinfo = reinterpret_cast<const uint32_t*>(b + sizeof(BlobHeader) + (size_t)ep * 8);
    v.val = reinterpret_cast<const double*>(b + sizeof(BlobHeader) + (size_t)ep * 12);
    v.cells = reinterpret_cast<const uint16_t*>(b + sizeof(BlobHeader) + (size_t)ep * 12 + kValSlots * 8);
    return v;
}

// The annotator's peak pick for one theoretical m/z (A2-A4): most intense matchable peak with |mz - t| <= itol,
// equal intensity -> smaller |error|.  Returns pinfo of the winner, 0 if none; *mz_out = its m/z.
__device__ __forceinline__ uint32_t match_ion(const BlobView& B, double t, double itol, double itol_pad, double* mz_out) {
    double lo = __dsub_rn(t, itol_pad);
    uint32_t j = B.cells[cell_of(lo)];
    uint32_t best = 0; double best_abs = 0.0;
    for (; j < B.E; ++j) {
        double m = B.pmz[j];
        double err = __dsub_rn(m, t);
        if (err > itol) break;
        double ae = fabs(err);
        if (ae <= itol) {
            uint32_t info = B.pinfo[j];
            uint32_t r = info >> 16, br = best >> 16;
            if (r > br || (r == br && ae < best_abs)) { best = info; best_abs = ae; *mz_out = m; }
        }
    }
    return best;
}

struct Extra { int32_t site; double delta; };   // Step 5: one extra modification (site -1 = none)

struct PairScore { double score; int32_t a, b; };

template <int BLOCK>
__device__ __forceinline__ uint32_t row_byte(const uint32_t* codes, int idx) {
    uint32_t w = codes[(idx >> 2) * BLOCK];
    return (w >> ((idx & 3) * 8)) & 0xFFu;
}

// Walks the fragment ladder of one peptide row against the staged spectrum.
// METHOD 1: hyperscore (count_b, count_y, sum of normalised intensities); METHOD 2: MVH class hits.
template <int BLOCK, int METHOD>
__device__ __forceinline__ PairScore score_pair(const uint32_t* codes /* this thread's column */, const BlobView& B,
                                                const double* aa, const double* dl, const double* ntd, const double* ctd,
                                                int z, double itol, double itol_pad, Extra ex,
                                                const double* log10_fact, const double* ln_fact, int32_t predicted,
                                                uint32_t* usedw /* MVH: this thread's column of the used bitmap */) {
    const uint32_t head = codes[0];
    const int L = (int)(head & 0xFF);
    const int nch = z <= 1 ? 1 : z - 1;          // JMatch.java:398-407
    uint64_t used = 0ull;
    int ca = 0, cb = 0;
    int key0 = 0, key1 = 0, key2 = 0;
    double sum = 0.0;
    if (METHOD == 2) for (int w = 0; w < 10; ++w) usedw[w * BLOCK] = 0u;

    auto claim = [&](uint32_t info, double mz, bool is_b) {
        uint32_t sid = info & 0xFFFu;
        if (info == 0u || sid == kNoSurvivor) return;
        if (METHOD == 1) {
            uint64_t bit = 1ull << sid;
            if (used & bit) return;
            used |= bit;
            if (is_b) ++ca; else ++cb;
            sum = __dadd_rn(sum, B.val[sid]);
        } else {
            if (mz < 200.0 || mz > 2000.0) return;                    // MVHMatch.java:248-250
            uint32_t w = usedw[(sid >> 5) * BLOCK], bit = 1u << (sid & 31);
            if (w & bit) return;
            usedw[(sid >> 5) * BLOCK] = w | bit;
            uint32_t c = (info >> 12) & 3u;
            key0 += c == 0; key1 += c == 1; key2 += c == 2;
        }
    };

    // ---- b ions: forward sum, residue mass then its modification, N-term modification opens the sum ----
    double acc = ntd[(head >> 8) & 0xFF];
    if (ex.site == 0) acc = __dadd_rn(acc, ex.delta);
    for (int k = 1; k <= L - 1; ++k) {
        uint32_t code = row_byte<BLOCK>(codes, kRowHeader + k - 1);
        acc = __dadd_rn(acc, aa[code]);
        double d = dl[code];
        if (d != 0.0) acc = __dadd_rn(acc, d);
        if (k == ex.site) acc = __dadd_rn(acc, ex.delta);
        for (int c = 1; c <= nch; ++c) {
            if (c > 1 && c > k) break;                                 // A6 chargeValidated
            double t = __ddiv_rn(__dadd_rn(acc, __dmul_rn((double)c, kProton)), (double)c);
            double mz = 0.0;
            uint32_t info = match_ion(B, t, itol, itol_pad, &mz);
            claim(info, mz, true);
        }
    }
    // ---- y ions: rewind sum from the C-terminus, + H2O ----
    acc = ctd[(head >> 16) & 0xFF];
    if (ex.site == L + 1) acc = __dadd_rn(acc, ex.delta);
    for (int k = 1; k <= L - 1; ++k) {
        int site = L - k + 1;
        uint32_t code = row_byte<BLOCK>(codes, kRowHeader + site - 1);
        acc = __dadd_rn(acc, aa[code]);
        double d = dl[code];
        if (d != 0.0) acc = __dadd_rn(acc, d);
        if (site == ex.site) acc = __dadd_rn(acc, ex.delta);
        double ym = __dadd_rn(acc, kWater);
        for (int c = 1; c <= nch; ++c) {
            if (c > 1 && c > k) break;
            double t = __ddiv_rn(__dadd_rn(ym, __dmul_rn((double)c, kProton)), (double)c);
            double mz = 0.0;
            uint32_t info = match_ion(B, t, itol, itol_pad, &mz);
            claim(info, mz, false);
        }
    }

    PairScore r;
    if (METHOD == 1) {
        if (ca > 64) ca = 64;
        if (cb > 64) cb = 64;
        r.a = ca; r.b = cb;
        if (ca == 0 && cb == 0) { r.score = 0.0; return r; }
        double hs = jlog10(sum);
        if (ca > 0 && cb > 0) hs = __dadd_rn(__dadd_rn(hs, log10_fact[ca]), log10_fact[cb]);
        else if (cb > 0) hs = __dadd_rn(hs, log10_fact[cb]);
        else hs = __dadd_rn(hs, log10_fact[ca]);
        r.score = __dmul_rn(4.0, hs);
    } else {
        int matched = key0 + key1 + key2;
        r.a = matched; r.b = predicted;
        r.score = 0.0;
        if (!(B.flags & 1u)) { r.a = 0; r.b = 0; return r; }     // fewer than 7 peaks survived: calcMVH returns before scoring
        if (predicted <= 0) return r;
        int key[4] = {key0, key1, key2, predicted - matched};
        int N = (int)B.nsurv;
        double mvh = 0.0; bool died = false;
        for (int i = 0; i < 4; ++i) {
            int n = B.cls[i], k = key[i];
            if (n < 0 || k < 0) continue;
            if (n < k) { died = true; continue; }
            mvh = __dadd_rn(mvh, __dsub_rn(__dsub_rn(ln_fact[n], ln_fact[n - k]), ln_fact[k]));
        }
        int tb = B.cls[3] + N;
        if (tb >= 0 && predicted >= 0) {
            if (tb < predicted) died = true;
            else mvh = __dsub_rn(mvh, __dsub_rn(__dsub_rn(ln_fact[tb], ln_fact[tb - predicted]), ln_fact[predicted]));
        }
        mvh = __dsub_rn(0.0, mvh);
        r.score = died ? 0.0 : mvh;
    }
    return r;
}

// Precursor-window predicate exactly as DBSearchWorker.java:38-52 / FindCandidate...java:47-58:
// bucket round(10*pepMass) within +-1 of round(10*mass) AND |pepMass - mass| (ppm of pepMass) <= tol.
__device__ __forceinline__ bool in_window(double pm, double mass, long long va, const ScoreConfig& cfg) {
    long long key = jround(__dmul_rn(pm, 10.0));
    if (key < va - 1 || key > va + 1) return false;
    double del = fabs(__dsub_rn(pm, mass));
    if (cfg.tol_ppm) del = __ddiv_rn(__dmul_rn(__dmul_rn(1.0e6, 1.0), del), pm);
    return del <= cfg.tol;
}

template <int BLOCK, int METHOD>
__global__ void __launch_bounds__(BLOCK)
k1_score_kernel(ScoreLaunch L, ScoreConfig cfg, uint32_t stage_bytes, int nstage) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ double s_red_d[BLOCK / 32];
    __shared__ uint32_t s_red_u[3][BLOCK / 32];

    uint8_t* stage0 = smem;
    uint32_t* codes_all = reinterpret_cast<uint32_t*>(smem + (size_t)stage_bytes * nstage);           // [16][BLOCK]
    double* s_aa = reinterpret_cast<double*>(codes_all + 16 * BLOCK);
    double* s_dl = s_aa + kMaxCodes;
    double* s_nt = s_dl + kMaxCodes;
    double* s_ct = s_nt + kMaxTermCodes;
    double* s_l10 = s_ct + kMaxTermCodes;                                                              // [65]
    uint32_t* usedw_all = reinterpret_cast<uint32_t*>(s_l10 + 66);                                     // MVH [10][BLOCK]

    const int tid = threadIdx.x;
    for (int i = tid; i < kMaxCodes; i += BLOCK) { s_aa[i] = L.mods->aa[i]; s_dl[i] = L.mods->delta[i]; }
    for (int i = tid; i < kMaxTermCodes; i += BLOCK) { s_nt[i] = L.mods->nterm[i]; s_ct[i] = L.mods->cterm[i]; }
    for (int i = tid; i < 65; i += BLOCK) s_l10[i] = L.log10_fact[i];
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();

    const double itol = cfg.itol;
    const double itol_pad = itol + 1e-6;
    uint32_t* codes = codes_all + tid;
    uint32_t* usedw = usedw_all + tid;

    uint32_t job = blockIdx.x;
    uint32_t it = 0;
    // prologue: stage the first job's blob
    if (job < L.n_jobs && tid == 0) {
        const Job& J = L.jobs[job];
        uint32_t bytes = (uint32_t)(L.blob_off[J.blob + 1] - L.blob_off[J.blob]);
        mbar_expect_tx(&s_bar[0], bytes);
        tma_load_1d(stage0, L.blobs + L.blob_off[J.blob], bytes, &s_bar[0]);
    }
    for (; job < L.n_jobs; job += gridDim.x, ++it) {
        const uint32_t cur = nstage == 2 ? (it & 1u) : 0u;
        const uint32_t parity = nstage == 2 ? ((it >> 1) & 1u) : (it & 1u);
        // prefetch the next job's blob into the other stage
        if (nstage == 2) {
            uint32_t nj = job + gridDim.x;
            if (nj < L.n_jobs && tid == 0) {
                const Job& Jn = L.jobs[nj];
                uint32_t bytes = (uint32_t)(L.blob_off[Jn.blob + 1] - L.blob_off[Jn.blob]);
                mbar_expect_tx(&s_bar[cur ^ 1u], bytes);
                tma_load_1d(stage0 + (size_t)stage_bytes * (cur ^ 1u), L.blobs + L.blob_off[Jn.blob], bytes, &s_bar[cur ^ 1u]);
            }
        }
        const Job J = L.jobs[job];
        mbar_wait(&s_bar[cur], parity);
        const BlobView B = view_blob(stage0 + (size_t)stage_bytes * cur);
        const long long va = jround(__dmul_rn(J.mass, 10.0));

        uint32_t n_in = 0, n_better = 0;
        double best = -1.0; uint32_t best_cand = 0xFFFFFFFFu;
        for (uint32_t c0 = 0; c0 < J.cand_count; c0 += BLOCK) {
            uint32_t c = c0 + tid;
            bool active = c < J.cand_count;
            uint32_t row = J.cand_begin + (active ? c : 0u);
            Extra ex; ex.site = -1; ex.delta = 0.0;
            if (active) {
                const uint4* rp = reinterpret_cast<const uint4*>(L.rows + (size_t)row * kRowBytes);
                uint4 q0 = __ldg(rp), q1 = __ldg(rp + 1), q2 = __ldg(rp + 2), q3 = __ldg(rp + 3);
                codes[0 * BLOCK] = q0.x; codes[1 * BLOCK] = q0.y; codes[2 * BLOCK] = q0.z; codes[3 * BLOCK] = q0.w;
                codes[4 * BLOCK] = q1.x; codes[5 * BLOCK] = q1.y; codes[6 * BLOCK] = q1.z; codes[7 * BLOCK] = q1.w;
                codes[8 * BLOCK] = q2.x; codes[9 * BLOCK] = q2.y; codes[10 * BLOCK] = q2.z; codes[11 * BLOCK] = q2.w;
                codes[12 * BLOCK] = q3.x; codes[13 * BLOCK] = q3.y; codes[14 * BLOCK] = q3.z; codes[15 * BLOCK] = q3.w;
                if (cfg.check_window) active = in_window(L.row_mass[row], J.mass, va, cfg);
            }
            if (active) {
                int32_t pred = 0;
                if (METHOD == 2) pred = L.row_pred[2 * (size_t)row + (J.charge == 2 ? 0 : 1)];
                PairScore ps = score_pair<BLOCK, METHOD>(codes, B, s_aa, s_dl, s_nt, s_ct, J.charge, itol, itol_pad, ex,
                                                         s_l10, L.ln_fact, pred, usedw);
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
            if (nstage == 1) {
                uint32_t nj = job + gridDim.x;
                if (nj < L.n_jobs) {
                    const Job& Jn = L.jobs[nj];
                    uint32_t bytes = (uint32_t)(L.blob_off[Jn.blob + 1] - L.blob_off[Jn.blob]);
                    mbar_expect_tx(&s_bar[0], bytes);
                    tma_load_1d(stage0, L.blobs + L.blob_off[Jn.blob], bytes, &s_bar[0]);
                }
            }
        }
        __syncthreads();
    }
}

template <int BLOCK>
int launch_block(const ScoreLaunch& L, const ScoreConfig& cfg, cudaStream_t st) {
    uint32_t stage = (L.max_blob_bytes + 127u) & ~127u;
    size_t tail = (size_t)16 * BLOCK * 4 + (size_t)(2 * kMaxCodes + 2 * kMaxTermCodes + 66) * 8 + (size_t)10 * BLOCK * 4;
    int nstage = 2;
    if ((size_t)stage * 2 + tail > 100 * 1024) nstage = 1;
    size_t smem = (size_t)stage * nstage + tail;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int per_sm = (int)((220 * 1024) / (smem + 2048));
    int by_threads = 2048 / BLOCK;
    if (per_sm > by_threads) per_sm = by_threads;
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (uint32_t)(sms * per_sm);
    if (grid > L.n_jobs) grid = L.n_jobs;
    if (cfg.method == 2) {
        cudaFuncSetAttribute(k1_score_kernel<BLOCK, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k1_score_kernel<BLOCK, 2><<<grid, BLOCK, smem, st>>>(L, cfg, stage, nstage);
    } else {
        cudaFuncSetAttribute(k1_score_kernel<BLOCK, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k1_score_kernel<BLOCK, 1><<<grid, BLOCK, smem, st>>>(L, cfg, stage, nstage);
    }
    return 1;
}

}  // namespace

int launch_score(const ScoreLaunch& L, const ScoreConfig& cfg, cudaStream_t st) {
    if (L.n_jobs == 0) return 0;
    switch (L.kind) {
        case kJobShuffles: return launch_block<256>(L, cfg, st);
        case kJobCompetitors: return launch_block<64>(L, cfg, st);
        default: return launch_block<32>(L, cfg, st);
    }
}

}  // namespace pq

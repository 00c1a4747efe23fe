/*
 * pq_synth.cpp — deterministic synthetic inputs of the shape SURVEY.md §8d names (test + bench infrastructure).
 *
 *   proteins : lengths ~ lognormal(ln 400, 0.5) clipped [50,3000], residues i.i.d. UniProt-like frequencies
 *   targets  : "novel" tryptic-like peptides, length U[8,25], last residue K/R
 *   spectra  : HCD-like; 30 % from a target form, 30 % from a database peptide form, 40 % from a decoy
 *              (C-term-preserving shuffle of a database peptide); charge 2/3/4 = 0.6/0.35/0.05;
 *              precursor m/z with N(0,3 ppm) error; b/y peaks at c=1 (and c=2 for z>=3) kept w.p. 0.7,
 *              m/z jitter N(0, itol/6), intensity lognormal(8,1); Poisson(80)-like noise peaks with
 *              lognormal(6,1) intensity; 4-decimal strictly ascending m/z, 2-decimal intensity; >= 11 peaks.
 *
 * Every spectrum draws from its own counter-based RNG stream (seed, index), so the output does not depend on
 * the thread count, and a mass-range shard (shard s of S) can be produced without generating the others.
 * Masses here only need to be close (peaks are rounded to 4 decimals); this file shares no code with the
 * product or the oracle.
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t& x) { uint64_t z = (x += 0x9e3779b97f4a7c15ULL); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; return z ^ (z >> 31); }
    Rng(uint64_t seed, uint64_t stream) { uint64_t x = seed * 0x2545F4914F6CDD1DULL + stream * 0x9E3779B97F4A7C15ULL + 1; for (int i = 0; i < 4; ++i) s[i] = splitmix(x); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() { uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17; s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45); return r; }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    int below(int n) { return (int)(uni() * n); }
    double normal() { double u1 = uni(), u2 = uni(); if (u1 < 1e-300) u1 = 1e-300; return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2); }
};

const char AA[] = "ACDEFGHIKLMNPQRSTVWY";
const double FREQ[] = {8.25, 1.37, 5.45, 6.75, 3.86, 7.07, 2.27, 5.96, 5.84, 9.66, 2.42, 4.06, 4.70, 3.93, 5.53, 6.56, 5.34, 6.87, 1.08, 2.92};
double CUM[20];
double MASS[128];
const double PROTON = 1.007276466812, WATER = 18.0105646837;
struct Init {
    Init() {
        double t = 0; for (int i = 0; i < 20; ++i) t += FREQ[i];
        double c = 0; for (int i = 0; i < 20; ++i) { c += FREQ[i] / t; CUM[i] = c; }
        CUM[19] = 1.0;
        for (int i = 0; i < 128; ++i) MASS[i] = 0;
        MASS['G'] = 57.02146372057; MASS['A'] = 71.03711378471; MASS['S'] = 87.03202840427; MASS['P'] = 97.05276384885;
        MASS['V'] = 99.06841391299; MASS['T'] = 101.04767846841; MASS['C'] = 103.00918478471; MASS['L'] = 113.08406397713;
        MASS['I'] = 113.08406397713; MASS['N'] = 114.04292744114; MASS['D'] = 115.02694302383; MASS['Q'] = 128.05857750528;
        MASS['K'] = 128.094963014; MASS['E'] = 129.04259308797; MASS['M'] = 131.04048491299; MASS['H'] = 137.05891185845;
        MASS['F'] = 147.06841391299; MASS['R'] = 156.1011110236; MASS['Y'] = 163.06332853255; MASS['W'] = 186.07931294986;
    }
} init_;
char draw_aa(Rng& r) { double u = r.uni(); for (int i = 0; i < 20; ++i) if (u < CUM[i]) return AA[i]; return 'Y'; }

struct Source { std::string seq; std::vector<double> delta; };   // per-residue modification mass

// a random database peptide: tryptic, 0-1 missed cleavages, length 7..30, I->L as the digest does
bool db_peptide(Rng& r, int n_prot, const uint64_t* poff, const char* pres, std::string& out) {
    for (int attempt = 0; attempt < 64; ++attempt) {
        int p = r.below(n_prot);
        const char* s = pres + poff[p]; int n = (int)(poff[p + 1] - poff[p]);
        if (n < 20) continue;
        int a = r.below(n - 7);
        // move a to the next cleavage boundary
        while (a > 0 && a < n && !((s[a - 1] == 'K' || s[a - 1] == 'R') && s[a] != 'P')) ++a;
        if (a >= n - 6) continue;
        int missed = r.below(2);
        int e = a;
        for (int m = 0; m <= missed; ++m) {
            ++e;
            while (e < n && !((s[e - 1] == 'K' || s[e - 1] == 'R') && (e == n || s[e] != 'P'))) ++e;
            if (e >= n) { e = n; break; }
        }
        int len = e - a;
        if (len < 7 || len > 30) continue;
        out.assign(s + a, s + e);
        for (auto& c : out) if (c == 'I') c = 'L';
        return true;
    }
    return false;
}
void apply_mods(Rng& r, Source& src) {
    src.delta.assign(src.seq.size(), 0.0);
    int nox = 0;
    for (size_t i = 0; i < src.seq.size(); ++i) {
        if (src.seq[i] == 'C') src.delta[i] = 57.02146372057;
        else if (src.seq[i] == 'M' && nox < 3 && r.uni() < 0.2) { src.delta[i] = 15.99491461956; ++nox; }
    }
}
double neutral_mass(const Source& s) { double m = WATER; for (size_t i = 0; i < s.seq.size(); ++i) m += MASS[(int)s.seq[i]] + s.delta[i]; return m; }

struct SpecOut { double prec_mz; int charge; std::vector<std::pair<double, double>> peaks; };

bool make_source(Rng& r, int n_targets, const uint32_t* toff, const char* tres, int n_prot, const uint64_t* poff, const char* pres, Source& src) {
    double u = r.uni();
    if (u < 0.3 && n_targets > 0) {
        int t = r.below(n_targets);
        src.seq.assign(tres + toff[t], tres + toff[t + 1]);
    } else {
        if (!db_peptide(r, n_prot, poff, pres, src.seq)) return false;
        if (u >= 0.6) {   // decoy: shuffle all but the C-terminal residue
            for (int i = (int)src.seq.size() - 2; i > 0; --i) { int j = r.below(i + 1); std::swap(src.seq[i], src.seq[j]); }
        }
    }
    apply_mods(r, src);
    return true;
}

// Optional shard rule for the multi-GPU bench: the precursor-mass axis is cut into strips (quantile bounds) and shard
// `g_rank` of `g_world` owns every g_world-th strip (interleaved contiguous mass ranges of equal expected work).
std::vector<double> g_strips; int g_world = 1, g_rank = 0;
bool in_shard(double M, double mlo, double mhi) {
    if (g_strips.empty() || g_world <= 1) return M >= mlo && M < mhi;
    int strip = (int)(std::upper_bound(g_strips.begin(), g_strips.end(), M) - g_strips.begin()) - 1;
    if (strip < 0) strip = 0;
    return strip % g_world == g_rank;
}

void make_spectrum(uint64_t seed, int64_t index, int n_targets, const uint32_t* toff, const char* tres, int n_prot,
                   const uint64_t* poff, const char* pres, double itol, double mlo, double mhi, SpecOut& o) {
    Rng r(seed, (uint64_t)index);
    Source src;
    double M = 0;
    for (int attempt = 0;; ++attempt) {
        if (!make_source(r, n_targets, toff, tres, n_prot, poff, pres, src)) { src.seq = "PEPTLDEK"; apply_mods(r, src); }
        M = neutral_mass(src);
        if (in_shard(M, mlo, mhi) || attempt > 4000) break;
    }
    double uz = r.uni();
    int z = uz < 0.6 ? 2 : (uz < 0.95 ? 3 : 4);
    double mz0 = (M + z * PROTON) / z;
    o.charge = z;
    o.prec_mz = std::round(mz0 * (1.0 + r.normal() * 3e-6) * 1e5) / 1e5;
    int L = (int)src.seq.size();
    std::vector<std::pair<double, double>> pk;
    double fwd = 0, rew = WATER;
    for (int k = 1; k <= L - 1; ++k) {
        fwd += MASS[(int)src.seq[k - 1]] + src.delta[k - 1];
        rew += MASS[(int)src.seq[L - k]] + src.delta[L - k];
        for (int type = 0; type < 2; ++type) {
            double m = type == 0 ? fwd : rew;
            for (int c = 1; c <= (z >= 3 ? 2 : 1); ++c) {
                if (r.uni() < 0.7) {
                    double mz = (m + c * PROTON) / c + r.normal() * itol / 6.0;
                    double in = std::exp(8.0 + r.normal());
                    if (mz > 50.0) pk.push_back({mz, in});
                }
            }
        }
    }
    int nn = (int)std::lround(80.0 + std::sqrt(80.0) * r.normal()); if (nn < 0) nn = 0;
    double hi = std::min(M, 2000.0);
    for (int i = 0; i < nn; ++i) pk.push_back({100.0 + r.uni() * (hi - 100.0), std::exp(6.0 + r.normal())});
    while ((int)pk.size() < 11) pk.push_back({100.0 + r.uni() * (hi - 100.0), std::exp(6.0 + r.normal())});
    for (auto& p : pk) { p.first = std::round(p.first * 1e4) / 1e4; p.second = std::max(0.01, std::round(p.second * 1e2) / 1e2); }
    std::sort(pk.begin(), pk.end());
    o.peaks.clear();
    for (auto& p : pk) if (o.peaks.empty() || p.first > o.peaks.back().first) o.peaks.push_back(p);
    // de-duplication may drop below 11 peaks only in degenerate cases; pad deterministically above the top
    while ((int)o.peaks.size() < 11) o.peaks.push_back({std::round((o.peaks.back().first + 1.2345) * 1e4) / 1e4, 100.0});
}

}  // namespace

extern "C" {

// Proteins.  Call with res == NULL to get the residue count; off has n+1 entries.
int64_t syn_proteins(int32_t n, uint64_t seed, uint64_t* off, char* res, int64_t cap) {
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
        Rng r(seed, (uint64_t)i);
        int len = (int)std::lround(std::exp(std::log(400.0) + 0.5 * r.normal()));
        len = std::max(50, std::min(3000, len));
        if (off) off[i] = (uint64_t)total;
        if (res) { if (total + len > cap) return -1; for (int k = 0; k < len; ++k) res[total + k] = draw_aa(r); }
        total += len;
    }
    if (off) off[n] = (uint64_t)total;
    return total;
}

int64_t syn_targets(int32_t n, uint64_t seed, uint32_t* off, char* res, int64_t cap) {
    int64_t total = 0;
    for (int i = 0; i < n; ++i) {
        Rng r(seed, (uint64_t)i);
        int len = 8 + r.below(18);
        if (off) off[i] = (uint32_t)total;
        if (res) {
            if (total + len > cap) return -1;
            for (int k = 0; k < len - 1; ++k) { char c = draw_aa(r); if (c == 'I') c = 'L'; res[total + k] = c; }
            res[total + len - 1] = r.uni() < 0.5 ? 'K' : 'R';
        }
        total += len;
    }
    if (off) off[n] = (uint32_t)total;
    return total;
}

// Shard rule of the following syn_spectra calls: n_bounds strip boundaries (ascending), this process owns strips
// rank, rank + world, ...  n_bounds == 0 switches back to the plain [mass_lo, mass_hi) rule.
void syn_set_strips(int32_t n_bounds, const double* bounds, int32_t world, int32_t rank) {
    g_strips.assign(bounds, bounds + (n_bounds > 0 ? n_bounds : 0)); g_world = world; g_rank = rank;
}

// Mass quantile boundaries for S shards (S+1 doubles), estimated from a fixed sample of sources.
int32_t syn_mass_bounds(uint64_t seed, int32_t shards, int32_t n_targets, const uint32_t* toff, const char* tres,
                        int32_t n_prot, const uint64_t* poff, const char* pres, double* bounds) {
    const int N = 200000;
    std::vector<double> m; m.reserve(N);
    for (int i = 0; i < N; ++i) {
        Rng r(seed ^ 0xABCDEF12345ULL, (uint64_t)i);
        Source s;
        if (!make_source(r, n_targets, toff, tres, n_prot, poff, pres, s)) continue;
        m.push_back(neutral_mass(s));
    }
    std::sort(m.begin(), m.end());
    bounds[0] = 0.0; bounds[shards] = 1e9;
    for (int s = 1; s < shards; ++s) bounds[s] = m[(size_t)((double)m.size() * s / shards)];
    return 0;
}

// Spectra.  Pass 1 (mz == NULL): fills prec_mz/charge/peak_off and returns the total peak count.
// Pass 2: fills mz/intensity.  mass_lo/mass_hi restrict the source peptide's neutral mass (a shard).
int64_t syn_spectra(int64_t n, uint64_t seed, int64_t first_index, double mass_lo, double mass_hi,
                    int32_t n_targets, const uint32_t* toff, const char* tres,
                    int32_t n_prot, const uint64_t* poff, const char* pres, double itol,
                    double* prec_mz, int32_t* charge, uint64_t* peak_off, double* mz, double* inten, int32_t threads) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    std::vector<uint32_t> cnt;
    if (!mz) cnt.assign((size_t)n, 0);
    std::atomic<int64_t> next{0};
    auto work = [&]() {
        SpecOut o;
        for (;;) {
            int64_t b = next.fetch_add(256);
            if (b >= n) break;
            int64_t e = std::min(n, b + 256);
            for (int64_t i = b; i < e; ++i) {
                make_spectrum(seed, first_index + i, n_targets, toff, tres, n_prot, poff, pres, itol, mass_lo, mass_hi, o);
                if (!mz) { cnt[(size_t)i] = (uint32_t)o.peaks.size(); prec_mz[i] = o.prec_mz; charge[i] = o.charge; }
                else {
                    uint64_t base = peak_off[i];
                    for (size_t k = 0; k < o.peaks.size(); ++k) { mz[base + k] = o.peaks[k].first; inten[base + k] = o.peaks[k].second; }
                }
            }
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(work);
    for (auto& t : th) t.join();
    if (!mz) {
        uint64_t tot = 0;
        for (int64_t i = 0; i < n; ++i) { peak_off[i] = tot; tot += cnt[(size_t)i]; }
        peak_off[n] = tot;
        return (int64_t)tot;
    }
    return (int64_t)peak_off[n];
}


// MGF text of a CSR spectrum set in the reference's own form (index/IndexWorker.asMgf, IndexWorker.java:431-457): one block per
// spectrum, peaks as "%.4f %.2f".  Built by `threads` threads into a buffer the library owns: syn_mgf_build returns its size,
// syn_mgf_copy copies it out and frees it.
static std::vector<std::string> g_mgf_parts;
int64_t syn_mgf_build(int64_t n, const double* prec_mz, const int32_t* charge, const uint64_t* off, const double* mz, const double* inten, int32_t threads) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads <= 0) threads = 1;
    g_mgf_parts.assign((size_t)threads, std::string());
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back([&, t]() {
        std::string& out = g_mgf_parts[(size_t)t];
        const int64_t a = n * t / threads, b = n * (t + 1) / threads;
        out.reserve((size_t)((off[b] - off[a]) * 18 + (uint64_t)(b - a) * 96));
        char buf[128];
        for (int64_t i = a; i < b; ++i) {
            int k = snprintf(buf, sizeof buf, "BEGIN IONS\nTITLE=syn:%lld:%d\nPEPMASS=%.5f 1.0\n", (long long)(i + 1), (int)charge[i], prec_mz[i]);
            out.append(buf, (size_t)k);
            if (charge[i] > 0) { k = snprintf(buf, sizeof buf, "CHARGE=%d+\n", (int)charge[i]); out.append(buf, (size_t)k); }
            for (uint64_t p = off[i]; p < off[i + 1]; ++p) { k = snprintf(buf, sizeof buf, "%.4f %.2f\n", mz[p], inten[p]); out.append(buf, (size_t)k); }
            out.append("END IONS\n\n");
        }
    });
    for (auto& t : th) t.join();
    int64_t total = 0;
    for (auto& s2 : g_mgf_parts) total += (int64_t)s2.size();
    return total;
}
int64_t syn_mgf_copy(char* dst, int64_t cap) {
    int64_t w = 0;
    for (auto& s2 : g_mgf_parts) { if (w + (int64_t)s2.size() > cap) return -1; memcpy(dst + w, s2.data(), s2.size()); w += (int64_t)s2.size(); }
    g_mgf_parts.clear(); g_mgf_parts.shrink_to_fit();
    return w;
}

}  // extern "C"

/*
 * pq_oracle.cpp — CPU restatement of PepQuery's pair-scoring path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load it.  The product (pepquery_b200/csrc) never links it.
 *
 * PARITY UNPINNED at the compomics boundary: the reference is Java + the un-vendored jar
 * com.compomics:utilities:5.0.39 (pom.xml:18-31); it ships no tests or golden vectors and no JVM exists in
 * the build image.  What IS pinned: JDK-specified java.util.Random, the formulae/operators in the Java
 * source, the modification masses in resources/top_modifications.tsv, and the worked examples in
 * SURVEY.md §8c.  Every compomics-dependent assumption (A1..A12 of SURVEY.md §8c) is a named switch in
 * struct Assumptions so it can be flipped if a JVM ever shows otherwise.
 *
 * Structure deliberately follows the reference "as written": scorePeptide2Spectrum() re-does the spectrum
 * pre-processing for every pair, with list-of-peak objects, stable sorts and hash maps, exactly like
 * pg/PeptideSearchMT.java:1144-1251.  File:line citations are relative to /root/reference/src/main/java/.
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/pepquery_b200.h"

namespace orc {

// ---------------------------------------------------------------------------------------------------
// Assumption switches
// ---------------------------------------------------------------------------------------------------
struct Assumptions {
    double proton = 1.007276466812;          // A1 ElementaryIon.proton
    bool a2_percentile_interp = true;        // A2 linear-interpolated 5th percentile, >= limit is matchable
    bool a3_closed_interval = true;          // A3 |mz - theo| <= itol
    bool a4_tie_smaller_error = true;        // A4 equal intensity -> smaller |error|
    bool a5_b_before_y = true;               // A5 annotator order: all b (number asc, charge asc) then all y
    bool a8_mod_losses_in_prediction = true; // A8 IonFactory emits single mod-defined-loss ions
};
static Assumptions g_a;

// Residue masses (back-solved from resources/top_modifications.tsv atom masses).
static double g_res_mass[128];
static const double H2O = 18.0105646837;
static void init_masses() {
    for (int i = 0; i < 128; ++i) g_res_mass[i] = 0.0;
    g_res_mass['G'] = 57.02146372057;  g_res_mass['A'] = 71.03711378471;  g_res_mass['S'] = 87.03202840427;
    g_res_mass['P'] = 97.05276384885;  g_res_mass['V'] = 99.06841391299;  g_res_mass['T'] = 101.04767846841;
    g_res_mass['C'] = 103.00918478471; g_res_mass['L'] = 113.08406397713; g_res_mass['I'] = 113.08406397713;
    g_res_mass['N'] = 114.04292744114; g_res_mass['D'] = 115.02694302383; g_res_mass['Q'] = 128.05857750528;
    g_res_mass['K'] = 128.094963014;   g_res_mass['E'] = 129.04259308797; g_res_mass['M'] = 131.04048491299;
    g_res_mass['H'] = 137.05891185845; g_res_mass['F'] = 147.06841391299; g_res_mass['R'] = 156.1011110236;
    g_res_mass['Y'] = 163.06332853255; g_res_mass['W'] = 186.07931294986;
}
struct MassInit { MassInit() { init_masses(); } };
static MassInit g_mass_init;

// ---------------------------------------------------------------------------------------------------
// JDK / commons-math pieces the path depends on
// ---------------------------------------------------------------------------------------------------

// java.util.Random as specified by the JDK (48-bit LCG).  Used at pg/PeptideInput.java:432-433,499-500.
struct JavaRandom {
    uint64_t seed;
    explicit JavaRandom(int64_t s) { setSeed(s); }
    void setSeed(int64_t s) { seed = ((uint64_t)s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1); }
    int32_t next(int bits) {
        seed = (seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
        return (int32_t)(int64_t)(seed >> (48 - bits));
    }
    int32_t nextInt(int32_t bound) {
        int32_t r = next(31);
        int32_t m = bound - 1;
        if ((bound & m) == 0) {
            r = (int32_t)(((int64_t)bound * (int64_t)r) >> 31);
        } else {
            for (int32_t u = r;; u = next(31)) {
                r = u % bound;
                int64_t t = (int64_t)u - (int64_t)r + (int64_t)m;   // Java: int overflow -> negative
                if (t <= 0x7fffffffLL) break;
            }
        }
        return r;
    }
};

// StrictMath.log / log10 (fdlibm 5.3 e_log.c / e_log10.c) — the JDK's specified reference implementation of
// Math.log / Math.log10 (HyperscoreMatch.java:162-170,215; commons-math 2 MathUtils.factorialLog).
static inline uint32_t hi_word(double x) { uint64_t u; memcpy(&u, &x, 8); return (uint32_t)(u >> 32); }
static inline uint32_t lo_word(double x) { uint64_t u; memcpy(&u, &x, 8); return (uint32_t)u; }
static inline double with_hi(double x, uint32_t hi) { uint64_t u; memcpy(&u, &x, 8); u = (u & 0xffffffffULL) | ((uint64_t)hi << 32); double r; memcpy(&r, &u, 8); return r; }

#pragma GCC push_options
#pragma GCC optimize("fp-contract=off")
static double fd_log(double x) {
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        two54 = 1.80143985094819840000e+16,
                        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                        Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                        Lg7 = 1.479819860511658591e-01;
    double hfsq, f, s, z, R, w, t1, t2, dk;
    int32_t k, hx, i, j;
    uint32_t lx;
    hx = (int32_t)hi_word(x);
    lx = lo_word(x);
    k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -HUGE_VAL;
        if (hx < 0) return NAN;
        k -= 54; x *= two54;
        hx = (int32_t)hi_word(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    x = with_hi(x, (uint32_t)(hx | (i ^ 0x3ff00000)));
    k += (i >> 20);
    f = x - 1.0;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) { if (k == 0) return 0.0; dk = (double)k; return dk * ln2_hi + dk * ln2_lo; }
        R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        dk = (double)k; return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    s = f / (2.0 + f);
    dk = (double)k;
    z = s * s;
    i = hx - 0x6147a;
    w = z * z;
    j = 0x6b851 - hx;
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
        hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    } else {
        if (k == 0) return f - s * (f - R);
        return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
    }
}
static double fd_log10(double x) {
    static const double two54 = 1.80143985094819840000e+16, ivln10 = 4.34294481903251816668e-01,
                        log10_2hi = 3.01029995663611771306e-01, log10_2lo = 3.69423907715893078616e-13;
    double y, z;
    int32_t i, k, hx;
    uint32_t lx;
    hx = (int32_t)hi_word(x);
    lx = lo_word(x);
    k = 0;
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | lx) == 0) return -HUGE_VAL;
        if (hx < 0) return NAN;
        k -= 54; x *= two54;
        hx = (int32_t)hi_word(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
    hx = (hx & 0x000fffff) | ((0x3ff - i) << 20);
    y = (double)(k + i);
    x = with_hi(x, (uint32_t)hx);
    z = y * log10_2lo + ivln10 * fd_log(x);
    return z + y * log10_2hi;
}
#pragma GCC pop_options

// commons-math 2.x MathUtils (A12).  factorialLog: log(n!) from the exact long for n<21, else sum ln i.
static double factorial_log(int n) {
    static std::vector<double> cache;
    static std::atomic<bool> ready{false};
    static const int N = 20001;
    if (!ready.load()) {
        static std::atomic_flag lock = ATOMIC_FLAG_INIT;
        while (lock.test_and_set()) {}
        if (!ready.load()) {
            cache.assign(N, 0.0);
            uint64_t f = 1;
            for (int i = 0; i < 21; ++i) { if (i > 0) f *= (uint64_t)i; cache[i] = fd_log((double)f); }
            // n >= 21: for (i = 2..n) logSum += log(i)  — a fresh running sum per n in the reference, identical
            // to a prefix because the additions happen in the same order.
            double s = 0.0;
            for (int i = 2; i < N; ++i) { s += fd_log((double)i); if (i >= 21) cache[i] = s; }
            ready.store(true);
        }
        lock.clear();
    }
    if (n < 0 || n >= N) return NAN;
    return cache[n];
}
// factorialDouble(n): exact for n<21 else floor(exp(factorialLog(n)) + 0.5) (A12).  Only n<=60 is cached by
// HyperscoreMatch.generateFactorialValues(60) (HyperscoreMatch.java:213-218, CParameter.java:542).
static double log10_factorial(int n) {
    static double tab[65];
    static bool ready = false;
    if (!ready) {
        tab[0] = 0.0;
        uint64_t f = 1;
        for (int i = 1; i <= 64; ++i) {
            double fd;
            if (i < 21) { f *= (uint64_t)i; fd = (double)f; }
            else fd = std::floor(std::exp(factorial_log(i)) + 0.5);
            tab[i] = fd_log10(1.0 * fd);
        }
        ready = true;
    }
    return tab[n];
}

static inline long java_round(double x) { return (long)std::floor(x + 0.5); }   // Math.round, half up

// ---------------------------------------------------------------------------------------------------
// Parameters, modifications, peptides, spectra
// ---------------------------------------------------------------------------------------------------
struct Mod {
    int position = 0;
    std::string residues;
    double delta = 0.0;
    double loss = 0.0;
    int id = 0;
};
struct ModMatch { int mod; int site; };    // mod: index into Params::mods
struct Peptide {
    std::string seq;
    std::vector<ModMatch> mods;            // compomics "variable modifications" in list order
};
struct Spectrum {
    double prec_mz = 0.0;
    int charge = 0;                        // charge in use (a spectrum without CHARGE gets 2, then 3: FindCandidate...java:96-228)
    int given_charge = 0;                  // as loaded; 0 = the MGF block had no CHARGE line
    std::vector<double> mz, inten;         // file order
};
// "Hoisted" cost model (bench.py cpu_baseline): everything scorePeptide2Spectrum derives from the spectrum alone, computed
// once per spectrum instead of once per pair.  Same functions, same results; the as-written model is the default.
struct Prepared {
    double limit = 0.0; std::vector<int> order; std::vector<double> smz;      // annotator: matchable peaks, m/z ascending
    std::unordered_map<double, double> peakSet;                               // hyperscore survivors: m/z -> normalised intensity
    bool mvh_ok = false; std::unordered_map<double, int> peakCls; int classCounts[4] = {0, 0, 0, 0}; int npk = 0;   // MVH
};
struct Params {
    pq_params p;
    std::vector<Mod> mods;                 // var mods first (index = var idx), then fixed
    int n_var = 0, n_fixed = 0;
    std::vector<int> var_idx, fixed_idx;
};

static Mod to_mod(const pq_mod& m) {
    Mod r; r.position = m.position; r.residues = m.residues; r.delta = m.delta; r.loss = m.neutral_loss; r.id = m.id;
    return r;
}
static Params make_params(const pq_params& p) {
    Params P; P.p = p;
    P.n_var = p.n_var; P.n_fixed = p.n_fixed;
    for (int i = 0; i < p.n_var; ++i) { P.var_idx.push_back((int)P.mods.size()); P.mods.push_back(to_mod(p.var[i])); }
    for (int i = 0; i < p.n_fixed; ++i) { P.fixed_idx.push_back((int)P.mods.size()); P.mods.push_back(to_mod(p.fixed[i])); }
    return P;
}

// ModificationUtils.getPossibleModificationSites (A9): ascending sites.
static std::vector<int> possible_sites(const std::string& seq, const Mod& m) {
    std::vector<int> s;
    int L = (int)seq.size();
    switch (m.position) {
        case PQ_MOD_ANYWHERE:
            for (int i = 0; i < L; ++i) if (m.residues.find(seq[i]) != std::string::npos) s.push_back(i + 1);
            break;
        case PQ_MOD_PEP_NTERM: case PQ_MOD_PROT_NTERM: s.push_back(0); break;
        case PQ_MOD_PEP_CTERM: case PQ_MOD_PROT_CTERM: s.push_back(L + 1); break;
        case PQ_MOD_PEP_NTERM_AA: case PQ_MOD_PROT_NTERM_AA:
            if (L > 0 && m.residues.find(seq[0]) != std::string::npos) s.push_back(1);
            break;
        case PQ_MOD_PEP_CTERM_AA: case PQ_MOD_PROT_CTERM_AA:
            if (L > 0 && m.residues.find(seq[L - 1]) != std::string::npos) s.push_back(L);
            break;
    }
    return s;
}

// Peptide.getMass as used by JPeptide.getMass (PSMMatch/JPeptide.java:73-78): H2O + residues + mods in list order.
static double peptide_mass(const Params& P, const Peptide& pep) {
    double m = H2O;
    for (char c : pep.seq) m += g_res_mass[(int)c];
    for (const ModMatch& mm : pep.mods) m += P.mods[mm.mod].delta;
    return m;
}

// PeptideInput.addFixedModification(Peptide) (pg/PeptideInput.java:213-237)
static void add_fixed_modification(const Params& P, Peptide& pep) {
    if (P.n_fixed == 0) return;
    std::set<int> varSites;
    for (const ModMatch& mm : pep.mods) varSites.insert(mm.site);
    for (int fi : P.fixed_idx) {
        for (int k : possible_sites(pep.seq, P.mods[fi]))
            if (!varSites.count(k)) pep.mods.push_back({fi, k});
    }
}

// combinatoricslib3 Generator.combination(list).simple(k): lexicographic index order.
template <class F> static void for_each_combination(int n, int k, F f) {
    std::vector<int> idx(k);
    for (int i = 0; i < k; ++i) idx[i] = i;
    if (k > n) return;
    while (true) {
        f(idx);
        int i = k - 1;
        while (i >= 0 && idx[i] == n - k + i) --i;
        if (i < 0) break;
        ++idx[i];
        for (int j = i + 1; j < k; ++j) idx[j] = idx[j - 1] + 1;
    }
}

// PeptideInput.calcPeptideIsoforms (pg/PeptideInput.java:114-181)
static std::vector<Peptide> calc_peptide_isoforms(const Params& P, const std::string& seq) {
    std::vector<Peptide> out;
    if (P.n_var >= 1) {
        std::vector<ModMatch> all_sites;   // JPTM(pos, ptm) in var-mod order, sites ascending
        for (int vi : P.var_idx)
            for (int s : possible_sites(seq, P.mods[vi])) all_sites.push_back({vi, s});
        if (!all_sites.empty()) {
            int maxNum = std::min((int)all_sites.size(), (int)P.p.max_var_mods);
            for (int k = 1; k <= maxNum; ++k) {
                for_each_combination((int)all_sites.size(), k, [&](const std::vector<int>& idx) {
                    Peptide mp; mp.seq = seq;
                    std::map<int, int> pos2mods; int maxPerAA = 0;
                    for (int i : idx) {
                        mp.mods.push_back(all_sites[i]);
                        int c = ++pos2mods[all_sites[i].site];
                        if (c > maxPerAA) maxPerAA = c;
                    }
                    if (maxPerAA <= P.p.max_mods_per_aa) { add_fixed_modification(P, mp); out.push_back(mp); }
                });
            }
        }
    }
    Peptide plain; plain.seq = seq;
    add_fixed_modification(P, plain);
    out.push_back(plain);
    return out;
}

// ---------------------------------------------------------------------------------------------------
// Fragment ions and the annotator (compomics; SURVEY.md Appendix A.2, assumptions A1-A7)
// ---------------------------------------------------------------------------------------------------
struct Ion { int type; /*0=b,1=y*/ int number; double mass; bool has_loss; };
struct IonMatch { int type; int number; int charge; double peakMz; double peakIntensity; };

// Forward / rewind accumulation in the order IonFactory.getFragmentIons walks the sequence:
// residue mass first, then each modification on that site in list order; N-term (site 0) mods open the
// forward sum, C-term (site L+1) mods open the rewind sum.  y mass = rewind + H2O.
static void fragment_masses(const Params& P, const Peptide& pep, std::vector<double>& b, std::vector<double>& y) {
    int L = (int)pep.seq.size();
    b.assign(L, 0.0); y.assign(L, 0.0);    // index = ion number (1..L-1)
    double fwd = 0.0, rew = 0.0;
    for (const ModMatch& mm : pep.mods) if (mm.site == 0) fwd += P.mods[mm.mod].delta;
    for (const ModMatch& mm : pep.mods) if (mm.site == L + 1) rew += P.mods[mm.mod].delta;
    for (int k = 1; k <= L - 1; ++k) {
        int fi = k;            // 1-based site of the residue joining the forward ion
        fwd += g_res_mass[(int)pep.seq[fi - 1]];
        for (const ModMatch& mm : pep.mods) if (mm.site == fi) fwd += P.mods[mm.mod].delta;
        int ri = L - k + 1;    // 1-based site of the residue joining the rewind ion
        rew += g_res_mass[(int)pep.seq[ri - 1]];
        for (const ModMatch& mm : pep.mods) if (mm.site == ri) rew += P.mods[mm.mod].delta;
        b[k] = fwd;
        y[k] = rew + H2O;
    }
}
static inline double theoretic_mz(double mass, int charge) { return (mass + charge * g_a.proton) / charge; }

// Spectrum.getIntensityLimit(percentile, 0.05) -> BasicMathFunctions.percentile (A2)
static double intensity_limit(const Spectrum& sp, double frac) {
    std::vector<double> v(sp.inten);
    std::sort(v.begin(), v.end());
    size_t n = v.size();
    if (n == 0) return 0.0;
    if (n == 1) return v[0];
    double indexDouble = frac * (double)(n - 1);
    int index = (int)indexDouble;
    double valueAtIndex = v[index];
    double rest = indexDouble - index;
    if (index == (int)n - 1 || rest == 0) return valueAtIndex;
    return valueAtIndex + rest * (v[index + 1] - valueAtIndex);
}

// JMatch.getSpectrumAnnotation (PSMMatch/JMatch.java:388-466) -> PeptideSpectrumAnnotator on the RAW spectrum.
static void annotator_index(const Spectrum& sp, double& limit, std::vector<int>& order, std::vector<double>& smz) {
    limit = intensity_limit(sp, 0.05);             // JMatch.java:427,430 (intensityLimit 0.05, percentile)
    // SpectrumIndex: built per annotator call in the reference (new PeptideSpectrumAnnotator per pair); here the
    // matchable peaks (intensity >= limit, A2) in ascending m/z (stable), searched by bisection.
    order.clear();
    for (size_t i = 0; i < sp.mz.size(); ++i) if (sp.inten[i] >= limit) order.push_back((int)i);
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return sp.mz[a] < sp.mz[c]; });
    smz.resize(order.size());
    for (size_t i = 0; i < order.size(); ++i) smz[i] = sp.mz[order[i]];
}
static std::vector<IonMatch> annotate(const Params& P, const Peptide& pep, const Spectrum& sp, const Prepared* H = nullptr) {
    std::vector<IonMatch> out;
    int z = sp.charge;
    std::vector<int> charges;                      // JMatch.java:398-407
    if (z == 1) charges.push_back(1); else for (int c = 1; c < z; ++c) charges.push_back(c);
    std::vector<double> b, y;
    fragment_masses(P, pep, b, y);
    int L = (int)pep.seq.size();
    double itol = P.p.itol;
    double limit_l; std::vector<int> order_l; std::vector<double> smz_l;
    if (!H) annotator_index(sp, limit_l, order_l, smz_l);
    const std::vector<int>& order = H ? H->order : order_l;
    const std::vector<double>& smz = H ? H->smz : smz_l;
    for (int type = 0; type < 2; ++type) {         // A5: b then y
        const std::vector<double>& m = type == 0 ? b : y;
        for (int num = 1; num <= L - 1; ++num) {
            for (int c : charges) {
                if (!(c == 1 || (c <= num && c < z))) continue;      // A6 chargeValidated
                double theo = theoretic_mz(m[num], c);
                int best = -1; double bestI = 0.0, bestErr = 0.0;
                // start a little below the window; the exact membership test is the fabs() below
                size_t k = std::lower_bound(smz.begin(), smz.end(), theo - itol - 1e-6) - smz.begin();
                for (; k < smz.size() && smz[k] <= theo + itol + 1e-6; ++k) {
                    int pi = order[k];
                    double err = sp.mz[pi] - theo;
                    bool in = g_a.a3_closed_interval ? (std::fabs(err) <= itol) : (std::fabs(err) < itol);
                    if (!in) continue;
                    if (best < 0 || sp.inten[pi] > bestI ||
                        (g_a.a4_tie_smaller_error && sp.inten[pi] == bestI && std::fabs(err) < std::fabs(bestErr))) {
                        best = pi; bestI = sp.inten[pi]; bestErr = err;
                    }
                }
                if (best >= 0) out.push_back({type, num, c, sp.mz[best], sp.inten[best]});
            }
        }
    }
    return out;
}

// ---------------------------------------------------------------------------------------------------
// JSpectrum / JMatch filters (PSMMatch/JSpectrum.java, PSMMatch/JMatch.java)
// ---------------------------------------------------------------------------------------------------
struct JPeak { double mz, inten; int cls; };
struct JSpectrum {
    std::vector<JPeak> peaks, raw;
    void resetPeaks() { peaks = raw; }                                  // JSpectrum.java:88-95
    std::vector<JPeak>& getPeaks() { if (peaks.empty()) resetPeaks(); return peaks; }   // :71-76 (reload quirk)
    void sortByMZ() { std::stable_sort(peaks.begin(), peaks.end(), [](const JPeak& a, const JPeak& b) { return a.mz < b.mz; }); }
    void sortByIntensity() { std::stable_sort(peaks.begin(), peaks.end(), [](const JPeak& a, const JPeak& b) { return a.inten < b.inten; }); }
    double maxIntensity() const { double m = 0; for (auto& p : peaks) if (m < p.inten) m = p.inten; return m; }   // :108-116 (no reload)
    double tic() const { double t = 0.0; for (auto& p : peaks) t += p.inten; return t; }
};
struct JMatch {
    const Params& P; const Spectrum& sp; JSpectrum js; int psm_charge; double exp_mz;
    JMatch(const Params& P_, const Spectrum& s) : P(P_), sp(s) {
        // PSMT:1145-1153: raw peaks in file order, resetPeaks, sort by m/z
        for (size_t i = 0; i < s.mz.size(); ++i) js.raw.push_back({s.mz[i], s.inten[i], 0});
        js.resetPeaks(); js.sortByMZ();
        psm_charge = s.charge; exp_mz = s.prec_mz;
    }
    void removeParentPeak() {                                           // JMatch.java:101-127
        int pc = psm_charge; if (pc == 0) pc = 1;
        double tol_mz = 1.0 * P.p.itol / pc;
        auto& v = js.getPeaks();
        v.erase(std::remove_if(v.begin(), v.end(), [&](const JPeak& p) { return std::fabs(p.mz - exp_mz) <= tol_mz; }), v.end());
    }
    void removeLowMasses() {                                            // :130-142
        auto& v = js.getPeaks();
        v.erase(std::remove_if(v.begin(), v.end(), [&](const JPeak& p) { return p.mz <= 150.0; }), v.end());
    }
    void filterByPeakCount(int maxPeaks) {                              // :148-174
        int n = (int)js.getPeaks().size();
        if (n > maxPeaks) { js.sortByIntensity(); js.getPeaks().erase(js.getPeaks().begin(), js.getPeaks().begin() + (n - maxPeaks)); }
    }
    void removeLowIntensityPeak() {                                     // :178-193
        double lim = 1.0 * 0.05 * js.maxIntensity();
        auto& v = js.getPeaks();
        v.erase(std::remove_if(v.begin(), v.end(), [&](const JPeak& p) { return p.inten < lim; }), v.end());
    }
    void isotopePass(double gap) {                                      // removeIsotopes :223-248 / cleanIsotopes :255-280
        if (js.getPeaks().size() > 2) {
            js.sortByMZ();
            auto& v = js.getPeaks();
            JPeak p1 = v[0];
            std::vector<JPeak> tmp;
            double mz1 = p1.mz;
            for (size_t i = 1; i < v.size(); ++i) {
                const JPeak& p2 = v[i];
                if (p2.mz - mz1 >= gap || p2.mz < 200) { tmp.push_back(p1); p1 = p2; mz1 = p1.mz; }
                else if (p2.inten > p1.inten) { p1 = p2; mz1 = p1.mz; }
            }
            tmp.push_back(p1);
            js.peaks = tmp;
        }
    }
    void normalizePeaks(double maxAfter) {                              // :286-292
        double mx = js.maxIntensity();
        auto& v = js.getPeaks();
        for (auto& p : v) p.inten = 1.0 * maxAfter * p.inten / mx;
    }
    void filterByTIC() {                                                // :341-364
        if (js.getPeaks().size() > 7) {
            double tic = js.tic();
            double rel = 0.0;
            js.sortByIntensity();
            std::vector<JPeak> tmp;
            auto& v = js.getPeaks();
            for (int i = (int)v.size() - 1; i >= 0; --i) {
                rel += v[i].inten / tic;
                if (rel <= 0.95) tmp.push_back(v[i]);
            }
            js.peaks = tmp;
        }
    }
    void filterPeakByRange(double lo, double hi) {                      // :372-383
        auto& v = js.getPeaks();
        v.erase(std::remove_if(v.begin(), v.end(), [&](const JPeak& p) { return p.mz < lo || p.mz > hi; }), v.end());
    }
};

struct Score { double score = 0.0; int a = 0; int b = 0; bool died = false; };

// HyperscoreMatch.calcHyperScore (PSMMatch/HyperscoreMatch.java:61-187)
static void hyperscore_peak_set(const Params& P, const Spectrum& sp, std::unordered_map<double, double>& peakSet) {
    JMatch m(P, sp);
    m.js.resetPeaks();
    // SpectrumPreProcessing :47-59
    m.isotopePass(0.95); m.removeParentPeak(); m.removeLowMasses(); m.removeLowIntensityPeak();
    m.isotopePass(1.5); m.filterByPeakCount(50); m.normalizePeaks(100.0);
    peakSet.clear();                                                    // :71-78
    m.js.getPeaks(); m.js.sortByMZ();
    for (auto& p : m.js.getPeaks()) peakSet[p.mz] = p.inten;
}
static Score calc_hyperscore(const Params& P, const Peptide& pep, const Spectrum& sp, const Prepared* H = nullptr) {
    std::vector<IonMatch> matches = annotate(P, pep, sp, H);            // initialize(false,2) -> getSpectrumAnnotation
    std::unordered_map<double, double> peakSet_l;
    if (!H) hyperscore_peak_set(P, sp, peakSet_l);
    const std::unordered_map<double, double>& peakSet = H ? H->peakSet : peakSet_l;
    std::unordered_set<double> used;
    int cb = 0, cy = 0; double dScore = 0.0;
    for (const IonMatch& im : matches) {                                // :89-147
        auto it = peakSet.find(im.peakMz);
        if (it == peakSet.end()) continue;
        if (used.count(im.peakMz)) continue;
        used.insert(im.peakMz);
        if (im.type == 0) cb++; else cy++;
        dScore += it->second;
    }
    if (cb > 64) cb = 64;
    if (cy > 64) cy = 64;
    Score r; r.a = cb; r.b = cy;
    double hs;
    if (cb > 0 && cy > 0) hs = fd_log10(dScore) + log10_factorial(cb) + log10_factorial(cy);
    else if (cb == 0 && cy > 0) hs = fd_log10(dScore) + log10_factorial(cy);
    else if (cy == 0 && cb > 0) hs = fd_log10(dScore) + log10_factorial(cb);
    else { r.score = 0.0; return r; }
    r.score = 4.0 * hs;
    return r;
}

// MVHMatch.calcMVH (PSMMatch/MVHMatch.java:121-302)
static double ln_combin(int n, int k, bool& died) {                     // :91-101
    if (n < 0 || k < 0) return 0;
    if (n < k) { died = true; return 0; }    // commons-math throws -> the pool task dies; treated as "not scored"
    return factorial_log(n) - factorial_log(n - k) - factorial_log(k);
}
// the spectrum-only part of calcMVH: false = fewer than 7 peaks survive (mvh stays 0.0)
static bool mvh_peak_classes(const Params& P, const Spectrum& sp, std::unordered_map<double, int>& peakCls, int classCounts[4], int& npk) {
    JMatch m(P, sp);
    m.js.resetPeaks();
    // SpectrumPreProcessing :76-89
    m.filterPeakByRange(200.0, 2000.0); m.removeParentPeak(); m.filterByTIC(); m.isotopePass(0.95); m.filterByPeakCount(300);
    {   // classifyPeakIntensities :46-73
        int total = (int)m.js.getPeaks().size();
        if (total < 7) return false;
        int npeak = (int)std::floor(1.0 * total / 7.0);
        m.js.sortByIntensity();
        auto& v = m.js.getPeaks();
        int n1 = npeak, n2 = 2 * npeak;
        for (int i = total - 1; i >= total - n1; --i) v[i].cls = 0;
        for (int i = total - 1 - n1; i >= total - n1 - n2; --i) v[i].cls = 1;
        for (int i = total - 1 - n1 - n2; i >= 0; --i) v[i].cls = 2;
    }
    peakCls.clear();                                                    // :234-241
    for (int i = 0; i < 4; ++i) classCounts[i] = 0;
    for (auto& p : m.js.getPeaks()) { peakCls[p.mz] = p.cls; classCounts[p.cls]++; }
    npk = (int)m.js.getPeaks().size();
    classCounts[3] = (int)(java_round((int)(2000.0 - 200.0) / (P.p.itol * 2.0)) - npk);   // :242-244
    return true;
}
static Score calc_mvh(const Params& P, const Peptide& pep, const Spectrum& sp, const Prepared* H = nullptr) {
    std::vector<IonMatch> matches = annotate(P, pep, sp, H);            // initialize(false,3)
    Score r;
    std::unordered_map<double, int> peakCls_l; int classCounts_l[4]; int npk_l = 0;
    if (H ? !H->mvh_ok : !mvh_peak_classes(P, sp, peakCls_l, classCounts_l, npk_l)) return r;      // mvh stays 0.0
    const std::unordered_map<double, int>& peakCls = H ? H->peakCls : peakCls_l;
    const int* classCounts = H ? H->classCounts : classCounts_l;
    const int npk = H ? H->npk : npk_l;
    // predicted peaks :150-227
    std::vector<double> b, y;
    fragment_masses(P, pep, b, y);
    int L = (int)pep.seq.size();
    std::vector<double> losses;          // A8: one single-loss variant per distinct mod-defined loss on the peptide
    if (g_a.a8_mod_losses_in_prediction) {
        for (const ModMatch& mm : pep.mods) {
            double l = P.mods[mm.mod].loss;
            if (l != 0.0 && std::find(losses.begin(), losses.end(), l) == losses.end()) losses.push_back(l);
        }
    }
    int maxc = (sp.charge == 2) ? 1 : 2;                                // :165-172
    std::set<double> predicted;
    for (int type = 0; type < 2; ++type) {
        const std::vector<double>& mm = type == 0 ? b : y;
        for (int num = 1; num <= L - 1; ++num) {
            for (int li = -1; li < (int)losses.size(); ++li) {
                double mass = li < 0 ? mm[num] : mm[num] - losses[li];
                for (int c = 1; c <= maxc; ++c) {
                    double mz = theoretic_mz(mass, c);
                    if (mz >= 200.0 && mz <= 2000.0) predicted.insert(mz);
                }
            }
        }
    }
    int totalPredicted = 0;
    if (!predicted.empty()) {                                           // removeFuzzyValues :104-119
        std::vector<double> val(predicted.begin(), predicted.end());
        double a1 = val[0];
        for (size_t i = 1; i < val.size(); ++i) { double a2 = val[i]; if (a2 - a1 >= P.p.itol) totalPredicted++; a1 = a2; }
        totalPredicted++;
    } else {
        r.died = true;   // val[0] on an empty array throws in the reference
        return r;
    }
    int key[4] = {0, 0, 0, 0};
    std::unordered_set<double> used;
    for (const IonMatch& im : matches) {                                // :246-279
        if (im.peakMz < 200.0 || im.peakMz > 2000.0) continue;
        auto it = peakCls.find(im.peakMz);
        if (it == peakCls.end()) continue;
        if (used.count(im.peakMz)) continue;
        used.insert(im.peakMz);
        key[it->second]++;
    }
    int matched = (int)used.size();
    key[3] = totalPredicted - matched;
    int totalBins = classCounts[3] + npk;
    double mvh = 0.0; bool died = false;
    for (int i = 0; i <= 3; ++i) mvh += ln_combin(classCounts[i], key[i], died);
    mvh -= ln_combin(totalBins, totalPredicted, died);
    mvh = 0.0 - mvh;
    r.score = died ? 0.0 : mvh; r.died = died; r.a = matched; r.b = totalPredicted;
    return r;
}

// PeptideSearchMT.scorePeptide2Spectrum (pg/PeptideSearchMT.java:1144-1251)
static Score score_peptide_to_spectrum(const Params& P, const Peptide& pep, const Spectrum& sp, const Prepared* H = nullptr) {
    if (P.p.score_method == 2) return calc_mvh(P, pep, sp, H);
    return calc_hyperscore(P, pep, sp, H);
}
static void prepare_spectrum(const Params& P, const Spectrum& sp, Prepared& H) {
    annotator_index(sp, H.limit, H.order, H.smz);
    if (P.p.score_method == 2) H.mvh_ok = mvh_peak_classes(P, sp, H.peakCls, H.classCounts, H.npk);
    else hyperscore_peak_set(P, sp, H.peakSet);
}

// ---------------------------------------------------------------------------------------------------
// Shuffles (pg/PeptideInput.java:376-557, pg/HashMapValueCompare.java)
// ---------------------------------------------------------------------------------------------------
static std::vector<std::pair<char, int>> hashmap_order_by_count(const std::string& s) {
    // java.util.HashMap<String,Integer> iteration order for single-char keys: bucket = code & (cap-1); cap 16,
    // doubled when size exceeds 0.75*cap; insertion order inside a bucket.
    std::vector<char> keys;          // insertion order of distinct keys
    std::map<char, int> count;
    for (char c : s) { if (!count.count(c)) keys.push_back(c); count[c]++; }
    int cap = 16;
    while ((int)keys.size() > (cap * 3) / 4) cap *= 2;
    std::vector<std::pair<char, int>> entries;
    for (int bkt = 0; bkt < cap; ++bkt)
        for (char c : keys) if (((int)c & (cap - 1)) == bkt) entries.push_back({c, count[c]});
    std::stable_sort(entries.begin(), entries.end(), [](const std::pair<char, int>& a, const std::pair<char, int>& b) { return a.second < b.second; });
    return entries;
}
static int64_t binomial(int n, int k) {
    if (k < 0 || k > n) return 0;
    if (k > n - k) k = n - k;
    unsigned __int128 r = 1;
    for (int i = 1; i <= k; ++i) { r = r * (unsigned)(n - k + i) / (unsigned)i; }
    return (int64_t)r;
}
static std::vector<std::string> generate_random_peptides_fast(const std::string& pep_seq, int num, int64_t seed) {
    std::vector<std::string> out;
    int len = (int)pep_seq.size() - 1;                                  // fixCtermAA = true
    std::string s = pep_seq.substr(0, len);
    auto aaList = hashmap_order_by_count(s);
    int64_t nComb = 1; int rn = len;
    for (auto& e : aaList) {                                            // :402-418
        int64_t c = binomial(rn, e.second);
        if (nComb >= 10000000) nComb = 10000000; else nComb = (int64_t)((uint64_t)nComb * (uint64_t)c);
        rn -= e.second;
    }
    std::unordered_set<std::string> pSet; pSet.insert(pep_seq);
    num = (int64_t)num > nComb ? (int)nComb : num;
    JavaRandom rnd(seed);
    for (int i = 1; i <= num; ++i) {
        std::vector<int> pos;                                           // HashSet<Integer> of 1..len iterates ascending
        for (int j = 1; j <= len; ++j) pos.push_back(j);
        std::string pep(pep_seq.size(), '?');
        for (auto& e : aaList) {
            for (int r_ = 0; r_ < e.second; ++r_) {
                int r = rnd.nextInt((int)pos.size());
                int pp = pos[r];
                pep[pp] = e.first;
                pos.erase(pos.begin() + r);
            }
        }
        std::string rpep = pep.substr(1) + pep_seq[len];                // pep[0] is null -> "" in StringUtils.join
        if (pSet.count(rpep)) continue;
        pSet.insert(rpep); out.push_back(rpep);
    }
    return out;
}
// PeptideInput.getModificationPeptide (pg/PeptideInput.java:485-557)
static Peptide get_modification_peptide(const Params& P, const Peptide& target, const std::string& seq) {
    Peptide pep; pep.seq = seq;
    std::set<int> used;
    JavaRandom rnd(P.p.mod_seed);
    for (const ModMatch& mm : target.mods) {
        std::vector<int> sites = possible_sites(seq, P.mods[mm.mod]);
        if (sites.empty()) continue;
        std::vector<int> free_;
        for (int s : sites) if (!used.count(s)) free_.push_back(s);
        if (!free_.empty()) {
            int idx = rnd.nextInt((int)free_.size());
            pep.mods.push_back({mm.mod, free_[idx]});
            used.insert(free_[idx]);
        }
    }
    add_fixed_modification(P, pep);
    return pep;
}

// ---------------------------------------------------------------------------------------------------
// Digest + reference index (pg/DatabaseInput.java:402-556, DigestProteinWorker.java:64-137)
// ---------------------------------------------------------------------------------------------------
static void digest_protein(const Params& P, std::string prot, std::set<std::string>& out) {
    for (auto& c : prot) c = (char)toupper(c);
    prot.erase(std::remove(prot.begin(), prot.end(), '*'), prot.end());
    for (auto& c : prot) if (c == 'I') c = 'L';
    int n = (int)prot.size();
    // Enzyme table of DatabaseInput.getEnzymeByIndex (pg/DatabaseInput.java:145-267): residues cleaved after ("before" the
    // cut), residues cleaved before ("after" the cut), and the residue that blocks a cut when it follows (restrictionAfter).
    // compomics Enzyme.isCleavageSite(x, y): (x in before and y not in restrictionAfter) or (y in after) (A10).
    static const struct { const char* before; const char* after; const char* restrict_after; } kEnzymes[18] = {
        {"", "", ""},              // 0 NoEnzyme: unspecific, not supported
        {"RK", "", "P"},           // 1 Trypsin
        {"RK", "", ""},            // 2 Trypsin (no P rule)
        {"R", "", "P"},            // 3 Arg-C
        {"R", "", ""},             // 4 Arg-C (no P rule)
        {"", "R", ""},             // 5 Arg-N
        {"E", "", ""},             // 6 Glu-C
        {"K", "", "P"},            // 7 Lys-C
        {"K", "", ""},             // 8 Lys-C (no P rule)
        {"", "K", ""},             // 9 Lys-N
        {"", "D", ""},             // 10 Asp-N
        {"", "DE", ""},            // 11 Asp-N (ambic)
        {"FYWL", "", "P"},         // 12 Chymotrypsin
        {"FYWL", "", ""},          // 13 Chymotrypsin (no P rule)
        {"FL", "", ""},            // 14 Pepsin A
        {"M", "", ""},             // 15 CNBr
        {"", "AFILMV", ""},        // 16 Thermolysin
        {"", "RK", ""},            // 17 LysargiNase
    };
    const auto& E = kEnzymes[(P.p.enzyme >= 1 && P.p.enzyme <= 17) ? P.p.enzyme : 1];
    auto in = [](const char* set, char c) { return strchr(set, c) != nullptr && c != 0; };
    // enzyme 0 = NoEnzyme (pg/DatabaseInput.java:117-121,149-156): CleavageParameter.unSpecific — every stretch of the protein whose mass
    // lies in the digest window is a peptide (A14: compomics' unspecific iterator; missed cleavages play no role); the length filter
    // of DigestProteinWorker.java:96 then keeps min_len .. max_len residues
    const bool unspecific = P.p.enzyme == 0;
    std::vector<int> starts; starts.push_back(0);
    for (int i = 0; i < n - 1; ++i)
        if (unspecific || (in(E.before, prot[i]) && !in(E.restrict_after, prot[i + 1])) || in(E.after, prot[i + 1])) starts.push_back(i + 1);
    starts.push_back(n);
    int ns = (int)starts.size() - 1;
    const int max_joined = unspecific ? P.p.max_len - 1 : P.p.max_missed;
    for (int i = 0; i < ns; ++i) {
        for (int mc = 0; mc <= max_joined && i + mc < ns; ++mc) {
            int a = starts[i], e = starts[i + mc + 1];
            int len = e - a;
            if (len <= 0) continue;
            std::string pep = prot.substr(a, len);
            double mass = H2O; bool bad = false;
            for (char c : pep) { if (g_res_mass[(int)c] == 0.0) bad = true; mass += g_res_mass[(int)c]; }
            if (bad) continue;                                          // B/J/O/U/X/Z: not handled (avoided in synthetic data)
            if (mass < P.p.min_mass || mass > P.p.max_mass) continue;
            if (len < P.p.min_len || len > P.p.max_len) continue;
            if (pep.find('X') != std::string::npos) continue;
            out.insert(pep);
        }
    }
}

struct RefForm { Peptide pep; double mass; int seq_index; };

// ---------------------------------------------------------------------------------------------------
// The search context (Steps 2-5)
// ---------------------------------------------------------------------------------------------------
struct Psm {
    int spectrum, form; Score sc; double exp_mass, pep_mass; int iso; int charge = 0;
    int n_db = 0, total_db = 0, n_random = -1, total_random = -1; bool valid = true; double p_value = 100.0;
    int rank = 0, n_ptm = -1, confident = 0;
};
struct Ctx {
    Params P;
    std::vector<Spectrum> spectra;
    std::vector<std::string> target_seqs;
    std::vector<Peptide> target_forms; std::vector<int> target_form_seq; std::vector<double> target_form_mass;
    std::vector<std::string> ref_seqs;
    std::vector<RefForm> ref_forms;                       // mass-sorted (mass, sequence, form order)
    std::map<long, std::vector<int>> ref_index;           // round(10*mass) -> ref form ids, mass ascending
    std::vector<Psm> psms;
    std::vector<pq_competitor> comp3, comp5;
    std::map<int, std::vector<std::string>> shuffles;     // target seq -> shuffles
    uint64_t pairs[4] = {0, 0, 0, 0};
    int threads = 1;
    std::string err;
    // cost model of the CPU baseline: false = as written (pre-processing per pair), true = hoisted (once per spectrum)
    bool hoisted = false;
    std::vector<Prepared> prep;                           // hoisted: per spectrum, filled by Step 2 for the spectra it scores
    // GeneratePeptideformWorker mass range forced by the caller (sampled-parity tests: the window of a larger run)
    bool forced_window = false; double forced_lo = 0.0, forced_hi = 0.0;
    // SpectraInput.pep2targeted_spectra (PSMT:264-309): target sequence -> spectra it may be scored against, with the PSM status
    std::vector<std::pair<int, int>> targeted;            // (target sequence index, spectrum index) in the caller's order
    std::map<std::string, std::map<int, int>> pep2targeted;   // sequence -> spectrum -> PsmType (0 no_spectrum_match, 1 spectrum_matched)
    const Prepared* H(int si) const { return hoisted ? &prep[(size_t)si] : nullptr; }
};

static double precursor_mass(const Spectrum& s) { return s.prec_mz * s.charge - s.charge * g_a.proton; }   // Precursor.getMass
static std::vector<std::pair<double, int>> isotope_masses(const Params& P, double m) {                     // CParameter.java:423-448
    std::vector<std::pair<double, int>> r;
    if (P.p.n_isotope == 1 && P.p.isotope[0] == 0) { r.push_back({m, 0}); return r; }
    for (int i = 0; i < P.p.n_isotope; ++i) r.push_back({m - (double)P.p.isotope[i] * 1.00335, P.p.isotope[i]});
    return r;
}
static bool in_window(const Params& P, double pepMass, double mass) {   // DBSearchWorker.java:47-52
    double del = std::fabs(pepMass - mass);
    if (P.p.tol_ppm) del = (1.0e6) * 1.0 * del / pepMass;
    return del <= P.p.tol;
}

template <class F> static void parallel_for(int threads, int64_t n, F f) {
    if (threads <= 1 || n < 2) { for (int64_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<int64_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back([&]() { for (;;) { int64_t i = next.fetch_add(1); if (i >= n) break; f(i); } });
    for (auto& t : th) t.join();
}

// Step 2: FindCandidateSpectraAndScoreWorker.run (pg/FindCandidateSpectraAndScoreWorker.java:33-229)
static void step2(Ctx& C) {
    const Params& P = C.P;
    std::map<long, std::vector<int>> pIndex;                            // SpectraInput.java:96-109
    for (size_t f = 0; f < C.target_forms.size(); ++f) pIndex[java_round(C.target_form_mass[f] * 10.0)].push_back((int)f);
    std::vector<std::vector<Psm>> per(C.spectra.size());
    std::vector<uint64_t> cnt(C.spectra.size(), 0);
    if (C.hoisted) C.prep.assign(C.spectra.size(), Prepared());
    for (auto& kv : C.pep2targeted) for (auto& st : kv.second) st.second = 0;      // PsmType.no_spectrum_match
    for (auto& sp : C.spectra) sp.charge = sp.given_charge;
    std::vector<std::vector<std::pair<const std::string*, int>>> touched(C.spectra.size());     // targeted pairs seen in a window
    parallel_for(C.threads, (int64_t)C.spectra.size(), [&](int64_t si) {
        Spectrum& sp = C.spectra[si];
        if (P.p.library_mode ? (int)sp.mz.size() < P.p.min_peaks : (int)sp.mz.size() <= P.p.min_peaks) return;   // SpectraInput.java:183; library: MsLibrarySearchWorker.java:239
        if (P.p.library_mode && sp.given_charge < 1) return;            // MsLibrarySearchWorker.java:234-236
        // one pass of the window loop with the precursor charge set to z; returns whether a PSM was saved
        auto pass = [&](int z, bool update_status) -> bool {
            sp.charge = z;
            bool prepared = false, find = false;
            double exp_mass = precursor_mass(sp);
            for (auto& im : isotope_masses(P, exp_mass)) {
                double mass = im.first;
                long va = java_round(mass * 10.0);
                // a library search (msio/MsLibrarySearchWorker.java:70-120) knows no buckets: every form with |mass error| <= tol
                const double reach = P.p.tol_ppm ? P.p.tol * 1e-6 * (std::fabs(mass) + 1.0) * 1.01 + 1e-6 : P.p.tol * 1.01 + 1e-6;
                const long k_lo = P.p.library_mode ? java_round((mass - reach) * 10.0) - 1 : va - 1, k_hi = P.p.library_mode ? java_round((mass + reach) * 10.0) + 1 : va + 1;
                for (long i = k_lo; i <= k_hi; ++i) {
                    auto it = pIndex.find(i);
                    if (it == pIndex.end()) continue;
                    for (int f : it->second) {
                        if (!in_window(P, C.target_form_mass[f], mass)) continue;
                        if (!C.pep2targeted.empty()) {                  // :60-74
                            const std::string& seq = C.target_forms[f].seq;
                            auto pt = C.pep2targeted.find(seq);
                            if (pt != C.pep2targeted.end()) {
                                if (!pt->second.count((int)si)) continue;                       // not a spectrum named for this peptide
                                if (update_status) touched[si].push_back({&pt->first, (int)si});  // PsmType.spectrum_matched
                            }
                        }
                        if (C.hoisted && !prepared) { prepare_spectrum(P, sp, C.prep[(size_t)si]); prepared = true; }
                        Score sc = score_peptide_to_spectrum(P, C.target_forms[f], sp, C.H((int)si));
                        cnt[si]++;
                        if (!sc.died && sc.score > P.p.min_score) {
                            Psm p; p.spectrum = (int)si; p.form = f; p.sc = sc; p.exp_mass = mass; p.pep_mass = C.target_form_mass[f]; p.iso = im.second; p.charge = z;
                            per[si].push_back(p);
                            find = true;
                        }
                    }
                }
            }
            return find;
        };
        if (sp.given_charge >= 1) pass(sp.given_charge, true);
        else if (!pass(2, false)) pass(3, false);                       // no CHARGE: z = 2, and z = 3 only if z = 2 saved nothing (:96-228)
    });
    for (auto& v : touched) for (auto& t : v) C.pep2targeted[*t.first][t.second] = 1;
    C.psms.clear();
    for (size_t si = 0; si < per.size(); ++si) { C.pairs[0] += cnt[si]; for (auto& p : per[si]) C.psms.push_back(p); }
    std::stable_sort(C.psms.begin(), C.psms.end(), [](const Psm& a, const Psm& b) { return a.spectrum != b.spectrum ? a.spectrum < b.spectrum : a.form < b.form; });
}

// Reference index: DatabaseInput.generate_reference_peptide_index (pg/DatabaseInput.java:490-556)
static void build_reference(Ctx& C, const std::vector<std::string>& proteins) {
    const Params& P = C.P;
    std::set<std::string> seqs;
    for (auto& pr : proteins) digest_protein(P, pr, seqs);
    for (auto& t : C.target_seqs) {                                     // DatabaseInput.java:480-488; the targets were I -> L folded
        std::string f(t);                                               // when they were registered (add_target_peptides :47-51)
        for (auto& c : f) if (c == 'I') c = 'L';
        seqs.erase(f);
    }
    // GeneratePeptideformWorker.set_mass_range over SpectraInput.spectraMap (= spectra that produced a PSM)
    double mn = 1.7976931348623157e308, mx = 0;
    std::set<int> matched;
    for (auto& p : C.psms) matched.insert(p.spectrum);
    for (int si : matched) { double m = precursor_mass(C.spectra[si]); mn = std::min(mn, m); mx = std::max(mx, m); }
    double min_mass = mn - 250.0 - 5.0, max_mass = mx + 250.0 + 5.0;    // ModificationDB.leftMass/rightMass = -250/250
    if (C.forced_window) { min_mass = C.forced_lo; max_mass = C.forced_hi; }
    C.ref_seqs.assign(seqs.begin(), seqs.end());
    C.ref_forms.clear();
    for (size_t si = 0; si < C.ref_seqs.size(); ++si) {
        for (auto& f : calc_peptide_isoforms(P, C.ref_seqs[si])) {
            double m = peptide_mass(P, f);
            if (m > min_mass && m < max_mass) C.ref_forms.push_back({f, m, (int)si});
        }
    }
    // Canonical order: mass ascending; equal masses by (sequence order, form order) — the reference's own
    // equal-mass order comes from ConcurrentHashMap iteration and is unspecified.
    std::stable_sort(C.ref_forms.begin(), C.ref_forms.end(), [](const RefForm& a, const RefForm& b) { return a.mass < b.mass; });
    C.ref_index.clear();
    for (size_t i = 0; i < C.ref_forms.size(); ++i) C.ref_index[java_round(C.ref_forms[i].mass * 10.0)].push_back((int)i);
}

// Step 3: DBSearchWorker.run (pg/DBSearchWorker.java:25-113) + DatabaseInput.db_search (:672-758)
static void step3(Ctx& C) {
    const Params& P = C.P;
    std::map<int, double> spectra2score;                                // get_spectra2score(..., "max")
    for (auto& p : C.psms) { auto it = spectra2score.find(p.spectrum); if (it == spectra2score.end() || it->second < p.sc.score) spectra2score[p.spectrum] = p.sc.score; }
    std::vector<int> titles; for (auto& kv : spectra2score) titles.push_back(kv.first);
    struct Res { int n_matched = 0, n_better = 0; uint64_t pairs = 0; std::vector<pq_competitor> rows; };
    std::vector<Res> res(titles.size());
    parallel_for(C.threads, (int64_t)titles.size(), [&](int64_t ti) {
        int si = titles[ti];
        const Spectrum& sp = C.spectra[si];
        double target = spectra2score[si];
        Res& R = res[ti];
        double exp_mass = precursor_mass(sp);
        double bestScore = -1; pq_competitor bestRow{}; bool haveBest = false;   // ScoreResult.score default -1
        for (auto& im : isotope_masses(P, exp_mass)) {
            double mass = im.first;
            long va = java_round(mass * 10.0);
            for (long i = va - 1; i <= va + 1; ++i) {
                auto it = C.ref_index.find(i);
                if (it == C.ref_index.end()) continue;
                bool spectra_matched = false;
                for (int rf : it->second) {
                    double pm = C.ref_forms[rf].mass;
                    if (in_window(P, pm, mass)) {
                        spectra_matched = true;
                        R.n_matched++;
                        Score sc = score_peptide_to_spectrum(P, C.ref_forms[rf].pep, sp, C.H(si));
                        R.pairs++;
                        double tmp = sc.score;
                        pq_competitor row{si, rf, -1, -1, tmp, pm};
                        if (bestScore < tmp) { bestScore = tmp; bestRow = row; haveBest = true; }
                        if (target <= tmp) { R.n_better++; if (tmp >= 20) R.rows.push_back(row); }
                        if (P.p.fast && R.n_better >= 1) break;
                    } else if (spectra_matched) break;
                }
            }
        }
        if (haveBest && bestScore > 0) {
            bool dup = false;
            for (auto& r : R.rows) if (r.ref_form == bestRow.ref_form && r.score == bestRow.score) dup = true;
            if (!dup) R.rows.push_back(bestRow);
        }
    });
    C.comp3.clear();
    std::map<int, int> t2i; for (size_t i = 0; i < titles.size(); ++i) t2i[titles[i]] = (int)i;
    for (auto& R : res) { C.pairs[1] += R.pairs; for (auto& r : R.rows) C.comp3.push_back(r); }
    for (auto& p : C.psms) {
        Res& R = res[t2i[p.spectrum]];
        p.total_db = R.n_matched; p.n_db = R.n_better;
        if (R.n_better >= 1) p.valid = false;                           // DatabaseInput.java:747-750
    }
}

// Step 4: StatisticScoreCalc.run (pg/StatisticScoreCalc.java:19-63) + StatisticScoreWorker.run (:36-63)
static void step4(Ctx& C) {
    const Params& P = C.P;
    std::set<int> need;                                                 // PeptideInputs with >= 1 valid PSM
    for (auto& p : C.psms) if (p.valid) need.insert(C.target_form_seq[p.form]);
    for (int ts : need) C.shuffles[ts] = generate_random_peptides_fast(C.target_seqs[ts], P.p.n_random, P.p.shuffle_seed);
    std::vector<int> work;
    for (size_t i = 0; i < C.psms.size(); ++i) if (C.psms[i].valid) work.push_back((int)i);
    std::vector<uint64_t> cnt(work.size(), 0);
    parallel_for(C.threads, (int64_t)work.size(), [&](int64_t wi) {
        Psm& p = C.psms[work[wi]];
        const std::vector<std::string>& rp = C.shuffles[C.target_form_seq[p.form]];
        const Spectrum& sp = C.spectra[p.spectrum];
        int nRand = 0;
        for (const std::string& rs : rp) {
            Peptide rpep = get_modification_peptide(P, C.target_forms[p.form], rs);
            Score sc = score_peptide_to_spectrum(P, rpep, sp, C.H(p.spectrum));
            cnt[wi]++;
            if (sc.score >= p.sc.score) nRand++;
        }
        p.p_value = 1.0 * (nRand + 1) / ((int)rp.size() + 1);
        p.total_random = (int)rp.size(); p.n_random = nRand;
    });
    for (auto c : cnt) C.pairs[2] += c;
}

// RankPSM.rankPSMs / RankPSMCompare (pg/RankPSM.java:36-148) + CParameter.passed_psm_validation (:913-940)
static std::string mod_string(const Ctx& C, const Peptide& pep) {       // ModificationDB.getModificationString :1695-1708
    if (pep.mods.empty()) return "-";
    std::string s;
    for (size_t i = 0; i < pep.mods.size(); ++i) {
        char buf[96];
        snprintf(buf, sizeof buf, "%d@%d[%.4f]", C.P.mods[pep.mods[i].mod].id, pep.mods[i].site, C.P.mods[pep.mods[i].mod].delta);
        if (i) s += ";";
        s += buf;
    }
    return s;
}
static double tol_ppm_of(const Psm& p) { return (1.0e6) * (p.pep_mass - p.exp_mass) / p.pep_mass; }   // CParameter.get_mass_error
static void rank_and_validate(Ctx& C) {
    std::map<int, std::vector<int>> by;
    for (size_t i = 0; i < C.psms.size(); ++i) by[C.psms[i].spectrum].push_back((int)i);
    for (auto& kv : by) {
        std::vector<int>& v = kv.second;
        std::stable_sort(v.begin(), v.end(), [&](int ia, int ib) {
            const Psm& a = C.psms[ia]; const Psm& b = C.psms[ib];
            if (a.sc.score != b.sc.score) return a.sc.score > b.sc.score;
            if (a.p_value != b.p_value) return a.p_value < b.p_value;
            double ta = std::fabs(tol_ppm_of(a)), tb = std::fabs(tol_ppm_of(b));
            if (ta != tb) return ta < tb;
            std::string sa = C.target_forms[a.form].seq + "|" + mod_string(C, C.target_forms[a.form]);
            std::string sb = C.target_forms[b.form].seq + "|" + mod_string(C, C.target_forms[b.form]);
            return sa < sb;
        });
        for (size_t r = 0; r < v.size(); ++r) C.psms[v[r]].rank = (int)r + 1;
    }
    for (auto& p : C.psms) {
        int len = (int)C.target_forms[p.form].seq.size();
        bool pv = (len <= 8 && p.p_value <= 0.05) || (len > 8 && p.p_value <= 0.01);   // CParameter.java:84-94
        bool ok = pv && p.n_db == 0 && p.rank == 1;
        if (p.n_ptm >= 0) ok = ok && p.n_ptm == 0;
        p.confident = ok ? 1 : 0;
    }
}

// Step 5: ModificationDB.doPTMValidation (:1320-1466) -> ModPepQueryAndScoringWorker.do_score (:149-276),
// ModificationDB.getCandidatePTMs (:263-326), n_ptm from PSMT.summariseResult (PSMT:920-943).
static void step5(Ctx& C, const std::vector<pq_ptm>& ptms) {
    const Params& P = C.P;
    // spectra needing validation: PSMs passing passed_psm_validation(len,p,n_db,rank); spectra2score = MIN score
    std::map<int, double> spectra2score;
    for (auto& p : C.psms) {
        int len = (int)C.target_forms[p.form].seq.size();
        bool pv = (len <= 8 && p.p_value <= 0.05) || (len > 8 && p.p_value <= 0.01);
        if (!(pv && p.n_db == 0 && p.rank == 1)) continue;
        auto it = spectra2score.find(p.spectrum);
        if (it == spectra2score.end() || it->second > p.sc.score) spectra2score[p.spectrum] = p.sc.score;
    }
    std::vector<int> titles; for (auto& kv : spectra2score) titles.push_back(kv.first);
    struct Res { uint64_t pairs = 0; std::vector<pq_competitor> rows; };
    std::vector<Res> res(titles.size());
    // an extra Mod slot per PTM so Peptide/fragment code can carry it
    Params PX = P;
    int base = (int)PX.mods.size();
    for (auto& t : ptms) { Mod m; m.position = t.position; m.residues = t.residue ? std::string(1, t.residue) : std::string(); m.delta = t.delta; m.id = t.id; PX.mods.push_back(m); }
    parallel_for(C.threads, (int64_t)titles.size(), [&](int64_t ti) {
        int si = titles[ti];
        const Spectrum& sp = C.spectra[si];
        double target = spectra2score[si];
        Res& R = res[ti];
        double precursorMass = precursor_mass(sp);
        bool stop = false;                                              // ModificationDB.validation_status (fast mode)
        for (size_t ti2 = 0; ti2 < ptms.size() && !stop; ++ti2) {
            const pq_ptm& t = ptms[ti2];
            if (t.position >= PQ_MOD_PROT_NTERM) continue;              // :1037-1040
            if (!(t.delta >= -250.0 && t.delta <= 250.0)) continue;     // :1042 leftMass/rightMass
            const Mod& M = PX.mods[base + (int)ti2];
            for (auto& im : isotope_masses(P, precursorMass)) {
                if (stop) break;
                double mono = im.first;
                double delta_mass = mono - t.delta;
                long ind = java_round(10.0 * delta_mass);
                bool exceed = false;
                bool spectra_matched = false;                           // NOT reset per bucket (:163-165)
                for (long pi = ind - 1; pi <= ind + 1 && !stop; ++pi) {
                    auto it = C.ref_index.find(pi);
                    if (it == C.ref_index.end()) continue;
                    for (int rf : it->second) {
                        double pep_mass = C.ref_forms[rf].mass + t.delta;
                        double d = P.p.tol_ppm ? (1.0e6) * (pep_mass - mono) / pep_mass : pep_mass - mono;
                        double del = std::fabs(d);
                        if (del <= P.p.tol) spectra_matched = true;
                        else { if (spectra_matched) { exceed = true; break; } continue; }
                        const Peptide& base_pep = C.ref_forms[rf].pep;
                        if (t.residue && base_pep.seq.find(t.residue) == std::string::npos) continue;   // :207-209
                        std::set<int> occupied; for (auto& mm : base_pep.mods) occupied.insert(mm.site);
                        bool better = false;
                        for (int site : possible_sites(base_pep.seq, M)) {
                            if (occupied.count(site)) continue;
                            Peptide mp = base_pep; mp.mods.push_back({base + (int)ti2, site});
                            Score sc = score_peptide_to_spectrum(PX, mp, sp, C.H(si));
                            R.pairs++;
                            if (sc.score >= P.p.min_score) R.rows.push_back({si, rf, (int)ti2, site, sc.score, pep_mass});
                            if (P.p.fast) {
                                if (P.p.hc ? sc.score >= target : sc.score > target) { better = true; break; }
                            }
                        }
                        if (better) { stop = true; break; }
                    }
                    if (exceed) break;
                }
            }
        }
    });
    C.comp5.clear();
    std::map<int, int> t2i; for (size_t i = 0; i < titles.size(); ++i) t2i[titles[i]] = (int)i;
    for (size_t i = 0; i < titles.size(); ++i) {
        C.pairs[3] += res[i].pairs;
        // ptm.txt rows: score >= spectra2score (ModificationDB.save :1660-1663; the extra "best row" is not a competitor)
        for (auto& r : res[i].rows) if (r.score >= spectra2score[titles[i]]) C.comp5.push_back(r);
    }
    std::sort(C.comp5.begin(), C.comp5.end(), [](const pq_competitor& a, const pq_competitor& b) {
        if (a.spectrum != b.spectrum) return a.spectrum < b.spectrum;
        if (a.ref_form != b.ref_form) return a.ref_form < b.ref_form;
        if (a.ptm != b.ptm) return a.ptm < b.ptm;
        return a.ptm_site < b.ptm_site;
    });
    // PSMT.summariseResult (:920-1038): validated PSMs get n_ptm = #rows scoring > (>= with -hc) the PSM, others -1
    for (auto& p : C.psms) {
        p.n_ptm = -1;
        int len = (int)C.target_forms[p.form].seq.size();
        bool pv = (len <= 8 && p.p_value <= 0.05) || (len > 8 && p.p_value <= 0.01);
        if (!(pv && p.n_db == 0 && p.rank == 1)) continue;
        int n = 0;
        auto it = t2i.find(p.spectrum);
        if (it != t2i.end()) for (auto& r : res[it->second].rows) if (P.p.hc ? r.score >= p.sc.score : r.score > p.sc.score) n++;
        p.n_ptm = n;
    }
    rank_and_validate(C);
}

static void fill_form(const Params& P, const Peptide& pep, int seq_index, double mass, pq_form* o) {
    memset(o, 0, sizeof *o);
    o->seq_index = seq_index; o->length = (int)pep.seq.size(); o->mass = mass;
    o->n_mods = (int)std::min<size_t>(pep.mods.size(), PQ_MAX_FORM_MODS);
    for (int i = 0; i < o->n_mods; ++i) {
        o->mod_site[i] = (uint8_t)pep.mods[i].site;
        int mi = pep.mods[i].mod;
        o->mod_index[i] = (int8_t)(mi < P.n_var ? mi : -(mi - P.n_var) - 1);
    }
    strncpy(o->seq, pep.seq.c_str(), PQ_MAX_PEP_LEN);
}
static Peptide form_to_peptide(const Params& P, const pq_form& f) {
    Peptide p; p.seq = f.seq;
    for (int i = 0; i < f.n_mods; ++i) {
        int mi = f.mod_index[i] >= 0 ? f.mod_index[i] : P.n_var + (-(int)f.mod_index[i] - 1);
        p.mods.push_back({mi, (int)f.mod_site[i]});
    }
    return p;
}

}  // namespace orc

// ---------------------------------------------------------------------------------------------------
// C entry points for ctypes (same shapes as the product ABI, prefix orc_)
// ---------------------------------------------------------------------------------------------------
using namespace orc;
extern "C" {

int32_t orc_java_random_ints(int64_t seed, int32_t bound, int32_t n, int32_t* out) {
    JavaRandom r(seed);
    for (int i = 0; i < n; ++i) out[i] = bound > 0 ? r.nextInt(bound) : r.next(31);
    return 0;
}
double orc_log10(double x) { return fd_log10(x); }
double orc_log(double x) { return fd_log(x); }
double orc_log10_factorial(int32_t n) { return log10_factorial(n); }
double orc_factorial_log(int32_t n) { return factorial_log(n); }
double orc_residue_mass(char c) { return g_res_mass[(int)c]; }
double orc_proton(void) { return g_a.proton; }
void orc_set_assumption(int32_t which, int32_t value) {
    switch (which) {
        case 3: g_a.a3_closed_interval = value != 0; break;
        case 4: g_a.a4_tie_smaller_error = value != 0; break;
        case 8: g_a.a8_mod_losses_in_prediction = value != 0; break;
        default: break;
    }
}

// shuffles of one sequence, NUL separated
int32_t orc_shuffles(const char* seq, int32_t num, int64_t seed, char* out, int64_t cap, int32_t* n) {
    auto v = generate_random_peptides_fast(seq, num, seed);
    int64_t L = (int64_t)strlen(seq) + 1;
    *n = (int32_t)v.size();
    if ((int64_t)v.size() * L > cap) return PQ_ERR_CAPACITY;
    for (size_t i = 0; i < v.size(); ++i) memcpy(out + i * L, v[i].c_str(), L);
    return 0;
}

struct orc_ctx { Ctx c; };

int32_t orc_create(const pq_params* p, int32_t threads, orc_ctx** out) {
    orc_ctx* o = new orc_ctx();
    o->c.P = make_params(*p);
    o->c.threads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    *out = o;
    return 0;
}
int32_t orc_destroy(orc_ctx* o) { delete o; return 0; }

int32_t orc_load_spectra(orc_ctx* o, int64_t n, const double* pmz, const int32_t* charge, const uint64_t* off, const double* mz, const double* inten) {
    o->c.spectra.resize(n);
    for (int64_t i = 0; i < n; ++i) {
        Spectrum& s = o->c.spectra[i];
        s.prec_mz = pmz[i]; s.charge = charge[i]; s.given_charge = charge[i];
        s.mz.assign(mz + off[i], mz + off[i + 1]); s.inten.assign(inten + off[i], inten + off[i + 1]);
    }
    return 0;
}
int32_t orc_set_targets(orc_ctx* o, int32_t n, const uint32_t* off, const char* res) {
    Ctx& C = o->c;
    C.target_seqs.clear(); C.target_forms.clear(); C.target_form_seq.clear(); C.target_form_mass.clear();
    for (int i = 0; i < n; ++i) {
        std::string s(res + off[i], res + off[i + 1]);
        C.target_seqs.push_back(s);
        for (auto& f : calc_peptide_isoforms(C.P, s)) {
            C.target_forms.push_back(f); C.target_form_seq.push_back(i); C.target_form_mass.push_back(peptide_mass(C.P, f));
        }
    }
    return 0;
}
int32_t orc_get_target_forms(orc_ctx* o, pq_form* out, int64_t cap, int64_t* n) {
    Ctx& C = o->c; *n = (int64_t)C.target_forms.size();
    if (!out) return 0;
    if (cap < *n) return PQ_ERR_CAPACITY;
    for (size_t i = 0; i < C.target_forms.size(); ++i) fill_form(C.P, C.target_forms[i], C.target_form_seq[i], C.target_form_mass[i], out + i);
    return 0;
}
int32_t orc_build_reference(orc_ctx* o, int32_t n, const uint64_t* off, const char* res) {
    std::vector<std::string> prots;
    for (int i = 0; i < n; ++i) prots.emplace_back(res + off[i], res + off[i + 1]);
    build_reference(o->c, prots);
    return 0;
}
int32_t orc_get_reference_forms(orc_ctx* o, pq_form* out, int64_t cap, int64_t* n) {
    Ctx& C = o->c; *n = (int64_t)C.ref_forms.size();
    if (!out) return 0;
    if (cap < *n) return PQ_ERR_CAPACITY;
    for (size_t i = 0; i < C.ref_forms.size(); ++i) fill_form(C.P, C.ref_forms[i].pep, C.ref_forms[i].seq_index, C.ref_forms[i].mass, out + i);
    return 0;
}
int32_t orc_get_reference_sequences(orc_ctx* o, char* out, int64_t cap, int64_t* n_seq, int64_t* n_bytes) {
    Ctx& C = o->c; int64_t b = 0;
    for (auto& s : C.ref_seqs) b += (int64_t)s.size() + 1;
    *n_seq = (int64_t)C.ref_seqs.size(); *n_bytes = b;
    if (!out) return 0;
    if (cap < b) return PQ_ERR_CAPACITY;
    char* w = out;
    for (auto& s : C.ref_seqs) { memcpy(w, s.c_str(), s.size() + 1); w += s.size() + 1; }
    return 0;
}
int32_t orc_step2_match_targets(orc_ctx* o, int64_t* n) { step2(o->c); *n = (int64_t)o->c.psms.size(); return 0; }
int32_t orc_step3_competitors(orc_ctx* o) { step3(o->c); return 0; }
int32_t orc_step4_shuffles(orc_ctx* o) { step4(o->c); return 0; }
int32_t orc_rank_and_validate(orc_ctx* o) { rank_and_validate(o->c); return 0; }
int32_t orc_step5_ptm(orc_ctx* o, int32_t n, const pq_ptm* t) { step5(o->c, std::vector<pq_ptm>(t, t + n)); return 0; }
int32_t orc_get_psms(orc_ctx* o, pq_psm* out, int64_t cap, int64_t* n) {
    Ctx& C = o->c; *n = (int64_t)C.psms.size();
    if (!out) return 0;
    if (cap < *n) return PQ_ERR_CAPACITY;
    for (size_t i = 0; i < C.psms.size(); ++i) {
        const Psm& p = C.psms[i]; pq_psm& q = out[i]; memset(&q, 0, sizeof q);
        q.spectrum = p.spectrum; q.form = p.form; q.score = p.sc.score; q.exp_mass = p.exp_mass; q.pep_mass = p.pep_mass;
        q.tol_ppm = tol_ppm_of(p); q.isotope_error = p.iso; q.count_b = p.sc.a; q.count_y = p.sc.b;
        q.n_db = p.n_db; q.total_db = p.total_db; q.n_random = p.n_random; q.total_random = p.total_random;
        if (p.rank >= 2) { q.n_db = -1; q.total_db = -1; }                  // psm_rank.txt: RankPSM.rankPSMs (pg/RankPSM.java:89-92)
        q.valid = p.valid ? 1 : 0; q.p_value = p.p_value; q.rank = p.rank; q.n_ptm = p.n_ptm; q.confident = p.confident; q.charge = p.charge;
    }
    return 0;
}
int32_t orc_get_competitors(orc_ctx* o, int32_t step, pq_competitor* out, int64_t cap, int64_t* n) {
    auto& v = step == 5 ? o->c.comp5 : o->c.comp3; *n = (int64_t)v.size();
    if (!out) return 0;
    if (cap < *n) return PQ_ERR_CAPACITY;
    memcpy(out, v.data(), v.size() * sizeof(pq_competitor));
    return 0;
}
// PeptideSearchMT.add_delta_score (pg/PeptideSearchMT.java:1365-1475): per row of psm_rank.txt, the target's score minus the
// best score among that spectrum's rows of detail.txt (ref_delta_score) / ptm.txt (mod_delta_score); the plain score when
// the spectrum has no row there.
int32_t orc_get_delta_scores(orc_ctx* o, double* ref_delta, double* mod_delta, int64_t cap, int64_t* n) {
    Ctx& C = o->c; *n = (int64_t)C.psms.size();
    if (!ref_delta || !mod_delta) return 0;
    if (cap < *n) return PQ_ERR_CAPACITY;
    std::map<int32_t, double> ref_best, mod_best;
    auto fill = [](const std::vector<pq_competitor>& rows, std::map<int32_t, double>& best) {
        for (const pq_competitor& r : rows) {
            auto it = best.find(r.spectrum);
            if (it == best.end()) best[r.spectrum] = r.score;
            else if (it->second < r.score) it->second = r.score;             // :1389-1392
        }
    };
    fill(C.comp3, ref_best); fill(C.comp5, mod_best);
    for (size_t i = 0; i < C.psms.size(); ++i) {
        const Psm& p = C.psms[i];
        auto a = ref_best.find(p.spectrum); auto b = mod_best.find(p.spectrum);
        ref_delta[i] = a != ref_best.end() ? p.sc.score - a->second : p.sc.score;   // :1456-1463
        mod_delta[i] = b != mod_best.end() ? p.sc.score - b->second : p.sc.score;
    }
    return 0;
}
// Step-1 novelty filter with the default FM-index mapping (pg/InputProcessor.java:416-427, pg/PepMapping.java:60-107):
// hasProtein(seq) = seq occurs in some protein with I and L indistinguishable (A13: compomics FMIndex,
// MatchingType.indistiguishableAminoAcids, treated as exact substring search after I -> L; case folded).
int32_t orc_find_in_proteome(int32_t n_seq, const uint32_t* seq_off, const char* res, int32_t n_prot, const uint64_t* prot_off, const char* prot_res,
                             int32_t* nhit) {
    auto fold = [](char c) { char u = (char)toupper((unsigned char)c); return u == 'I' ? 'L' : u; };
    std::vector<std::string> prots((size_t)n_prot);
    for (int32_t i = 0; i < n_prot; ++i) { std::string& p = prots[(size_t)i]; for (uint64_t k = prot_off[i]; k < prot_off[i + 1]; ++k) p.push_back(fold(prot_res[k])); }
    for (int32_t t = 0; t < n_seq; ++t) {
        std::string q; for (uint32_t k = seq_off[t]; k < seq_off[t + 1]; ++k) q.push_back(fold(res[k]));
        nhit[t] = 0;
        for (const std::string& p : prots) if (p.find(q) != std::string::npos) { nhit[t] = 1; break; }
    }
    return 0;
}
// cost model of the timed baseline: 0 = as written (pre-processing redone per pair), 1 = hoisted (once per spectrum)
int32_t orc_set_cost_model(orc_ctx* o, int32_t hoisted) { o->c.hoisted = hoisted != 0; return 0; }
// force the mass range of the reference table (open interval), e.g. the range of a larger run a sample is compared with
int32_t orc_set_reference_window(orc_ctx* o, int32_t on, double lo, double hi) { o->c.forced_window = on != 0; o->c.forced_lo = lo; o->c.forced_hi = hi; return 0; }
// SpectraInput.pep2targeted_spectra (PSMT:264-309): pairs (target sequence, spectrum); n = 0 clears the list
int32_t orc_set_targeted_psms(orc_ctx* o, int64_t n, const int32_t* target_seq, const int32_t* spectrum) {
    Ctx& C = o->c;
    C.targeted.clear(); C.pep2targeted.clear();
    for (int64_t i = 0; i < n; ++i) {
        if (target_seq[i] < 0 || (size_t)target_seq[i] >= C.target_seqs.size()) return PQ_ERR_BAD_ARG;
        C.targeted.push_back({target_seq[i], spectrum[i]});
        C.pep2targeted[C.target_seqs[(size_t)target_seq[i]]][spectrum[i]] = 0;
    }
    return 0;
}
// PeptideSearchMT.export_targeted_psm_status (PSMT:1484-1589) + TargetedPSM.get_psm_status (pg/TargetedPSM.java:29-46): one
// PsmType ordinal per targeted pair (0 no_spectrum_match, 2 low_score, 3 high_p_value, 4 better_ref_pep, 5 better_ref_pep_with_mod,
// 6 not_top_rank, 7 confident; 8 invalid_peptide is the Java host's call, it owns the Step-1 verdict).
int32_t orc_get_targeted_status(orc_ctx* o, int32_t* out, int64_t cap, int64_t* n) {
    Ctx& C = o->c; *n = (int64_t)C.targeted.size();
    if (!out) return 0;
    if (cap < *n) return PQ_ERR_CAPACITY;
    std::map<std::pair<std::string, int>, const Psm*> best;             // (peptide, spectrum) -> row with the highest rank (:1513-1527)
    for (const Psm& p : C.psms) {
        auto key = std::make_pair(C.target_forms[(size_t)p.form].seq, p.spectrum);
        auto it = best.find(key);
        if (it == best.end() || it->second->rank > p.rank) best[key] = &p;
    }
    for (size_t i = 0; i < C.targeted.size(); ++i) {
        const std::string& seq = C.target_seqs[(size_t)C.targeted[i].first];
        auto it = best.find(std::make_pair(seq, C.targeted[i].second));
        if (it != best.end()) {
            const Psm& p = *it->second;
            const int len = (int)seq.size();
            const bool pv = (len <= 8 && p.p_value <= 0.05) || (len > 8 && p.p_value <= 0.01);      // passed_p_value_validation
            out[i] = p.rank >= 2 ? 6 : (p.n_db >= 1 ? 4 : (!pv ? 3 : (p.n_ptm >= 1 ? 5 : 7)));
        } else {
            out[i] = C.pep2targeted[seq][C.targeted[i].second] == 0 ? 0 : 2;
        }
    }
    return 0;
}
int32_t orc_get_pairs(orc_ctx* o, uint64_t* out4) { for (int i = 0; i < 4; ++i) out4[i] = o->c.pairs[i]; return 0; }

// scorePeptide2Spectrum on an explicit pair list (multi-threaded; the CPU-baseline timing leg uses this too)
int32_t orc_score_pairs(orc_ctx* o, int64_t n, const int32_t* spectrum_index, const pq_form* forms, double* score, int32_t* ca, int32_t* cb) {
    Ctx& C = o->c;
    parallel_for(C.threads, n, [&](int64_t i) {
        Peptide p = form_to_peptide(C.P, forms[i]);
        Score s = score_peptide_to_spectrum(C.P, p, C.spectra[spectrum_index[i]]);
        score[i] = s.score; if (ca) ca[i] = s.a; if (cb) cb[i] = s.b;
    });
    return 0;
}
// form of a shuffle: getModificationPeptide(target form, shuffled sequence)
int32_t orc_shuffle_form(orc_ctx* o, int32_t target_form, const char* seq, pq_form* out) {
    Ctx& C = o->c;
    Peptide p = get_modification_peptide(C.P, C.target_forms[target_form], seq);
    fill_form(C.P, p, C.target_form_seq[target_form], peptide_mass(C.P, p), out);
    return 0;
}
// spectrum pre-processing products, for kernel K0 parity: hyperscore survivors (mz, normalised intensity),
// MVH survivors (mz, class), annotator intensity limit.
int32_t orc_preprocess(orc_ctx* o, int32_t si, int32_t method, double* out_mz, double* out_val, int32_t cap, int32_t* n, double* limit) {
    Ctx& C = o->c; const Spectrum& sp = C.spectra[si];
    JMatch m(C.P, sp);
    m.js.resetPeaks();
    *limit = intensity_limit(sp, 0.05);
    if (method == 1) {
        m.isotopePass(0.95); m.removeParentPeak(); m.removeLowMasses(); m.removeLowIntensityPeak();
        m.isotopePass(1.5); m.filterByPeakCount(50); m.normalizePeaks(100.0);
        m.js.getPeaks(); m.js.sortByMZ();
        *n = (int32_t)m.js.peaks.size();
        if (*n > cap) return PQ_ERR_CAPACITY;
        for (int i = 0; i < *n; ++i) { out_mz[i] = m.js.peaks[i].mz; out_val[i] = m.js.peaks[i].inten; }
    } else {
        m.filterPeakByRange(200.0, 2000.0); m.removeParentPeak(); m.filterByTIC(); m.isotopePass(0.95); m.filterByPeakCount(300);
        int total = (int)m.js.getPeaks().size();
        if (total < 7) { *n = -1; return 0; }
        int npeak = (int)std::floor(1.0 * total / 7.0);
        m.js.sortByIntensity();
        auto& v = m.js.getPeaks();
        for (int i = 0; i < total; ++i) v[i].cls = i >= total - npeak ? 0 : (i >= total - 3 * npeak ? 1 : 2);
        m.js.sortByMZ();
        *n = total;
        if (*n > cap) return PQ_ERR_CAPACITY;
        for (int i = 0; i < *n; ++i) { out_mz[i] = v[i].mz; out_val[i] = (double)v[i].cls; }
    }
    return 0;
}
}  // extern "C"

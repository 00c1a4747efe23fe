// pq_common.h — constants, exact-arithmetic helpers and device data layouts shared by the kernels and the host side
// of libpepquery_b200.so.  Product code: nothing here comes from oracle/.
#pragma once
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

#include "../../include/pepquery_b200.h"

#define PQ_HD __host__ __device__ __forceinline__

namespace pq {

// ---- physical constants (compomics 5.0.39 values; SURVEY.md §8c, pinned by resources/top_modifications.tsv) ----
constexpr double kProton = 1.007276466812;      // ElementaryIon.proton (A1)
constexpr double kWater = 18.0105646837;        // H2O
constexpr double kIsotope = 1.00335;            // CParameter.isotope_error_mass (pg/CParameter.java:296)

// ---- peptide rows on device -------------------------------------------------------------------------------------
// One peptide form = one 64-byte row: byte0 length, byte1 N-term modification code, byte2 C-term modification code,
// byte3 reserved, bytes 4.. residue codes.  A residue code indexes ModTable::aa/delta: 0..25 plain 'A'..'Z',
// 26.. = (residue, modification) combinations of this search.
constexpr int kRowBytes = 64;
constexpr int kRowHeader = 4;
constexpr int kMaxLen = kRowBytes - kRowHeader;   // 60 residues
constexpr int kMaxCodes = 128;
constexpr int kMaxTermCodes = 16;

struct ModTable {
    double aa[kMaxCodes];         // residue mass of the code
    double delta[kMaxCodes];      // modification mass of the code (0 for plain residues)
    double nterm[kMaxTermCodes];  // terminal modification masses, index = code (0 = none)
    double cterm[kMaxTermCodes];
    double loss[kMaxCodes];       // modification-defined neutral loss of a residue code (MVH prediction), 0 = none
    double nterm_loss[kMaxTermCodes];
    double cterm_loss[kMaxTermCodes];
};

// ---- prepared spectrum blob (output of K0, input of K1) -----------------------------------------------------------
// Everything K1 needs about one spectrum, contiguous and 16-byte aligned so one cp.async.bulk stages it:
//   BlobHeader | ev f64[NE] | ans u16[NE pad 8] | pmz f64[E+1 pad 2] | pin f64[E pad 2] | pcode u16[E pad 8] | val f64[kValSlots] | cells u16[kCells + 8] | cbits u32[kCells / 32]
//
// The annotator's question "which matchable peak is the most intense one with |mz - t| <= itol" (JMatch.java:411-451,
// A2-A4) has a piecewise-constant answer in t.  K0 inverts the predicate once per spectrum: for every matchable peak j
// the set {t : fabs(fl(mz_j - t)) <= itol} is an interval of doubles [lo_j, hi_j] whose end points are found by testing
// the predicate itself on neighbouring doubles, so membership is bit-exact.  The sorted end points are the events:
//   ev    : ev[0] = -inf, then every lo_j (peak j enters at t >= lo_j) and nextup(hi_j) (leaves), ascending, then +inf
//   ans   : answer for ev[k] <= t < ev[k+1]: 0xFFFF = nothing to claim; 0x8000|a = two in-window peaks tie on intensity,
//           take the exact scan from peak a (pmz/pinfo); else the claim code of the winner: hyperscore = survivor id,
//           MVH = (class << 12) | survivor id (a winner that did not survive pre-processing, or an MVH winner outside
//           200..2000 Th, is already folded into 0xFFFF)
//   pmz   : m/z of the matchable peaks (intensity >= 5th-percentile limit), ascending, +inf after the last
//   pin   : raw intensity of the matchable peaks (only the tie scan reads it)
//   pcode : claim code of each matchable peak: (class << 12) | survivor id (0xFFF = the peak did not survive pre-processing)
//   val   : hyperscore: normalised intensity of survivor id; MVH: unused
//   cells : cells[c] = index of the last event below the lower edge of cell c of the monotone m/z -> cell map,
//           | kCellClean when the whole cell answers 0xFFFF (K1 then skips the event scan); one more entry follows the table,
//           always kCellClean: the precomputed cell lists of the reference rows (k1_ladder.cu) point unused slots at it
//   cbits : the kCellClean flags of cells[] once more as a bit map (bit c of word c / 32): a kernel that reads the blob from
//           global memory (Step 2, one or two pairs per spectrum) filters its ions from these 320 bytes instead of pulling the
//           5 KB table through L2
constexpr int kCellShift = 11;                              // 512 cells per binade
constexpr uint32_t kCellBase = 0x40600000u >> kCellShift;   // high word of 128.0
constexpr int kCells = 5 * 512;                             // [128, 4096) Th; below / above clamp to the first / last cell
constexpr int kValSlots = 64;
constexpr uint32_t kNoSurvivor = 0xFFFu;
constexpr uint32_t kAnsNone = 0xFFFFu;
constexpr uint32_t kAnsTie = 0x8000u;
constexpr uint32_t kCellClean = 0x8000u;     // cells[] flag: every t of the cell answers kAnsNone (no event inside, nothing to claim at its lower edge)
constexpr uint32_t kCellSentinel = 2u * kCells;   // byte offset of the always-clean entry behind the table
constexpr int kLadderCharges = 3;                // fragment charges held in the precomputed ladders of the reference rows (k1_ladder.cu)

struct BlobHeader {
    uint32_t n_peaks;        // E
    uint32_t n_events;       // NE = 2E + 2
    uint32_t n_survivors;    // hyperscore: distinct surviving m/z (<= 50); MVH: surviving peaks N (<= 300)
    uint32_t flags;          // bit0: MVH pre-processing succeeded (N >= 7)
    int32_t  cls_count[4];   // MVH intenClassCounts[0..2], [3] = bins - N
    uint32_t total_bytes;
    uint32_t raw_peaks;      // P_s, for the algorithmic-bytes counter
    uint32_t pad[2];
};
static_assert(sizeof(BlobHeader) == 48, "BlobHeader must stay 16-byte aligned");

struct BlobLayout { uint32_t ev, ans, pmz, pin, pcode, val, cells, cbits, total; };
PQ_HD BlobLayout blob_layout(uint32_t E) {
    BlobLayout L;
    const uint32_t ne = 2u * E + 2u;
    L.ev = (uint32_t)sizeof(BlobHeader);
    L.ans = L.ev + ne * 8u;
    L.pmz = L.ans + ((ne * 2u + 15u) & ~15u);
    L.pin = L.pmz + (((E + 1u) * 8u + 15u) & ~15u);
    L.pcode = L.pin + ((E * 8u + 15u) & ~15u);
    L.val = L.pcode + ((E * 2u + 15u) & ~15u);
    L.cells = L.val + kValSlots * 8u;
    L.cbits = L.cells + kCells * 2u + 16u;                      // behind the always-clean sentinel entry (padded to 16 bytes)
    L.total = L.cbits + kCells / 8u;
    return L;
}
PQ_HD uint32_t blob_bytes(uint32_t n_peaks) { return blob_layout(n_peaks).total; }

PQ_HD uint32_t hi32(double x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2hiint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (uint32_t)(u >> 32);
#endif
}
// Monotone map m/z -> cell: the high word of a positive double orders like the double itself.
PQ_HD int cell_of(double mz) {
    int c = ((int)hi32(mz) >> kCellShift) - (int)kCellBase;       // arithmetic shift: below 128 Th and for a negative mz (sign bit) c < 0 -> cell 0
#ifdef __CUDA_ARCH__
    return min(max(c, 0), kCells - 1);
#else
    return c < 0 ? 0 : (c >= kCells ? kCells - 1 : c);
#endif
}

// Cell for the "can anything be claimed here?" filter only: one unsigned minimum.  Everything outside [128, 4096) Th — below
// the table, negative, NaN — wraps or clamps to the LAST cell, whose clean flag K0 therefore clears unless cell 0 (which
// holds the events below 128 Th) is clean as well.
PQ_HD uint32_t cell_filter(double mz) {
    uint32_t c = (hi32(mz) >> kCellShift) - kCellBase;
#ifdef __CUDA_ARCH__
    return min(c, (uint32_t)(kCells - 1));
#else
    return c < (uint32_t)(kCells - 1) ? c : (uint32_t)(kCells - 1);
#endif
}

// ---- exact arithmetic ------------------------------------------------------------------------------------------------
// On the device every step is an explicit round-to-nearest intrinsic so that no FMA contraction can change a bit.
PQ_HD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
PQ_HD double dsub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
PQ_HD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
PQ_HD double ddiv(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
#ifdef __CUDACC__
// RN(x / 3) for a normal x, three FP64 operations instead of the division routine (the m/z of a triply charged fragment, once per
// ion).  r = RN(1/3) = (1/3)(1 - 2^-54); q0 = RN(x r) is within 1.5 ulp of x/3; rem = x - 3 q0 is exact (a multiple of ulp(q0) below
// 5 ulp); q0 + rem r = x/3 - (rem/3) 2^-54, i.e. x/3 moved by less than 2^-53 ulp, and x/3 is never closer than ulp/6 to a rounding
// boundary (a midpoint times 3 needs 55 bits, so x - 3 * midpoint is a non-zero multiple of ulp/2): the final rounding returns
// RN(x/3), the double __ddiv_rn(x, 3.0) returns (tests: pq_selftest compares the two on the device; tests/test_oracle_pins.py replays
// the three operations in exact rational arithmetic).
__device__ __forceinline__ double div3_rn(double x) {
    const double r = 0.33333333333333331;
    const double q0 = __dmul_rn(x, r);
    const double rem = __fma_rn(-3.0, q0, x);
    return __fma_rn(rem, r, q0);
}
#endif
PQ_HD double with_hi(double x, uint32_t hi) {
#ifdef __CUDA_ARCH__
    return __hiloint2double((int)hi, __double2loint(x));
#else
    uint64_t u; memcpy(&u, &x, 8); u = (u & 0xffffffffULL) | ((uint64_t)hi << 32); double r; memcpy(&r, &u, 8); return r;
#endif
}
PQ_HD uint32_t lo32(double x) {
#ifdef __CUDA_ARCH__
    return (uint32_t)__double2loint(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return (uint32_t)u;
#endif
}

// Natural log, the JDK's specified algorithm (StrictMath.log = fdlibm 5.3 e_log.c), positive finite normal inputs.
// java.lang.Math.log / log10 are what HyperscoreMatch.java:162-170 and commons-math MathUtils.factorialLog call.
PQ_HD double jlog(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    int32_t hx = (int32_t)hi32(x);
    int32_t k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    x = with_hi(x, (uint32_t)(hx | (i ^ 0x3ff00000)));
    k += (i >> 20);
    double f = dsub(x, 1.0);
    double dk = (double)k;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) { if (k == 0) return 0.0; return dadd(dmul(dk, ln2_hi), dmul(dk, ln2_lo)); }
        double R = dmul(dmul(f, f), dsub(0.5, dmul(0.33333333333333333, f)));
        if (k == 0) return dsub(f, R);
        return dsub(dmul(dk, ln2_hi), dsub(dsub(R, dmul(dk, ln2_lo)), f));
    }
    double s = ddiv(f, dadd(2.0, f));
    double z = dmul(s, s);
    i = hx - 0x6147a;
    double w = dmul(z, z);
    int32_t j = 0x6b851 - hx;
    double t1 = dmul(w, dadd(Lg2, dmul(w, dadd(Lg4, dmul(w, Lg6)))));
    double t2 = dmul(z, dadd(Lg1, dmul(w, dadd(Lg3, dmul(w, dadd(Lg5, dmul(w, Lg7)))))));
    i |= j;
    double R = dadd(t2, t1);
    if (i > 0) {
        double hfsq = dmul(dmul(0.5, f), f);
        if (k == 0) return dsub(f, dsub(hfsq, dmul(s, dadd(hfsq, R))));
        return dsub(dmul(dk, ln2_hi), dsub(dsub(hfsq, dadd(dmul(s, dadd(hfsq, R)), dmul(dk, ln2_lo))), f));
    } else {
        if (k == 0) return dsub(f, dmul(s, dsub(f, R)));
        return dsub(dmul(dk, ln2_hi), dsub(dsub(dmul(s, dsub(f, R)), dmul(dk, ln2_lo)), f));
    }
}
// Base-10 log (StrictMath.log10 = fdlibm e_log10.c), positive finite normal inputs.
PQ_HD double jlog10(double x) {
    const double ivln10 = 4.34294481903251816668e-01, log10_2hi = 3.01029995663611771306e-01,
                 log10_2lo = 3.69423907715893078616e-13;
    int32_t hx = (int32_t)hi32(x);
    if (hx >= 0x7ff00000) return dadd(x, x);          // +inf / NaN propagate like fdlibm
    int32_t k = (hx >> 20) - 1023;
    int32_t i = (int32_t)(((uint32_t)k & 0x80000000u) >> 31);
    hx = (hx & 0x000fffff) | ((0x3ff - i) << 20);
    double y = (double)(k + i);
    x = with_hi(x, (uint32_t)hx);
    double z = dadd(dmul(y, log10_2lo), dmul(ivln10, jlog(x)));
    return dadd(z, dmul(y, log10_2hi));
}

// Math.round(double) for the magnitudes used here (bucket keys): floor(x + 0.5)
PQ_HD long long jround(double x) {
#ifdef __CUDA_ARCH__
    return (long long)floor(__dadd_rn(x, 0.5));
#else
    return (long long)__builtin_floor(x + 0.5);
#endif
}

}  // namespace pq

// pq_mgf.cu — MGF text parser into CSR buffers.  Replaces compomics MgfFileIterator as used by
// SpectraInput.readMSMSfast (pg/SpectraInput.java:150-195); the writer-side format is index/IndexWorker.java:431-457
// (BEGIN IONS / TITLE / PEPMASS / CHARGE / "mz intensity" lines / END IONS).  Host code, no device work.
//
// Once scoring runs on the GPU the text parse is the upstream bottleneck (1 M spectra ~ 3 GB of text), so the parser is
// parallel: the text is cut at "BEGIN IONS" lines into one chunk per host thread, every chunk is parsed independently
// (count pass when the caller asks for sizes, fill pass into the caller's arrays at prefix-summed offsets), and decimal
// numbers take Clinger's exact fast path — up to 19 digits collected in an integer, one correctly rounded multiply or
// divide by an exactly representable power of ten when the integer is below 2^53 and |exponent| <= 22 — which returns the
// same double as strtod / Java's Double.parseDouble; anything else falls back to strtod.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "pq_common.h"

namespace {

const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                           1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};

// Parses one decimal number at p (p < end).  *next = first character after it; *next == p when nothing was parsed.
inline double parse_double(const char* p, const char* end, const char** next) {
    const char* s = p;
    while (s < end && (*s == ' ' || *s == '\t')) ++s;              // strtod skips leading blanks
    bool neg = false;
    if (s < end && (*s == '+' || *s == '-')) { neg = *s == '-'; ++s; }
    if (s < end && ((*s | 0x20) == 'i' || (*s | 0x20) == 'n')) {   // "inf" / "nan": strtod's business
        char buf[64]; size_t k = 0; while (k < 63 && p + k < end && p[k] != '\n' && p[k] != '\r') { buf[k] = p[k]; ++k; }
        buf[k] = 0;
        char* e = nullptr; double v = strtod(buf, &e);
        *next = p + (e - buf);
        return v;
    }
    uint64_t m = 0; int digits = 0, dropped = 0, frac = 0; bool any = false;
    while (s < end && *s >= '0' && *s <= '9') { any = true; if (digits < 19) { m = m * 10 + (uint64_t)(*s - '0'); if (m) ++digits; } else ++dropped; ++s; }
    if (s < end && *s == '.') {
        ++s;
        while (s < end && *s >= '0' && *s <= '9') { any = true; if (digits < 19) { m = m * 10 + (uint64_t)(*s - '0'); if (m) ++digits; ++frac; } ++s; }
    }
    if (!any) { *next = p; return 0.0; }
    int e10 = dropped - frac;
    bool simple = true;
    if (s < end && (*s == 'e' || *s == 'E')) {
        const char* t = s + 1; bool eneg = false;
        if (t < end && (*t == '+' || *t == '-')) { eneg = *t == '-'; ++t; }
        if (t < end && *t >= '0' && *t <= '9') {
            int ev = 0;
            while (t < end && *t >= '0' && *t <= '9') { if (ev < 100000) ev = ev * 10 + (*t - '0'); ++t; }
            e10 += eneg ? -ev : ev; s = t;
        }
    }
    // hex floats ("0x..."), "inf"/"nan", more than 19 significant digits, big exponents: let strtod decide
    if (s < end && (*s == 'x' || *s == 'X' || *s == 'p' || *s == 'P')) simple = false;
    if (dropped || m >= (1ull << 53) || e10 < -22 || e10 > 22) simple = false;
    if (!simple) {
        char buf[400]; size_t n = (size_t)std::min<ptrdiff_t>(end - p, 399);
        size_t k = 0; while (k < n && p[k] != '\n' && p[k] != '\r') ++k;          // never read past the line
        memcpy(buf, p, k); buf[k] = 0;
        char* e = nullptr; double v = strtod(buf, &e);
        *next = p + (e - buf);
        return v;
    }
    double v = (double)m;
    v = e10 < 0 ? v / kPow10[-e10] : v * kPow10[e10];
    *next = s;
    return neg ? -v : v;
}

// The common peak-line number: digits [. digits], at most 18 digits in all, no sign / exponent.  Same value as parse_double
// (the integer below 2^53 divided by an exact power of ten: one correctly rounded IEEE division); false = not that shape,
// the caller takes the general path from the same position.
inline bool fast_number(const char*& s, const char* end, double& v) {
    const char* p = s;
    uint64_t m = 0; int nd = 0, frac = 0;
    while (p < end && (unsigned)(*p - '0') <= 9u) { m = m * 10 + (uint64_t)(*p - '0'); ++nd; ++p; }
    if (p < end && *p == '.') {
        ++p;
        const char* f0 = p;
        while (p < end && (unsigned)(*p - '0') <= 9u) { m = m * 10 + (uint64_t)(*p - '0'); ++p; }
        frac = (int)(p - f0); nd += frac;
    }
    if (nd == 0 || nd > 18 || m >= (1ull << 53)) return false;
    if (p < end) { const char c = *p; if (c != ' ' && c != '\t' && c != '\r' && c != '\n') return false; }     // 'e', 'x', letters ...: general path
    v = frac ? (double)m / kPow10[frac] : (double)m;
    s = p;
    return true;
}

struct Out { double* prec; int32_t* charge; uint64_t* off; double* mz; double* inten; };

// The sequential state machine over [p, end): a chunk always starts at a BEGIN IONS line (or at the start of the text).
// fill == nullptr: count only.  s0 / p0: index of the chunk's first spectrum / peak in the caller's arrays.
void parse_chunk(const char* p, const char* end, const Out* fill, int64_t s0, uint64_t p0, int64_t* n_spec, uint64_t* n_peak) {
    int64_t ns = 0; uint64_t np = 0, np_begin = 0;
    bool in_ions = false; double pmz = 0.0; int32_t z = 0;
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        const char* q = p;
        while (q < eol && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
        const size_t len = (size_t)(eol - q);
        if (len > 0) {
            const char c = *q;
            if (in_ions && ((c >= '0' && c <= '9') || c == '.')) {
                if (!fill) { if (c != '.' || (len > 1 && q[1] >= '0' && q[1] <= '9')) ++np; }      // what the fill pass will accept
                else {
                    const char* e1 = q; double a;
                    if (!fast_number(e1, eol, a)) a = parse_double(q, eol, &e1);
                    if (e1 != q) {
                        const char* t = e1; while (t < eol && (*t == ' ' || *t == '\t')) ++t;
                        const char* e2 = t; double b = 0.0;
                        if (t < eol && !fast_number(e2, eol, b)) b = parse_double(t, eol, &e2);
                        fill->mz[p0 + np] = a; fill->inten[p0 + np] = (e2 != t) ? b : 0.0;
                        ++np;
                    }
                }
            } else if (c == 'B' && len >= 10 && !memcmp(q, "BEGIN IONS", 10)) {
                if (in_ions) np = np_begin;                    // a block without END IONS: its peaks belong to no spectrum, drop them (both passes)
                in_ions = true; pmz = 0.0; z = 0; np_begin = np;
                if (fill) fill->off[s0 + ns] = p0 + np;
            } else if (c == 'E' && len >= 8 && !memcmp(q, "END IONS", 8)) {
                if (in_ions) {
                    if (fill) { fill->prec[s0 + ns] = pmz; fill->charge[s0 + ns] = z; }
                    ++ns; in_ions = false;
                }
            } else if (in_ions && c == 'P' && len >= 8 && !memcmp(q, "PEPMASS=", 8)) {
                const char* e; pmz = parse_double(q + 8, eol, &e);                                   // first token = m/z (A11)
            } else if (in_ions && c == 'C' && len >= 7 && !memcmp(q, "CHARGE=", 7)) {
                z = (int32_t)strtol(q + 7, nullptr, 10);                                            // "2+" -> 2
                if (z < 0) z = -z;
            }
        }
        p = eol + 1;
    }
    if (in_ions) np = np_begin;                                // the text ends inside a block
    *n_spec = ns; *n_peak = np;
}

// first line at or after `from` whose first non-blank text is BEGIN IONS; `end` if none
const char* next_begin(const char* text, const char* from, const char* end) {
    const char* p = from;
    if (p > text) { const char* nl = (const char*)memchr(p - 1, '\n', (size_t)(end - (p - 1))); if (!nl) return end; p = nl + 1; }
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        const char* q = p;
        while (q < eol && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
        if ((size_t)(eol - q) >= 10 && !memcmp(q, "BEGIN IONS", 10)) return p;
        p = eol + 1;
    }
    return end;
}

}  // namespace

extern "C" int32_t pq_parse_mgf(const char* text, uint64_t n_bytes, int64_t* n_spectra, uint64_t* n_peaks, double* precursor_mz,
                                int32_t* charge, uint64_t* peak_offset, double* mz, double* intensity) {
    if (!text || !n_spectra || !n_peaks) return PQ_ERR_BAD_ARG;
    const bool fill = precursor_mz && charge && peak_offset && mz && intensity;
    const int64_t cap_s = *n_spectra; const uint64_t cap_p = *n_peaks;
    const char* end = text + n_bytes;
    int threads = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("PQ_MGF_THREADS")) threads = atoi(e);
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (n_bytes < (1u << 20)) threads = 1;
    // chunk boundaries at BEGIN IONS lines
    std::vector<const char*> cut; cut.push_back(text);
    for (int t = 1; t < threads; ++t) {
        const char* b = next_begin(text, text + n_bytes * (uint64_t)t / (uint64_t)threads, end);
        if (b > cut.back() && b < end) cut.push_back(b);
    }
    cut.push_back(end);
    const int nc = (int)cut.size() - 1;
    std::vector<int64_t> cs((size_t)nc, 0); std::vector<uint64_t> cp((size_t)nc, 0);
    auto run = [&](const Out* out, const std::vector<int64_t>& s0, const std::vector<uint64_t>& p0) {
        std::vector<std::thread> th;
        for (int c = 1; c < nc; ++c) th.emplace_back([&, c]() { parse_chunk(cut[(size_t)c], cut[(size_t)c + 1], out, s0[(size_t)c], p0[(size_t)c], &cs[(size_t)c], &cp[(size_t)c]); });
        parse_chunk(cut[0], cut[1], out, s0[0], p0[0], &cs[0], &cp[0]);
        for (auto& t : th) t.join();
    };
    std::vector<int64_t> s0((size_t)nc + 1, 0); std::vector<uint64_t> p0((size_t)nc + 1, 0);
    run(nullptr, s0, p0);                                  // count pass (cheap: no number parsing)
    for (int c = 0; c < nc; ++c) { s0[(size_t)c + 1] = s0[(size_t)c] + cs[(size_t)c]; p0[(size_t)c + 1] = p0[(size_t)c] + cp[(size_t)c]; }
    const int64_t ns = s0[(size_t)nc]; const uint64_t np = p0[(size_t)nc];
    if (fill) {
        if (ns > cap_s || np > cap_p) return PQ_ERR_CAPACITY;
        Out out{precursor_mz, charge, peak_offset, mz, intensity};
        run(&out, s0, p0);
        peak_offset[ns] = np;
    }
    *n_spectra = ns; *n_peaks = np;
    return PQ_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// On-disk MS/MS library.  index/BuildMSlibrary + IndexWorker write one MGF per 0.1-Da bin of the neutral
// precursor mass, <round(10 * mass)>.mgf (IndexWorker.java:422-486, BuildMSlibrary.getIntMass :349-352); msio/MsLibrarySearchWorker
// reads the bins a peptide's precursor window touches (:70-99, loadSpectra :154-252: spectra without a charge are ignored, the
// TITLE line is the spectrum's identity).  Here the reader takes a mass RANGE — the range a GPU's shard of a sweep covers — and
// returns every spectrum of the bins round(10 lo) .. round(10 hi) whose mass lies inside it, in bin order then file order, as the
// CSR arrays pq_load_spectra takes, plus the titles.
#include <cstdio>
#include <string>

namespace {

struct LibSpectra { std::vector<double> prec, mz, inten; std::vector<int32_t> charge; std::vector<uint64_t> off{0}; std::string titles; std::vector<uint64_t> toff{0}; };

bool read_file(const std::string& path, std::string& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    size_t got = n > 0 ? fread(&out[0], 1, (size_t)n, f) : 0;
    fclose(f);
    out.resize(got);
    return true;
}

void parse_library_file(const std::string& text, double lo, double hi, LibSpectra& L) {
    const char* p = text.data(); const char* end = p + text.size();
    bool in_ions = false; double pmz = 0.0; int32_t z = 0; std::string title;
    size_t peaks0 = L.mz.size();
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        const char* q = p;
        while (q < eol && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
        const char* r = eol; while (r > q && (r[-1] == '\r' || r[-1] == ' ' || r[-1] == '\t')) --r;
        const size_t len = (size_t)(r - q);
        if (len > 0) {
            const char c = *q;
            if (in_ions && ((c >= '0' && c <= '9') || c == '.')) {
                const char* e1 = q; double a;
                if (!fast_number(e1, eol, a)) a = parse_double(q, eol, &e1);
                if (e1 != q) {
                    const char* t = e1; while (t < eol && (*t == ' ' || *t == '\t')) ++t;
                    const char* e2 = t; double b = 0.0;
                    if (t < eol && !fast_number(e2, eol, b)) b = parse_double(t, eol, &e2);
                    L.mz.push_back(a); L.inten.push_back(e2 != t ? b : 0.0);
                }
            } else if (c == 'B' && len >= 10 && !memcmp(q, "BEGIN IONS", 10)) {
                if (in_ions) { L.mz.resize(peaks0); L.inten.resize(peaks0); }
                in_ions = true; pmz = 0.0; z = 0; title.clear(); peaks0 = L.mz.size();
            } else if (c == 'E' && len >= 8 && !memcmp(q, "END IONS", 8)) {
                if (in_ions) {
                    const double mass = pmz * z - z * pq::kProton;                       // Precursor.getMass
                    if (z >= 1 && mass >= lo && mass <= hi) {                           // MsLibrarySearchWorker.java:234-241
                        L.prec.push_back(pmz); L.charge.push_back(z); L.off.push_back(L.mz.size());
                        L.titles += title; L.toff.push_back(L.titles.size());
                    } else { L.mz.resize(peaks0); L.inten.resize(peaks0); }
                    in_ions = false;
                }
            } else if (in_ions && c == 'P' && len >= 8 && !memcmp(q, "PEPMASS=", 8)) {
                const char* e; pmz = parse_double(q + 8, eol, &e);
            } else if (in_ions && c == 'C' && len >= 7 && !memcmp(q, "CHARGE=", 7)) {
                z = (int32_t)strtol(q + 7, nullptr, 10); if (z < 0) z = -z;
            } else if (in_ions && c == 'T' && len >= 6 && !memcmp(q, "TITLE=", 6)) {
                title.assign(q + 6, r);
            }
        }
        p = eol + 1;
    }
    if (in_ions) { L.mz.resize(peaks0); L.inten.resize(peaks0); }
}

}  // namespace

extern "C" int32_t pq_read_ms_library(const char* dir, double mass_lo, double mass_hi, int64_t* n_spectra, uint64_t* n_peaks, uint64_t* title_bytes,
                                      double* precursor_mz, int32_t* charge, uint64_t* peak_offset, double* mz, double* intensity,
                                      uint64_t* title_offset, char* titles) {
    if (!dir || !n_spectra || !n_peaks || !title_bytes || !(mass_hi >= mass_lo) || !(mass_lo > -1e9) || mass_hi - mass_lo > 1e7) return PQ_ERR_BAD_ARG;
    const bool fill = precursor_mz && charge && peak_offset && mz && intensity && title_offset && titles;
    const long long b0 = pq::jround(mass_lo * 10.0), b1 = pq::jround(mass_hi * 10.0);
    const size_t nb = (size_t)(b1 - b0 + 1);
    std::vector<LibSpectra> part(nb);
    int threads = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("PQ_MGF_THREADS")) threads = atoi(e);
    threads = std::max(1, std::min(threads, 64));
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back([&, t]() {
        std::string text;
        for (size_t k = (size_t)t; k < nb; k += (size_t)threads) {
            const std::string path = std::string(dir) + "/" + std::to_string(b0 + (long long)k) + ".mgf";
            if (!read_file(path, text)) continue;                                    // a bin nobody wrote to: no file (MsLibrarySearchWorker.java:203-206)
            parse_library_file(text, mass_lo, mass_hi, part[k]);
        }
    });
    for (auto& t : th) t.join();
    int64_t ns = 0; uint64_t np = 0, tb = 0;
    for (auto& L : part) { ns += (int64_t)L.prec.size(); np += L.mz.size(); tb += L.titles.size(); }
    if (fill) {
        if (ns > *n_spectra || np > *n_peaks || tb > *title_bytes) return PQ_ERR_CAPACITY;
        int64_t s = 0; uint64_t pk = 0, t0 = 0;
        for (auto& L : part) {
            for (size_t i = 0; i < L.prec.size(); ++i) {
                precursor_mz[s] = L.prec[i]; charge[s] = L.charge[i]; peak_offset[s] = pk + L.off[i]; title_offset[s] = t0 + L.toff[i]; ++s;
            }
            if (!L.mz.empty()) { memcpy(mz + pk, L.mz.data(), L.mz.size() * 8); memcpy(intensity + pk, L.inten.data(), L.inten.size() * 8); }
            if (!L.titles.empty()) memcpy(titles + t0, L.titles.data(), L.titles.size());
            pk += L.mz.size(); t0 += L.titles.size();
        }
        peak_offset[ns] = np; title_offset[ns] = tb;
    }
    *n_spectra = ns; *n_peaks = np; *title_bytes = tb;
    return PQ_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// MGF text parsed ON THE DEVICE.  The text crosses PCIe in chunks; while later chunks are still on their way the
// line starts of the chunks that have arrived are selected (cub).  Then, over the whole text at once: one thread per line
// classifies it (peak line / BEGIN IONS / END IONS / PEPMASS= / CHARGE= / other), two prefix scans give every line its spectrum
// and every peak line its peak index, and one thread per line writes its numbers straight into the caller-order CSR arrays on
// the device.  Numbers take Clinger's exact fast path in FP64 — at most 18 digits collected in an integer below 2^53, one
// correctly rounded IEEE division by an exactly representable power of ten (__ddiv_rn) — which gives the same double as strtod
// / Double.parseDouble; the rare number of another shape (exponent, > 18 digits, "inf") is left to the host's strtod and
// patched in afterwards.  Blocks must be well formed (every BEGIN IONS closed by END IONS): anything else is refused and the
// caller takes the host parser.
#include <cub/cub.cuh>

#include "pq_host.h"

namespace pq {
namespace {

enum : uint8_t { kLineOther = 0, kLinePeak = 1, kLineBegin = 2, kLineEnd = 3, kLinePepmass = 4, kLineCharge = 5 };

struct GlobalLineStart {        // is the byte at global offset g the first of a line?  lo = first byte of the chunk (chunks begin at line starts)
    const char* text; uint64_t lo;
    __host__ __device__ bool operator()(uint64_t g) const { return g == lo || text[g - 1] == '\n'; }
};
struct IsNewline { const char* text; __host__ __device__ uint64_t operator()(uint64_t g) const { return text[g] == '\n' ? 1ull : 0ull; } };
struct AddBase { uint64_t base; __host__ __device__ uint64_t operator()(uint64_t i) const { return base + i; } };

__device__ __forceinline__ bool d_blank(char c) { return c == ' ' || c == '\t' || c == '\r'; }

__global__ void k_mgf_classify(const char* __restrict__ text, uint64_t n_bytes, const uint64_t* __restrict__ line_pos, uint64_t n_lines,
                               uint8_t* __restrict__ type, int2* __restrict__ delta /* x: +1 BEGIN, y: +1 BEGIN / -1 END */) {
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    uint64_t p = line_pos[l];
    const uint64_t end = l + 1 < n_lines ? line_pos[l + 1] : n_bytes;
    while (p < end && d_blank(text[p])) ++p;
    uint8_t t = kLineOther;
    if (p < end) {
        const char c = text[p];
        const uint64_t len = end - p;
        auto is = [&](const char* w, int n) { if (len < (uint64_t)n) return false; for (int i = 0; i < n; ++i) if (text[p + i] != w[i]) return false; return true; };
        if ((c >= '0' && c <= '9') || (c == '.' && len > 1 && text[p + 1] >= '0' && text[p + 1] <= '9')) t = kLinePeak;
        else if (c == 'B' && is("BEGIN IONS", 10)) t = kLineBegin;
        else if (c == 'E' && is("END IONS", 8)) t = kLineEnd;
        else if (c == 'P' && is("PEPMASS=", 8)) t = kLinePepmass;
        else if (c == 'C' && is("CHARGE=", 7)) t = kLineCharge;
    }
    type[l] = t;
    delta[l] = make_int2(t == kLineBegin ? 1 : 0, t == kLineBegin ? 1 : (t == kLineEnd ? -1 : 0));
}
struct AddInt2 { __host__ __device__ int2 operator()(const int2& a, const int2& b) const { return make_int2(a.x + b.x, a.y + b.y); } };

// peak lines inside a block; the block state must stay in {0, 1}
__global__ void k_mgf_peak_flags(const uint8_t* __restrict__ type, const int2* __restrict__ incl, uint64_t n_lines, uint8_t* __restrict__ pflag, uint32_t* __restrict__ bad) {
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    const int st = incl[l].y;
    if (st < 0 || st > 1) atomicAdd(bad, 1u);
    pflag[l] = (type[l] == kLinePeak && st == 1) ? 1 : 0;
}

// digits [. digits] with at most 18 digits, below 2^53, ended by a blank or the end of the line: value, next position.  false: another shape.
__device__ __forceinline__ bool d_fast_number(const char* __restrict__ text, uint64_t& p, uint64_t end, double& v) {
    const double pow10[19] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18};
    uint64_t q = p; unsigned long long m = 0; int nd = 0, frac = 0;
    while (q < end) { const unsigned d = (unsigned)(text[q] - '0'); if (d > 9u) break; m = m * 10ull + d; ++nd; ++q; }
    if (q < end && text[q] == '.') {
        ++q; const uint64_t f0 = q;
        while (q < end) { const unsigned d = (unsigned)(text[q] - '0'); if (d > 9u) break; m = m * 10ull + d; ++q; }
        frac = (int)(q - f0); nd += frac;
    }
    if (nd == 0 || nd > 18 || m >= (1ull << 53)) return false;
    if (q < end) { const char c = text[q]; if (c != ' ' && c != '\t' && c != '\r' && c != '\n') return false; }
    v = frac ? __ddiv_rn((double)m, pow10[frac]) : (double)m;
    p = q;
    return true;
}

struct HardNumber { uint64_t pos; uint64_t slot; uint32_t what; uint32_t pad; };      // what: 0 m/z, 1 intensity, 2 precursor m/z

__global__ void k_mgf_fill(const char* __restrict__ text, uint64_t n_bytes, const uint64_t* __restrict__ line_pos, uint64_t n_lines,
                           const uint8_t* __restrict__ type, const int2* __restrict__ incl, const uint8_t* __restrict__ pflag,
                           const uint64_t* __restrict__ pidx, double* __restrict__ prec, int32_t* __restrict__ charge, uint64_t* __restrict__ off,
                           double* __restrict__ mz, double* __restrict__ inten, HardNumber* __restrict__ hard, uint32_t* __restrict__ n_hard, uint32_t hard_cap) {
    uint64_t l = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_lines) return;
    const uint8_t t = type[l];
    if (t == kLineOther || t == kLineEnd) return;
    uint64_t p = line_pos[l];
    const uint64_t end = l + 1 < n_lines ? line_pos[l + 1] : n_bytes;
    while (p < end && d_blank(text[p])) ++p;
    const int64_t s = (int64_t)incl[l].x - 1;                       // spectrum of this line (index of the last BEGIN at or before it)
    auto defer = [&](uint64_t pos, uint64_t slot, uint32_t what) { uint32_t k = atomicAdd(n_hard, 1u); if (k < hard_cap) hard[k] = HardNumber{pos, slot, what, 0u}; };
    if (t == kLineBegin) { off[s] = pidx[l]; return; }
    if (incl[l].y != 1 || s < 0) return;                            // outside a block
    if (t == kLinePepmass) { uint64_t q = p + 8; while (q < end && (text[q] == ' ' || text[q] == '\t')) ++q; double v; if (d_fast_number(text, q, end, v)) prec[s] = v; else defer(p + 8, (uint64_t)s, 2u); return; }
    if (t == kLineCharge) {
        uint64_t q = p + 7; while (q < end && (text[q] == ' ' || text[q] == '\t')) ++q;
        bool neg = false; if (q < end && (text[q] == '+' || text[q] == '-')) { neg = text[q] == '-'; ++q; }
        int z = 0; while (q < end && (unsigned)(text[q] - '0') <= 9u && z < 100000) { z = z * 10 + (text[q] - '0'); ++q; }
        (void)neg; charge[s] = z;                                   // "2+" -> 2, "2-" -> 2 (pq_parse_mgf takes the magnitude too)
        return;
    }
    if (!pflag[l]) return;
    const uint64_t k = pidx[l];
    double a = 0.0, b = 0.0;
    uint64_t q = p;
    if (!d_fast_number(text, q, end, a)) { defer(p, k, 0u); defer(p, k, 1u); return; }      // the host parses the whole line
    while (q < end && (text[q] == ' ' || text[q] == '\t')) ++q;
    if (q < end && text[q] != '\r' && text[q] != '\n') { uint64_t q2 = q; if (!d_fast_number(text, q2, end, b)) { b = 0.0; defer(p, k, 1u); } }
    mz[k] = a; inten[k] = b;
}

#define RB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return false; } } while (0)

}  // namespace

bool parse_mgf_on_device(cudaStream_t st, cudaStream_t sc, const char* text, uint64_t n_bytes, MgfDeviceWork& W, DevBuf& d_prec, DevBuf& d_charge,
                         DevBuf& d_off, DevBuf& d_mz, DevBuf& d_in, int64_t* n_spectra, uint64_t* n_peaks, uint64_t* launches, std::string& err) {
    *n_spectra = 0; *n_peaks = 0;
    if (n_bytes == 0) return true;
    RB(W.text.reserve(n_bytes + 16));
    char* d_text = W.text.as<char>();
    // upper bound of the line count: a line is at least one byte; in practice ~16 bytes — reserve for n_bytes / 2 + 1 lines (a text of
    // empty lines only would need more and is refused)
    const uint64_t line_cap = n_bytes / 2 + 16;
    RB(W.line_pos.reserve(line_cap * 8)); RB(W.scal.reserve(256));
    RB(cudaMemsetAsync(W.scal.p, 0, 256, st));
    const uint64_t chunk = 64ull << 20;
    std::vector<uint64_t> cuts; cuts.push_back(0);
    while (cuts.back() < n_bytes) {                                  // chunks end at a line end
        uint64_t e = std::min(n_bytes, cuts.back() + chunk);
        if (e < n_bytes) { const char* nl = (const char*)memchr(text + e, '\n', (size_t)(n_bytes - e)); e = nl ? (uint64_t)(nl - text) + 1 : n_bytes; }
        cuts.push_back(e);
    }
    const size_t nc = cuts.size() - 1;
    while (W.ev.size() < nc) { cudaEvent_t e; RB(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); W.ev.push_back(e); }
    for (size_t c = 0; c < nc; ++c) {
        RB(cudaMemcpyAsync(d_text + cuts[c], text + cuts[c], cuts[c + 1] - cuts[c], cudaMemcpyHostToDevice, sc));
        RB(cudaEventRecord(W.ev[c], sc));
    }
    uint64_t n_lines = 0;
    uint64_t* d_num = reinterpret_cast<uint64_t*>(W.scal.as<uint8_t>() + 64);
    for (size_t c = 0; c < nc; ++c) {
        RB(cudaStreamWaitEvent(st, W.ev[c], 0));
        const uint64_t len = cuts[c + 1] - cuts[c];
        cub::CountingInputIterator<uint64_t> iota(0);
        cub::TransformInputIterator<uint64_t, AddBase, cub::CountingInputIterator<uint64_t>> pos_it(iota, AddBase{cuts[c]});
        {   // the chunk's line count first: the select below must not write past the reserved positions (a text of very short lines)
            cub::TransformInputIterator<uint64_t, IsNewline, decltype(pos_it)> nl_it(pos_it, IsNewline{d_text});
            size_t bytes = 0;
            cub::DeviceReduce::Sum(nullptr, bytes, nl_it, d_num, (int64_t)len, st);
            RB(W.tmp.reserve(bytes + 64));
            RB(cub::DeviceReduce::Sum(W.tmp.p, bytes, nl_it, d_num, (int64_t)len, st));
            uint64_t nl = 0;
            RB(cudaMemcpyAsync(&nl, d_num, 8, cudaMemcpyDeviceToHost, st));
            RB(cudaStreamSynchronize(st));
            if (n_lines + nl + 1 > line_cap) { err = "too many lines for the reserved work space (lines shorter than two bytes on average)"; return false; }
        }
        // select the global offsets of the line starts of this chunk
        size_t bytes = 0;
        GlobalLineStart gs{d_text, cuts[c]};
        cub::DeviceSelect::If(nullptr, bytes, pos_it, W.line_pos.as<uint64_t>() + n_lines, d_num, (int64_t)len, gs, st);
        RB(W.tmp.reserve(bytes + 64));
        RB(cub::DeviceSelect::If(W.tmp.p, bytes, pos_it, W.line_pos.as<uint64_t>() + n_lines, d_num, (int64_t)len, gs, st));
        uint64_t got = 0;
        RB(cudaMemcpyAsync(&got, d_num, 8, cudaMemcpyDeviceToHost, st));
        RB(cudaStreamSynchronize(st));
        n_lines += got;
        *launches += 3;
        if (n_lines > line_cap) { err = "too many lines for the reserved work space"; return false; }
    }
    if (n_lines == 0) return true;
    // classify, scan, count
    RB(W.type.reserve(n_lines + 64)); RB(W.delta.reserve(n_lines * 8)); RB(W.incl.reserve(n_lines * 8)); RB(W.pflag.reserve(n_lines + 64)); RB(W.pidx.reserve((n_lines + 1) * 8));
    const unsigned nb = (unsigned)((n_lines + 255) / 256);
    k_mgf_classify<<<nb, 256, 0, st>>>(d_text, n_bytes, W.line_pos.as<uint64_t>(), n_lines, W.type.as<uint8_t>(), W.delta.as<int2>());
    {
        size_t bytes = 0;
        cub::DeviceScan::InclusiveScan(nullptr, bytes, W.delta.as<int2>(), W.incl.as<int2>(), AddInt2(), (int64_t)n_lines, st);
        RB(W.tmp.reserve(bytes + 64));
        RB(cub::DeviceScan::InclusiveScan(W.tmp.p, bytes, W.delta.as<int2>(), W.incl.as<int2>(), AddInt2(), (int64_t)n_lines, st));
    }
    uint32_t* d_bad = W.scal.as<uint32_t>();
    k_mgf_peak_flags<<<nb, 256, 0, st>>>(W.type.as<uint8_t>(), W.incl.as<int2>(), n_lines, W.pflag.as<uint8_t>(), d_bad);
    {
        size_t bytes = 0;
        cub::TransformInputIterator<uint64_t, ByteToU64, const uint8_t*> pf(W.pflag.as<uint8_t>(), ByteToU64());
        cub::DeviceScan::ExclusiveSum(nullptr, bytes, pf, W.pidx.as<uint64_t>(), (int64_t)n_lines + 1, st);
        RB(W.tmp.reserve(bytes + 64));
        RB(cub::DeviceScan::ExclusiveSum(W.tmp.p, bytes, pf, W.pidx.as<uint64_t>(), (int64_t)n_lines + 1, st));
    }
    struct { uint32_t bad; } hb = {0}; int2 last = make_int2(0, 0); uint64_t np = 0;
    RB(cudaMemcpyAsync(&hb, d_bad, 4, cudaMemcpyDeviceToHost, st));
    RB(cudaMemcpyAsync(&last, W.incl.as<int2>() + (n_lines - 1), 8, cudaMemcpyDeviceToHost, st));
    RB(cudaMemcpyAsync(&np, W.pidx.as<uint64_t>() + n_lines, 8, cudaMemcpyDeviceToHost, st));
    RB(cudaStreamSynchronize(st));
    *launches += 6;
    if (hb.bad || last.y != 0) { err = "malformed MGF blocks (a BEGIN IONS without its END IONS, or the reverse)"; return false; }
    const int64_t ns = last.x;
    *n_spectra = ns; *n_peaks = np;
    if (ns == 0) return true;
    RB(d_prec.reserve((size_t)ns * 8)); RB(d_charge.reserve((size_t)ns * 4)); RB(d_off.reserve(((size_t)ns + 1) * 8));
    RB(d_mz.reserve(std::max<uint64_t>(np, 1) * 8)); RB(d_in.reserve(std::max<uint64_t>(np, 1) * 8));
    RB(cudaMemsetAsync(d_prec.p, 0, (size_t)ns * 8, st)); RB(cudaMemsetAsync(d_charge.p, 0, (size_t)ns * 4, st));
    const uint32_t hard_cap = 1u << 20;
    RB(W.hard.reserve((size_t)hard_cap * sizeof(HardNumber)));
    uint32_t* d_nhard = W.scal.as<uint32_t>() + 4;
    k_mgf_fill<<<nb, 256, 0, st>>>(d_text, n_bytes, W.line_pos.as<uint64_t>(), n_lines, W.type.as<uint8_t>(), W.incl.as<int2>(), W.pflag.as<uint8_t>(), W.pidx.as<uint64_t>(),
                                   d_prec.as<double>(), d_charge.as<int32_t>(), d_off.as<uint64_t>(), d_mz.as<double>(), d_in.as<double>(),
                                   W.hard.as<HardNumber>(), d_nhard, hard_cap);
    RB(cudaMemcpyAsync(d_off.as<uint64_t>() + ns, &np, 8, cudaMemcpyHostToDevice, st));
    uint32_t n_hard = 0;
    RB(cudaMemcpyAsync(&n_hard, d_nhard, 4, cudaMemcpyDeviceToHost, st));
    RB(cudaStreamSynchronize(st));
    RB(cudaGetLastError());
    *launches += 2;
    if (n_hard > hard_cap) { err = "too many numbers outside the fast path for the device parser"; return false; }
    if (n_hard) {                                                    // the host's strtod decides the rare number of another shape
        std::vector<HardNumber> h(n_hard);
        RB(cudaMemcpyAsync(h.data(), W.hard.p, (size_t)n_hard * sizeof(HardNumber), cudaMemcpyDeviceToHost, st));
        RB(cudaStreamSynchronize(st));
        const char* end = text + n_bytes;
        for (const HardNumber& x : h) {
            const char* p = text + x.pos; const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p)); if (!eol) eol = end;
            const char* e1 = p; double a = ::parse_double(p, eol, &e1), v = a;
            if (x.what == 1) { const char* t = e1; while (t < eol && (*t == ' ' || *t == '\t')) ++t; const char* e2 = t; double b = t < eol ? ::parse_double(t, eol, &e2) : 0.0; v = (e2 != t) ? b : 0.0; }
            double* dst = x.what == 2 ? d_prec.as<double>() + x.slot : (x.what == 0 ? d_mz.as<double>() + x.slot : d_in.as<double>() + x.slot);
            RB(cudaMemcpyAsync(dst, &v, 8, cudaMemcpyHostToDevice, st));
            RB(cudaStreamSynchronize(st));
        }
    }
    return true;
}

}  // namespace pq

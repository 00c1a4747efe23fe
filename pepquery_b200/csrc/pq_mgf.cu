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

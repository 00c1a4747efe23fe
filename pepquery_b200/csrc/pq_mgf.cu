// pq_mgf.cu — single-pass MGF text parser into CSR buffers.  Replaces compomics MgfFileIterator as
// used by SpectraInput.readMSMSfast (pg/SpectraInput.java:150-195); the writer-side format is index/IndexWorker.java:431-457
// (BEGIN IONS / TITLE / PEPMASS / CHARGE / "mz intensity" lines / END IONS).  Host code, no device work.
#include <cstdlib>
#include <cstring>

#include "pq_common.h"

extern "C" int32_t pq_parse_mgf(const char* text, uint64_t n_bytes, int64_t* n_spectra, uint64_t* n_peaks, double* precursor_mz,
                                int32_t* charge, uint64_t* peak_offset, double* mz, double* intensity) {
    if (!text || !n_spectra || !n_peaks) return PQ_ERR_BAD_ARG;
    const bool fill = precursor_mz && charge && peak_offset && mz && intensity;
    const int64_t cap_s = *n_spectra; const uint64_t cap_p = *n_peaks;
    int64_t ns = 0; uint64_t np = 0;
    bool in_ions = false; double pmz = 0.0; int32_t z = 0;
    const char* p = text; const char* end = text + n_bytes;
    while (p < end) {
        const char* eol = (const char*)memchr(p, '\n', (size_t)(end - p));
        if (!eol) eol = end;
        const char* q = p;
        while (q < eol && (*q == ' ' || *q == '\t' || *q == '\r')) ++q;
        size_t len = (size_t)(eol - q);
        if (len >= 10 && !memcmp(q, "BEGIN IONS", 10)) {
            in_ions = true; pmz = 0.0; z = 0;
            if (fill) { if (ns >= cap_s) return PQ_ERR_CAPACITY; peak_offset[ns] = np; }
        } else if (len >= 8 && !memcmp(q, "END IONS", 8)) {
            if (in_ions) {
                if (fill) { precursor_mz[ns] = pmz; charge[ns] = z; }
                ++ns; in_ions = false;
            }
        } else if (in_ions && len > 0) {
            if ((*q >= '0' && *q <= '9') || *q == '.') {
                char* e1 = nullptr; double a = strtod(q, &e1);
                char* e2 = nullptr; double b = (e1 && e1 < eol) ? strtod(e1, &e2) : 0.0;
                if (e1 != q) {
                    if (fill) { if (np >= cap_p) return PQ_ERR_CAPACITY; mz[np] = a; intensity[np] = (e2 && e2 != e1) ? b : 0.0; }
                    ++np;
                }
            } else if (len >= 8 && !memcmp(q, "PEPMASS=", 8)) {
                pmz = strtod(q + 8, nullptr);                       // first token = m/z (A11)
            } else if (len >= 7 && !memcmp(q, "CHARGE=", 7)) {
                z = (int32_t)strtol(q + 7, nullptr, 10);            // "2+" -> 2
                if (z < 0) z = -z;
            }
        }
        p = eol + 1;
    }
    if (fill) peak_offset[ns] = np;
    *n_spectra = ns; *n_peaks = np;
    return PQ_OK;
}

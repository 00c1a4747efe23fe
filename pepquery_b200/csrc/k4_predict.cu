// k4_predict.cu — MVH "total predicted peaks" per peptide row (peptide-only, so computed once per row, not per pair).
//
// MVHMatch.calcMVH (PSMMatch/MVHMatch.java:150-227): every b/y ion the ion factory emits with fewer than two neutral
// losses and no H2O/NH3 loss — plain ions plus one single-loss variant per modification-defined loss carried by the
// peptide (assumption A8) — at fragment charges {1,2} ({1} when the precursor charge is 2), kept when
// 200 <= m/z <= 2000, made unique, then removeFuzzyValues(set, itol) (:104-119) counts the values whose successor is
// at least itol away, plus the last one.
//
// Each (ion type, loss, charge) series is ascending in the ion number, so the sorted unique stream is an S-way merge of
// at most 2*3*2 series; one thread per row, two results per row: [2*row] for precursor charge 2, [2*row+1] otherwise.
#include "k4_device.cuh"

namespace pq {

namespace {

__global__ void k4_predict_kernel(const uint8_t* __restrict__ rows, uint32_t n_rows, const ModTable* __restrict__ M, double itol,
                                  int32_t* __restrict__ out) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const uint8_t* row = rows + (size_t)r * kRowBytes;
    dev::Extra ex; ex.site = -1; ex.delta = 0.0;
    double b[kMaxLen], y[kMaxLen], loss[dev::kMaxLoss]; int nloss;
    dev::predict_ladders(row, M, ex, b, y, loss, nloss);                    // one ladder walk for both charge sets
    out[2 * (size_t)r] = dev::predict_count(b, y, (int)row[0], loss, nloss, itol, 1);          // precursor charge 2: fragment charge 1 only
    out[2 * (size_t)r + 1] = dev::predict_count(b, y, (int)row[0], loss, nloss, itol, 2);
}

}  // namespace

int launch_predict(const uint8_t* rows, uint32_t n_rows, const ModTable* mods, double itol, int32_t* out, cudaStream_t st) {
    if (n_rows == 0) return 0;
    k4_predict_kernel<<<(n_rows + 127) / 128, 128, 0, st>>>(rows, n_rows, mods, itol, out);
    return 1;
}

}  // namespace pq

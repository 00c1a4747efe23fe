// k_cells.cuh — per-row "cell lists": the spectrum-independent half of the Step-4 count pass, hoisted out of K1.
//
// The score bound of Step 4 (k1_device.cuh, below_score_bound) only needs, per shuffled peptide, how many of its theoretical
// fragment m/z land in a cell of the spectrum's m/z -> cell map where something can be claimed.  The m/z themselves — the
// fragment ladder of PeptideSpectrumAnnotator (JMatch.java:388-466), walked in the reference's order with explicit
// round-to-nearest adds — depend on the peptide row and the fragment charge only, and a shuffle row is scored against ~25
// spectra.  So the kernel that WRITES a row (K2 for shuffles) also writes, once, the cell of every (ion, fragment charge):
// K1's count pass then is "load entry, test one flag of the staged spectrum, add".
//
// One row's lists = 2 * chunks_side 16-byte units of eight u16 entries:
//   side 0 (b ions) | side 1 (y ions); inside a side: charge 1 ions k = 1..L-1, charge 2 ions k = 2..L-1, charge 3 ions
//   k = 3..L-1 (A6: fragment charge c needs c <= ion number), concatenated, the tail of the last unit padded.
//   entry = byte offset of the cell in BlobLayout::cells = 2 * cell_filter(m/z).
// A job with nch = precursor charge - 1 fragment charges reads the first T(nch) = sum_{c<=nch} max(L - c, 0) entries of each side.
// The rows of one peptide form are stored unit-major: unit u of row k at base + u * stride + k (stride = rows reserved for the
// form), so the 32 rows a warp of K1 (one thread per row) reads, and a warp of K2 writes, are 512 contiguous bytes per unit.
#pragma once
#include "pq_common.h"

namespace pq {

constexpr int kListMaxCharge = 3;       // fragment charges held in the lists (precursor charge <= 4); above: K1 walks the ladder

PQ_HD uint32_t list_entries(int L, int nch) {      // entries of one side for nch fragment charges
    uint32_t t = 0;
    for (int c = 1; c <= nch; ++c) t += (uint32_t)(L - c > 0 ? L - c : 0);
    return t;
}
PQ_HD uint32_t list_chunks_side(int L, int nch_max) { return (list_entries(L, nch_max) + 7u) >> 3; }

#ifdef __CUDACC__
// Writes both sides of one row's lists.  `row` = the 64 row bytes (any address space), M = the search's residue table.
__device__ __forceinline__ void write_cell_lists(const uint8_t* row, const ModTable* __restrict__ M, int nch_max, uint4* __restrict__ dst /* unit 0 of this row */,
                                                 uint32_t chunks_side, size_t stride) {
    const int L = (int)row[0];
    for (int side = 0; side < 2; ++side) {
        unsigned long long lo = 0ull, hi = 0ull; int fill = 0;            // eight u16 entries = one 16-byte unit
        uint4* out = dst + (size_t)side * chunks_side * stride;
        for (int c = 1; c <= nch_max; ++c) {
            double acc = side == 0 ? M->nterm[row[1]] : M->cterm[row[2]];
            for (int k = 1; k <= L - 1; ++k) {
                const uint32_t code = row[kRowHeader + (side == 0 ? k - 1 : L - k)];
                acc = __dadd_rn(__dadd_rn(acc, M->aa[code]), M->delta[code]);
                if (k < c) continue;
                const double m = side == 0 ? acc : __dadd_rn(acc, kWater);
                double t;
                if (c == 1) t = __dadd_rn(m, kProton);
                else if (c == 2) t = __dmul_rn(__dadd_rn(m, 2.0 * kProton), 0.5);
                else t = __ddiv_rn(__dadd_rn(m, __dmul_rn(3.0, kProton)), 3.0);
                const uint32_t e = 2u * cell_filter(t);
                if (fill < 4) lo |= (unsigned long long)e << (16 * fill); else hi |= (unsigned long long)e << (16 * (fill - 4));
                if (++fill == 8) { *out = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32)); out += stride; lo = hi = 0ull; fill = 0; }
            }
        }
        if (fill) *out = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));      // padding entries are offset 0: readable, never counted
    }
}
#endif

}  // namespace pq

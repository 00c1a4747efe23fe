// k_cells.cuh — per-row "cell lists": the spectrum-independent half of the Step-4 count pass, hoisted out of K1.
//
// The score bound of Step 4 (k1_device.cuh, below_score_bound) only needs, per shuffled peptide, how many of its theoretical
// fragment m/z land in a cell of the spectrum's m/z -> cell map where something can be claimed.  The m/z themselves — the
// fragment ladder of PeptideSpectrumAnnotator (JMatch.java:388-466), walked in the reference's order with explicit
// round-to-nearest adds — depend on the peptide row and the fragment charge only, and a shuffle row is scored against ~25
// spectra.  So the kernel that WRITES a row (K2 for shuffles) also writes, once, the cell of every (ion, fragment charge):
// K1's count pass then is "load entry, test one flag of the staged spectrum, add".
//
// One row's lists = 2 sides x nch_max charge streams x U 16-byte units of eight u16 entries, U = ceil((L - 1) / 8):
//   side 0 (b ions) | side 1 (y ions); inside a side: stream of charge 1, of charge 2, of charge 3; slot k - 1 of a stream = ion k,
//   entry = the ion's cell, cell_filter(m/z).  K1 turns the staged spectrum's clean flags into a byte per cell (1 = nothing can be
//   claimed) once per job, so an entry costs one byte load and an add.  Slots that are no ion (k < charge: A6, fragment charge c
//   needs c <= ion number; the padding of the last unit) hold kListSentinel, an always-clean extra cell behind that byte table, so a
//   reader counts whole units and needs no length.
// A job with nch = precursor charge - 1 fragment charges reads the first nch streams of each side.
// The rows of one peptide form are stored unit-major: unit u of row k at base + u * stride + k (stride = rows reserved for the
// form), so the 32 rows a warp of K1 (one thread per row) reads, and a warp of K2 writes, are 512 contiguous bytes per unit.
#pragma once
#include <type_traits>

#include "pq_common.h"

namespace pq {

constexpr int kListMaxCharge = 3;       // fragment charges held in the lists (precursor charge <= 4); above: K1 walks the ladder
constexpr uint32_t kListSentinel = kCells;      // entry of a slot that is no ion: the always-clean cell behind K1's byte table

PQ_HD uint32_t list_entries(int L, int nch) {      // ions x charges of one side for nch fragment charges
    uint32_t t = 0;
    for (int c = 1; c <= nch; ++c) t += (uint32_t)(L - c > 0 ? L - c : 0);
    return t;
}
PQ_HD uint32_t list_stream_units(int L) { return L > 1 ? (uint32_t)(L - 1 + 7) >> 3 : 0u; }            // U
PQ_HD uint32_t list_chunks_side(int L, int nch_max) { return (uint32_t)nch_max * list_stream_units(L); }

#ifdef __CUDACC__
// entry e into slot J (a compile-time number, 0..7) of a unit held as four 32-bit words
template <int J> __device__ __forceinline__ void unit_put(uint32_t (&w)[4], uint32_t e) {
    if (J & 1) w[J >> 1] = (w[J >> 1] & 0x0000FFFFu) | (e << 16);
    else w[J >> 1] = (w[J >> 1] & 0xFFFF0000u) | e;
}

// Writes both sides of one row's lists: one ladder walk per side, every charge of an ion from the same running sum, a unit (eight
// ions) at a time so that every slot of the packed words is a compile-time position.
// `row` = the 64 row bytes, tab[code] = {residue mass, modification mass}, nt / ct = terminal masses (any address space).
__device__ __forceinline__ void write_cell_lists(const uint8_t* row, const double2* tab, const double* nt, const double* ct, int nch_max,
                                                 uint4* __restrict__ dst /* unit 0 of this row */, uint32_t chunks_side, size_t stride) {
    const int L = (int)row[0], Lm1 = L - 1;
    const uint32_t U = list_stream_units(L);
    constexpr uint32_t kSent2 = 0x00010001u * kListSentinel;
    for (int side = 0; side < 2; ++side) {
        uint4* out = dst + (size_t)side * chunks_side * stride;
        double acc = side == 0 ? nt[row[1]] : ct[row[2]];
        for (uint32_t u = 0; u < U; ++u) {
            uint32_t w[kListMaxCharge][4];
#pragma unroll
            for (int c = 0; c < kListMaxCharge; ++c) { w[c][0] = kSent2; w[c][1] = kSent2; w[c][2] = kSent2; w[c][3] = kSent2; }
            auto ion = [&](int k, auto slot) {
                constexpr int J = decltype(slot)::value;
                const double2 t = tab[row[kRowHeader + (side == 0 ? k - 1 : L - k)]];
                acc = __dadd_rn(__dadd_rn(acc, t.x), t.y);
                const double m = side == 0 ? acc : __dadd_rn(acc, kWater);
                unit_put<J>(w[0], cell_filter(__dadd_rn(m, kProton)));
                if (nch_max >= 2 && k >= 2) unit_put<J>(w[1], cell_filter(__dmul_rn(__dadd_rn(m, 2.0 * kProton), 0.5)));
                if (nch_max >= 3 && k >= 3) unit_put<J>(w[2], cell_filter(div3_rn(__dadd_rn(m, __dmul_rn(3.0, kProton)))));      // skipped when no PSM of the form needs charge 3
            };
            const int k0 = 8 * (int)u + 1;
            if (k0 + 0 <= Lm1) ion(k0 + 0, std::integral_constant<int, 0>());
            if (k0 + 1 <= Lm1) ion(k0 + 1, std::integral_constant<int, 1>());
            if (k0 + 2 <= Lm1) ion(k0 + 2, std::integral_constant<int, 2>());
            if (k0 + 3 <= Lm1) ion(k0 + 3, std::integral_constant<int, 3>());
            if (k0 + 4 <= Lm1) ion(k0 + 4, std::integral_constant<int, 4>());
            if (k0 + 5 <= Lm1) ion(k0 + 5, std::integral_constant<int, 5>());
            if (k0 + 6 <= Lm1) ion(k0 + 6, std::integral_constant<int, 6>());
            if (k0 + 7 <= Lm1) ion(k0 + 7, std::integral_constant<int, 7>());
#pragma unroll
            for (int c = 0; c < kListMaxCharge; ++c)
                if (c < nch_max) out[((size_t)c * U + u) * stride] = make_uint4(w[c][0], w[c][1], w[c][2], w[c][3]);
        }
    }
}
#endif

}  // namespace pq

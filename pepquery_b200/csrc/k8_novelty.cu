// k8_novelty.cu — Step-1 novelty filter on the device: is a target sequence present in the reference proteome?
// Replaces InputProcessor.searchRefDB (pg/InputProcessor.java:379-448) with its default mapping, the FM index of
// pg/PepMapping.java:60-107: hasProtein(seq) = "seq occurs in some protein, I and L indistinguishable" (compomics
// SequenceMatchingParameters.MatchingType.indistiguishableAminoAcids).  Exact substring semantics, not digest membership.
//
// One thread per residue position of the concatenated proteome: the K-mer starting there (5 bits per letter, case folded,
// I -> L) is looked up in an open-addressing table of the targets' leading K-mers; every entry with that key is verified
// residue by residue inside the protein that owns the position.  The proteome is read once (HBM bound, ~9 MB for the 20 k
// protein database); the table (a few thousand targets) stays in L1/L2.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "pq_internal.h"

namespace pq {

namespace {

__host__ __device__ __forceinline__ uint32_t fold_letter(uint8_t c) {     // 0..25 for letters (I folded into L), 31 otherwise
    uint32_t u = (uint32_t)(c & 0xDFu) - 65u;                               // upper-case, 'A' -> 0
    if (u > 25u || ((c | 0x20u) < 97u)) return 31u;
    return u == 8u ? 11u : u;                                               // I -> L
}

__global__ void k8_scan(const uint8_t* __restrict__ prot, uint64_t n_res, const uint64_t* __restrict__ prot_off, int32_t n_prot,
                        const uint32_t* __restrict__ table, uint32_t table_mask, int K, const uint32_t* __restrict__ t_off,
                        const uint8_t* __restrict__ t_res, const uint32_t* __restrict__ t_key, int32_t* __restrict__ nhit) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p + (uint64_t)K > n_res) return;
    uint32_t key = 0;
    for (int i = 0; i < K; ++i) key = (key << 5) | fold_letter(prot[p + i]);
    uint32_t slot = (key * 2654435761u) & table_mask;
    int64_t end = -1;                                                       // end of the protein that owns p (found on first use)
    for (;; slot = (slot + 1u) & table_mask) {
        const uint32_t t = table[slot];
        if (t == 0xFFFFFFFFu) return;                                       // empty slot ends the probe sequence
        if (t_key[t] != key) continue;
        if (end < 0) {
            int lo = 0, hi = n_prot;                                        // last protein whose offset <= p
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (prot_off[mid] <= p) lo = mid; else hi = mid; }
            end = (int64_t)prot_off[lo + 1];
        }
        const uint32_t a = t_off[t], L = t_off[t + 1] - a;
        if ((int64_t)(p + L) > end) continue;
        bool same = true;
        for (uint32_t i = (uint32_t)K; i < L && same; ++i) same = fold_letter(prot[p + i]) == (uint32_t)t_res[a + i];
        if (same) nhit[t] = 1;                                              // benign race: every writer stores 1
    }
}

}  // namespace

// Host side: folded targets, the K-mer table, launch.  err is set on refusal.
bool find_in_proteome(int32_t n_seq, const uint32_t* seq_off, const char* res, int32_t n_prot, const uint64_t* prot_off, const char* prot_res,
                      int32_t* nhit, cudaStream_t st, uint64_t* h2d, uint64_t* d2h, int* launches, std::string& err) {
    for (int32_t i = 0; i < n_seq; ++i) nhit[i] = 0;
    if (n_seq == 0 || n_prot == 0) return true;
    const uint64_t n_res = prot_off[n_prot] - prot_off[0];
    uint32_t min_len = 0xFFFFFFFFu;
    for (int32_t i = 0; i < n_seq; ++i) {
        if (seq_off[i + 1] <= seq_off[i]) { err = "empty target sequence"; return false; }
        min_len = std::min(min_len, seq_off[i + 1] - seq_off[i]);
    }
    const int K = (int)std::min<uint32_t>(min_len, 6u);
    const uint32_t n_letters = seq_off[n_seq] - seq_off[0];
    std::vector<uint8_t> folded(n_letters); std::vector<uint32_t> off((size_t)n_seq + 1), key((size_t)n_seq);
    for (int32_t i = 0; i <= n_seq; ++i) off[(size_t)i] = seq_off[i] - seq_off[0];
    for (uint32_t i = 0; i < n_letters; ++i) {
        folded[i] = (uint8_t)fold_letter((uint8_t)res[seq_off[0] + i]);
        if (folded[i] == 31u) { err = "target sequences must be letters A-Z"; return false; }
    }
    uint32_t size = 16; while (size < 2u * (uint32_t)n_seq) size <<= 1;
    std::vector<uint32_t> table(size, 0xFFFFFFFFu);
    for (int32_t t = 0; t < n_seq; ++t) {
        uint32_t k = 0;
        for (int i = 0; i < K; ++i) k = (k << 5) | folded[off[(size_t)t] + (uint32_t)i];
        key[(size_t)t] = k;
        uint32_t slot = (k * 2654435761u) & (size - 1);
        while (table[slot] != 0xFFFFFFFFu) slot = (slot + 1u) & (size - 1);
        table[slot] = (uint32_t)t;
    }
    // proteins: offsets rebased to 0
    std::vector<uint64_t> poff((size_t)n_prot + 1);
    for (int32_t i = 0; i <= n_prot; ++i) poff[(size_t)i] = prot_off[i] - prot_off[0];
    DevBuf d_prot, d_poff, d_table, d_toff, d_tres, d_tkey, d_hit;
    auto ok = [&](cudaError_t e, const char* what) { if (e != cudaSuccess) { err = std::string(what) + ": " + cudaGetErrorString(e); return false; } return true; };
    if (!ok(d_prot.reserve(n_res + 16), "alloc") || !ok(d_poff.reserve(poff.size() * 8), "alloc") || !ok(d_table.reserve(table.size() * 4), "alloc") ||
        !ok(d_toff.reserve(off.size() * 4), "alloc") || !ok(d_tres.reserve(folded.size() + 16), "alloc") || !ok(d_tkey.reserve(key.size() * 4), "alloc") ||
        !ok(d_hit.reserve((size_t)n_seq * 4), "alloc")) return false;
    if (!ok(cudaMemcpyAsync(d_prot.p, prot_res + prot_off[0], n_res, cudaMemcpyHostToDevice, st), "copy proteins") ||
        !ok(cudaMemcpyAsync(d_poff.p, poff.data(), poff.size() * 8, cudaMemcpyHostToDevice, st), "copy") ||
        !ok(cudaMemcpyAsync(d_table.p, table.data(), table.size() * 4, cudaMemcpyHostToDevice, st), "copy") ||
        !ok(cudaMemcpyAsync(d_toff.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice, st), "copy") ||
        !ok(cudaMemcpyAsync(d_tres.p, folded.data(), folded.size(), cudaMemcpyHostToDevice, st), "copy") ||
        !ok(cudaMemcpyAsync(d_tkey.p, key.data(), key.size() * 4, cudaMemcpyHostToDevice, st), "copy") ||
        !ok(cudaMemsetAsync(d_hit.p, 0, (size_t)n_seq * 4, st), "memset")) return false;
    *h2d += n_res + poff.size() * 8 + table.size() * 4 + off.size() * 4 + folded.size() + key.size() * 4;
    if (n_res >= (uint64_t)K) {
        const uint64_t threads = n_res - (uint64_t)K + 1;
        k8_scan<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d_prot.as<uint8_t>(), n_res, d_poff.as<uint64_t>(), n_prot, d_table.as<uint32_t>(), size - 1, K,
                                                                   d_toff.as<uint32_t>(), d_tres.as<uint8_t>(), d_tkey.as<uint32_t>(), d_hit.as<int32_t>());
        ++*launches;
    }
    if (!ok(cudaGetLastError(), "k8_scan") || !ok(cudaMemcpyAsync(nhit, d_hit.p, (size_t)n_seq * 4, cudaMemcpyDeviceToHost, st), "copy back") ||
        !ok(cudaStreamSynchronize(st), "sync")) return false;
    *d2h += (uint64_t)n_seq * 4;
    return true;
}

}  // namespace pq

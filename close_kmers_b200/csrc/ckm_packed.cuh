// 5-bit packed residues: the host-side form of a batch for callers that pass every residue through a parser anyway
// (host/seq_parser.cc does) and want to move 5 bits instead of 8 per residue over PCIe -- end to end the calling path is
// bound by that copy, and several GPUs share one host's memory bandwidth.
//
// Format.  Sequence i is the bit stream in words [word_offsets[i], word_offsets[i+1]) of `packed` (32-bit little-endian
// words, residue r in bits [5r, 5r+5) of the stream): codes 0..19 = ACDEFGHIKLMNPQRSTVWY (kguts.cc:273-339), 31 = any other
// character, 30 = end of the sequence (an embedded NUL ends the reference's scan, kguts.cc:791; the packer also writes it
// into the slots left over in the last word).  A sequence of L residues takes ceil(5 L / 32) words.
//
// On the device the stream is unpacked to the ASCII layout every kernel reads, one warp per sequence: sequence i lands at
// residue offset 8 * word_offsets[i] (a word holds 6.4 residues, so eight slots per word always suffice), real residues
// first, NULs behind them -- which is exactly how the kernels already see a protein that ends early (strlen semantics), so
// neither true lengths nor residue offsets have to be uploaded.  0.1 ms per million proteins.
#pragma once
#include "ckm_common.cuh"

namespace ckm {

constexpr uint32_t kPackEnd = 30u, kPackInvalid = 31u;

__global__ void __launch_bounds__(256)
unpack5_kernel(const uint32_t *__restrict__ packed, const uint64_t *__restrict__ woff /* n + 1, rebased to word 0 of `packed` */,
               uint32_t n, uint64_t woff_base /* device word offset of woff[0]'s sequence */, uint8_t *__restrict__ residues,
               uint64_t *__restrict__ offsets /* n + 1 residue offsets, written here */) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    // code -> ASCII, four codes at a time from a 32-byte table in registers would cost more than this switch-free form:
    // "ACDEFGHIKLMNPQRSTVWY" + 10 x 'X' + NUL (30) + 'X' (31)
    for (uint32_t i = warp0; i <= n; i += n_warps) {
        const uint64_t w0 = __ldg(woff + i);
        if (lane == 0) offsets[i] = 8ull * (woff_base + w0);
        if (i == n) break;
        const uint32_t words = (uint32_t)(__ldg(woff + i + 1) - w0);
        const uint32_t slots = 8u * words, coded = (32u * words) / 5u;  // residue slots on the device; codes the stream holds
        const uint32_t *src = packed + w0;
        uint32_t *dst = reinterpret_cast<uint32_t *>(residues + 8ull * (woff_base + w0));
        for (uint32_t r0 = 4u * lane; r0 < slots; r0 += 128u) {
            uint32_t out = 0;
            if (r0 < coded) {
                const uint32_t bit = 5u * r0, wi = bit >> 5, sh = bit & 31u;
                const uint32_t a = __ldg(src + wi), b = (wi + 1u < words) ? __ldg(src + wi + 1u) : 0u;
                const uint32_t x = __funnelshift_r(a, b, sh);  // 20 bits: four codes
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t c = (x >> (5 * k)) & 31u;
                    uint32_t ch = 'X';
                    if (c < 20u) ch = (uint32_t)"ACDEFGHIKLMNPQRSTVWY"[c];
                    if (c == kPackEnd || r0 + k >= coded) ch = 0u;
                    out |= ch << (8 * k);
                }
            }
            dst[r0 >> 2] = out;
        }
    }
}

}  // namespace ckm

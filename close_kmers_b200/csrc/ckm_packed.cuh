// Packed residues: the host-side form of a batch for callers that pass every residue through a parser anyway
// (host/seq_parser.cc does) and want to move fewer than 8 bits per residue over PCIe -- end to end the calling path is bound
// by that copy, and several GPUs share one host's memory bandwidth.
//
// Format.  Sequence i is the 32-bit words [word_offsets[i], word_offsets[i+1]) of `packed`; a word holds SEVEN residues as
// the digits of a base-22 number, residue 7 w + k of the sequence = (word_w / 22^k) % 22 (22^7 < 2^32: 4.57 bits per residue,
// a quarter of a bit above the entropy of twenty letters): 0..19 = ACDEFGHIKLMNPQRSTVWY (kguts.cc:273-339), 20 = any other
// character, 21 = end of the sequence (an embedded NUL ends the reference's scan, kguts.cc:791; the packer also writes it
// into the digits left over in the last word).  A sequence of L residues takes ceil(L / 7) words.
//
// On the device the words are unpacked to the ASCII layout every kernel reads: sequence i lands at residue offset
// 8 * word_offsets[i] (eight slots per word always suffice), real residues first, NULs behind them -- which is exactly how
// the kernels already see a protein that ends early (strlen semantics), so neither true lengths nor residue offsets have to
// be uploaded.  A lane decodes one word (seven divisions by 22 as multiplications), the warp's 32 x 7 bytes are put in order in
// shared memory and leave as aligned 4-byte stores.
#pragma once
#include "ckm_common.cuh"

namespace ckm {

constexpr uint32_t kPackBase = 22u, kPackPerWord = 7u, kPackInvalid = 20u, kPackEnd = 21u;
constexpr int kUnpackThreads = 256;

__global__ void __launch_bounds__(kUnpackThreads)
unpack7_kernel(const uint32_t *__restrict__ packed, const uint64_t *__restrict__ woff /* n + 1, rebased to word 0 of `packed` */,
               uint32_t n, uint64_t woff_base /* device word offset of woff[0]'s sequence */, uint8_t *__restrict__ residues,
               uint64_t *__restrict__ offsets /* n + 1 residue offsets, written here */) {
    __shared__ __align__(16) uint8_t s_bytes[kUnpackThreads / 32][32 * kPackPerWord + 4];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    uint8_t *sb = s_bytes[wid];
    for (uint32_t i = warp0; i <= n; i += n_warps) {  // one warp per sequence
        const uint64_t w0 = __ldg(woff + i);
        if (lane == 0) offsets[i] = 8ull * (woff_base + w0);
        if (i == n) break;
        const uint32_t words = (uint32_t)(__ldg(woff + i + 1) - w0);
        const uint32_t slots = 8u * words, coded = kPackPerWord * words;  // residue slots on the device; digits the words hold
        const uint32_t *src = packed + w0;
        uint32_t *dst = reinterpret_cast<uint32_t *>(residues + 8ull * (woff_base + w0));
        // 32 words = 224 residues = 56 output words per round (7 * 32 is a multiple of four: rounds start word-aligned)
        for (uint32_t wb = 0; wb < words; wb += 32u) {
            uint32_t x = wb + lane < words ? __ldg(src + wb + lane) : 0u;
#pragma unroll
            for (int k = 0; k < (int)kPackPerWord; k++) {
                const uint32_t q = __umulhi(x, 0xBA2E8BA3u) >> 4;  // x / 22
                const uint32_t c = x - q * kPackBase;
                x = q;
                uint32_t ch = 'X';
                if (c < 20u) ch = (uint32_t)"ACDEFGHIKLMNPQRSTVWY"[c];
                if (c == kPackEnd || wb + lane >= words) ch = 0u;  // (lanes past the last word: the NULs behind the sequence)
                sb[kPackPerWord * lane + k] = (uint8_t)ch;
            }
            __syncwarp();
            const uint32_t r_base = kPackPerWord * wb;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                const uint32_t o = lane + 32u * t, r0 = r_base + 4u * o;  // output word within the round, its first residue slot
                if (o < 56u && r0 < coded) dst[r0 >> 2] = reinterpret_cast<const uint32_t *>(sb)[o];
            }
            __syncwarp();
        }
        // eight slots per word, seven digits: the slots behind them are NULs
        for (uint32_t r0 = ((coded + 3u) & ~3u) + 4u * lane; r0 < slots; r0 += 128u) dst[r0 >> 2] = 0u;
    }
}

}  // namespace ckm

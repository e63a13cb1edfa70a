// fastq reads on the GPU: 6-frame translation (TranslationTable code 11, trans_table.cc:8-84; frames and reverse
// complement, dna_seq.cc:9-47, dna_seq.h:28-111), fragment split on stops, then the standard calling + family
// pipeline per fragment and the best-frame selection of FqProcessRequest::on_parsed_seq
// (fq_process_request.cc:298-365).  Included at the end of ckm_api.cu.
//
// Layout: one thread per (read, frame slot), slots ordered {1,2,3,-1,-2,-3} like dna_seq.cc:13.  A first pass
// counts the fragments longer than `min_len` and their residues; two prefix sums place every frame's
// fragments; a second pass writes the fragment batch (residues + CSR offsets) that K1/K2 and the family
// kernels then consume unchanged -- a fragment is just a protein.
#pragma once

namespace ckm {

// NCBI genetic code 11 in the TCAG order of the published table (the text block of trans_table.cc:8-15)
__constant__ char kNcbi11[65] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";

// TranslationTable::encode_char, trans_table.h:47-68
__device__ __forceinline__ uint32_t base_code(uint8_t c) {
    switch (c) {
        case 'a': case 'A': return 0u;
        case 'c': case 'C': return 1u;
        case 'g': case 'G': return 2u;
        case 't': case 'u': case 'T': case 'U': return 3u;
        default: return 4u;
    }
}

// aa_table_ re-indexed by encode_triple (16*b1+4*b2+b3 with A=0,C=1,G=2,T=3), slot 64 = 'X' (trans_table.cc:41-62)
__device__ __forceinline__ void fill_aa11(char *tbl /* 65, shared */) {
    if (threadIdx.x < 64) {
        const uint32_t pos = threadIdx.x;
        const uint32_t tcag[4] = {3u, 1u, 0u, 2u};  // T, C, A, G as encode_char codes
        tbl[tcag[pos >> 4] * 16 + tcag[(pos >> 2) & 3] * 4 + tcag[pos & 3]] = kNcbi11[pos];
    }
    if (threadIdx.x == 64) tbl[64] = 'X';
}

// base_code for every byte value, for branch-free lookups
__device__ __forceinline__ void fill_base_lut(uint8_t *lut /* 256, shared */) {
    for (uint32_t c = threadIdx.x; c < 256u; c += blockDim.x) lut[c] = (uint8_t)base_code((uint8_t)c);
}

// amino acid of codon k of frame slot `slot` (0..2 forward, 3..5 reverse complement) of read[0..len).  CODES: `read`
// already holds base codes (the staged copy), else raw bytes that go through `blut`.  No branches: lanes of one warp
// serve forward and reverse frames side by side.
template <bool CODES>
__device__ __forceinline__ char frame_aa(const uint8_t *read, uint32_t len, uint32_t slot, uint32_t k, const char *tbl,
                                         const uint8_t *blut) {
    const uint32_t off = (slot % 3u) + 3u * k;
    const bool rev = slot >= 3u;
    // reverse_seq(): complement of the reversed read; complement() keeps non-ACGTU letters ambiguous
    const uint32_t i0 = rev ? len - 1u - off : off, i1 = rev ? i0 - 1u : i0 + 1u, i2 = rev ? i0 - 2u : i0 + 2u;
    uint32_t c0 = CODES ? read[i0] : blut[read[i0]], c1 = CODES ? read[i1] : blut[read[i1]], c2 = CODES ? read[i2] : blut[read[i2]];
    const bool ok = (c0 | c1 | c2) < 4u;
    if (rev) { c0 = 3u - c0; c1 = 3u - c1; c2 = 3u - c2; }
    return tbl[ok ? c0 * 16u + c1 * 4u + c2 : 64u];
}

constexpr uint32_t kFqReadsPerBlock = 42;   // 252 of the block's 256 threads
constexpr uint32_t kFqStageBytes = 12 * 1024;
constexpr uint32_t kFqOutBytes = 16 * 1024;  // 42 reads x 6 frames x 50 residues = 12.6 KB for 150-base reads

template <bool FILL>
__global__ void __launch_bounds__(256)
fq_frames_kernel(const uint8_t *__restrict__ bases, const uint64_t *__restrict__ offsets, uint32_t n, uint32_t min_len,
                 uint32_t *__restrict__ nfrag, uint32_t *__restrict__ naa,                      // count pass outputs
                 const uint64_t *__restrict__ frag_base, const uint64_t *__restrict__ res_base,  // fill pass inputs
                 uint64_t *__restrict__ frag_off, uint8_t *__restrict__ frag_res) {
    __shared__ char tbl[65];
    __shared__ uint8_t blut[256];
    __shared__ __align__(16) uint8_t s_bases[kFqStageBytes + 16];
    __shared__ __align__(16) uint8_t s_out[FILL ? kFqOutBytes + 32 : 16];
    fill_aa11(tbl);
    fill_base_lut(blut);
    // the block's reads (kFqReadsPerBlock consecutive ones, six threads each) are staged in shared memory with
    // coalesced word loads when they fit; every base is then read six times (once per frame) from there
    const uint32_t r0 = blockIdx.x * kFqReadsPerBlock;
    const uint32_t r1 = min(n, r0 + kFqReadsPerBlock);
    const uint64_t blk0 = r0 < n ? offsets[r0] : 0, blk1 = r0 < n ? offsets[r1] : 0;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(bases + blk0) & 3u);
    const bool staged = (blk1 - blk0) + mis <= kFqStageBytes;
    __syncthreads();  // blut is read below
    if (staged) {  // ... as base codes (trans_table.h:47-68), four per word
        const uint32_t *src = reinterpret_cast<const uint32_t *>(bases + blk0 - mis);
        const uint32_t words = (uint32_t)((blk1 - blk0) + mis + 3u) >> 2;
        for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) {
            const uint32_t v = __ldg(src + w);
            reinterpret_cast<uint32_t *>(s_bases)[w] = (uint32_t)blut[v & 0xFFu] | ((uint32_t)blut[(v >> 8) & 0xFFu] << 8) |
                                                       ((uint32_t)blut[(v >> 16) & 0xFFu] << 16) | ((uint32_t)blut[v >> 24] << 24);
        }
    }
    // fill pass: the block's fragments are one contiguous range of the output (frames are laid out in thread order), so
    // residues are assembled in shared memory, at the same 16-byte phase as their destination, and leave as whole
    // 16-byte stores instead of one byte per thread per instruction
    uint64_t out0 = 0, out1 = 0;
    uint32_t ophase = 0;
    bool out_staged = false;
    if (FILL && r0 < n) {
        out0 = res_base[6ull * r0];
        out1 = res_base[6ull * r1];
        ophase = (uint32_t)(reinterpret_cast<uintptr_t>(frag_res + out0) & 15u);
        out_staged = (out1 - out0) + ophase <= kFqOutBytes;
    }
    __syncthreads();
    const uint32_t r = r0 + threadIdx.x / 6u, slot = threadIdx.x % 6u;
    if (threadIdx.x < 6u * kFqReadsPerBlock && r < n) {
        const uint64_t t = 6ull * r + slot;
        const uint64_t b0 = offsets[r];
        const uint32_t len = (uint32_t)(offsets[r + 1] - b0);
        const uint8_t *read = staged ? s_bases + mis + (b0 - blk0) : bases + b0;
        const uint32_t skip = slot % 3u;
        const uint32_t ncod = len >= skip + 3u ? (len - skip) / 3u : 0u;  // complete codons only (trans_table.cc:70-81)
        uint32_t cnt = 0, aa_total = 0, run = 0;
        uint64_t fcur = FILL ? frag_base[t] : 0, rcur = FILL ? res_base[t] : 0;
        // boost::split(.., "*", token_compress_on) (dna_seq.cc:17): fragments are the maximal stop-free runs
        if (FILL && out_staged && ncod <= 64u) {
            // Lanes meet their stops at different codons, so copying "the run that just ended" would serialise the warp.
            // Instead: one pass marks the codons of kept runs in a 64-bit mask (and emits the fragment offsets), a second
            // pass -- the same trip count in every lane -- stores the marked residues.
            unsigned long long keep = 0ull;
            const uint64_t r_begin = rcur;
            for (uint32_t k = 0; k <= ncod; k++) {
                const char a = k < ncod ? (staged ? frame_aa<true>(read, len, slot, k, tbl, blut) : frame_aa<false>(read, len, slot, k, tbl, blut)) : '*';
                if (a != '*') { run++; continue; }
                if (run > min_len) {
                    frag_off[fcur++] = rcur;
                    rcur += run;
                    keep |= (run >= 64u ? ~0ull : ((1ull << run) - 1ull)) << (k - run);
                }
                run = 0;
            }
            uint8_t *dst = s_out + ophase + (uint32_t)(r_begin - out0);
            for (uint32_t k = 0; k < ncod; k++)
                if ((keep >> k) & 1ull) *dst++ = (uint8_t)(staged ? frame_aa<true>(read, len, slot, k, tbl, blut) : frame_aa<false>(read, len, slot, k, tbl, blut));
        } else {
            for (uint32_t k = 0; k <= ncod; k++) {
                const char a = k < ncod ? (staged ? frame_aa<true>(read, len, slot, k, tbl, blut) : frame_aa<false>(read, len, slot, k, tbl, blut)) : '*';
                if (a != '*') { run++; continue; }
                if (run > min_len) {  // prot.length() > 10 (fq_process_request.cc:331)
                    if (FILL) {
                        frag_off[fcur++] = rcur;
                        if (out_staged) {
                            uint8_t *dst = s_out + ophase + (uint32_t)(rcur - out0);
                            for (uint32_t j = k - run; j < k; j++) *dst++ = (uint8_t)(staged ? frame_aa<true>(read, len, slot, j, tbl, blut) : frame_aa<false>(read, len, slot, j, tbl, blut));
                            rcur += run;
                        } else {
                            for (uint32_t j = k - run; j < k; j++) frag_res[rcur++] = (uint8_t)(staged ? frame_aa<true>(read, len, slot, j, tbl, blut) : frame_aa<false>(read, len, slot, j, tbl, blut));
                        }
                    }
                    cnt++;
                    aa_total += run;
                }
                run = 0;
            }
        }
        if (!FILL) { nfrag[t] = cnt; naa[t] = aa_total; }
    }
    if (FILL) {
        __syncthreads();
        if (out_staged && out1 > out0) {
            uint8_t *gbase = frag_res + out0 - ophase;  // 16-byte aligned
            const uint32_t span = ophase + (uint32_t)(out1 - out0);
            const uint32_t first_full = ophase ? 1u : 0u, n_vec = span >> 4;  // vectors [first_full, n_vec) are fully ours
            for (uint32_t v = first_full + threadIdx.x; v < n_vec; v += blockDim.x)
                reinterpret_cast<uint4 *>(gbase)[v] = reinterpret_cast<const uint4 *>(s_out)[v];
            // the partial vectors at both ends belong partly to the neighbouring blocks: bytes
            if (ophase) {
                const uint32_t e = min(16u, span);
                for (uint32_t b = ophase + threadIdx.x; b < e; b += blockDim.x) gbase[b] = s_out[b];
            }
            const uint32_t tail0 = max(n_vec << 4, ophase ? 16u : 0u);
            for (uint32_t b = tail0 + threadIdx.x; b < span; b += blockDim.x) gbase[b] = s_out[b];
        }
    }
}

// best frame per read (fq_process_request.cc:319-348): frames in slot order, score is a double sum of the
// fragments' best-call scores, strictly-greater updates inside the fragment loop
__global__ void __launch_bounds__(256)
fq_reduce_kernel(const uint64_t *__restrict__ frag_base /* 6n+1 */, const ckm_family_match_t *__restrict__ m, uint32_t n,
                 int32_t *__restrict__ best_frame, double *__restrict__ best_score, uint32_t *__restrict__ best_n,
                 uint64_t *__restrict__ best_first) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double bs = 0.0;
    int32_t bf = 0;
    uint32_t bn = 0;
    uint64_t bfirst = 0;
    for (uint32_t slot = 0; slot < 6; slot++) {
        const uint64_t f0 = frag_base[6ull * r + slot], f1 = frag_base[6ull * r + slot + 1];
        double score = 0.0;
        for (uint64_t f = f0; f < f1; f++) {
            score += (double)m[f].score;
            if (score > bs) {
                bs = score;
                bf = slot < 3 ? (int32_t)slot + 1 : -((int32_t)slot - 2);
                bn = (uint32_t)(f - f0 + 1);
                bfirst = f0;
            }
        }
    }
    best_frame[r] = bs > 0.0 ? bf : 0;
    best_score[r] = bs;
    best_n[r] = bs > 0.0 ? bn : 0u;
    best_first[r] = bfirst;
}

__global__ void __launch_bounds__(256)
fq_gather_kernel(const uint64_t *__restrict__ match_off, const uint64_t *__restrict__ best_first,
                 const uint64_t *__restrict__ frag_off, const ckm_family_match_t *__restrict__ m, uint32_t n,
                 ckm_fq_match_t *__restrict__ out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const uint64_t o0 = match_off[r];
    const uint32_t cnt = (uint32_t)(match_off[r + 1] - o0);
    const uint64_t f0 = best_first[r];
    for (uint32_t k = 0; k < cnt; k++) {
        ckm_fq_match_t x;
        x.length = (uint32_t)(frag_off[f0 + k + 1] - frag_off[f0 + k]);
        x.m = m[f0 + k];
        out[o0 + k] = x;
    }
}

}  // namespace ckm

// reads -> fragment batch on the device; leaves c->fq.{frag_base, frag_off, frag_res} and the counts
static int fq_translate_device(ckm_ctx *c, const char *bases, const uint64_t *offsets, uint32_t n, uint32_t min_len,
                               uint64_t *n_frags_out, uint64_t *n_aa_out, uint32_t *max_read_out) {
    ckm_ctx::Fq &Q = c->fq;
    uint64_t total = 0;
    uint32_t max_len = 0;
    RC(upload_batch(c, bases, offsets, n, &total, &max_len));  // reads live in in_res / in_off during translation
    const uint64_t nt = 6ull * n;
    RC(Q.nfrag.ensure((nt + 1) * 4));
    RC(Q.naa.ensure((nt + 1) * 4));
    RC(Q.frag_base.ensure((nt + 2) * 8));
    RC(Q.res_base.ensure((nt + 2) * 8));
    uint64_t n_frags = 0, n_aa = 0;
    if (n) {
        const unsigned blocks = (n + ckm::kFqReadsPerBlock - 1) / ckm::kFqReadsPerBlock;
        fq_frames_kernel<false><<<blocks, 256, 0, c->stream>>>((const uint8_t *)c->in_res.p, (const uint64_t *)c->in_off.p, n,
                                                              min_len, (uint32_t *)Q.nfrag.p, (uint32_t *)Q.naa.p, nullptr,
                                                              nullptr, nullptr, nullptr);
        c->launches++;
        RC(prefix_sum(c, (const uint32_t *)Q.nfrag.p, nt, (uint64_t *)Q.frag_base.p));
        RC(prefix_sum(c, (const uint32_t *)Q.naa.p, nt, (uint64_t *)Q.res_base.p));
        CU(cudaMemcpyAsync(&n_frags, (const uint64_t *)Q.frag_base.p + nt, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(&n_aa, (const uint64_t *)Q.res_base.p + nt, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
    } else {
        CU(cudaMemsetAsync(Q.frag_base.p, 0, 8, c->stream));
    }
    RC(Q.frag_off.ensure((n_frags + 2) * 8));
    RC(Q.frag_res.ensure(n_aa + 64));
    if (n) {
        const unsigned blocks = (n + ckm::kFqReadsPerBlock - 1) / ckm::kFqReadsPerBlock;
        fq_frames_kernel<true><<<blocks, 256, 0, c->stream>>>((const uint8_t *)c->in_res.p, (const uint64_t *)c->in_off.p, n,
                                                             min_len, nullptr, nullptr, (const uint64_t *)Q.frag_base.p,
                                                             (const uint64_t *)Q.res_base.p, (uint64_t *)Q.frag_off.p,
                                                             (uint8_t *)Q.frag_res.p);
        c->launches++;
    }
    CU(cudaMemcpyAsync((uint64_t *)Q.frag_off.p + n_frags, &n_aa, 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemsetAsync((uint8_t *)Q.frag_res.p + n_aa, 0, 32, c->stream));
    CU(cudaStreamSynchronize(c->stream));  // n_aa is a stack variable
    CU(cudaGetLastError());
    *n_frags_out = n_frags;
    *n_aa_out = n_aa;
    *max_read_out = max_len;
    return 0;
}

extern "C" int ckm_fq_translate(ckm_ctx *c, const char *bases, const uint64_t *offsets, uint32_t n, uint32_t min_len,
                                ckm_fq_fragments_t *out) {
    if (!c || !out) return ckm_fail(CKM_EINVAL, "NULL argument");
    memset(out, 0, sizeof *out);
    uint64_t n_frags = 0, n_aa = 0;
    uint32_t max_read = 0;
    RC(fq_translate_device(c, bases, offsets, n, min_len, &n_frags, &n_aa, &max_read));
    ckm_ctx::Fq &Q = c->fq;
    RC(Q.h_frag_base.ensure((6ull * n + 2) * 8));
    RC(Q.h_frag_off.ensure((n_frags + 2) * 8));
    RC(Q.h_frag_res.ensure(n_aa + 64));
    CU(cudaMemcpyAsync(Q.h_frag_base.p, Q.frag_base.p, (6ull * n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(Q.h_frag_off.p, Q.frag_off.p, (n_frags + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    if (n_aa) CU(cudaMemcpyAsync(Q.h_frag_res.p, Q.frag_res.p, n_aa, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    out->n_reads = n;
    out->n_fragments = n_frags;
    out->frag_frame_offsets = (const uint64_t *)Q.h_frag_base.p;
    out->frag_offsets = (const uint64_t *)Q.h_frag_off.p;
    out->residues = (const char *)Q.h_frag_res.p;
    return 0;
}

extern "C" int ckm_fq_batch(ckm_ctx *c, const char *bases, const uint64_t *offsets, uint32_t n, ckm_fq_out_t *out) {
    if (!c || !out) return ckm_fail(CKM_EINVAL, "NULL argument");
    if (!c->fam.loaded) return ckm_fail(CKM_ESTATE, "ckm_fq_batch before ckm_family_load");
    memset(out, 0, sizeof *out);
    out->n = n;
    uint64_t n_frags = 0, n_aa = 0;
    uint32_t max_read = 0;
    RC(fq_translate_device(c, bases, offsets, n, 10u, &n_frags, &n_aa, &max_read));
    ckm_ctx::Fq &Q = c->fq;
    // every fragment through find_best_family_match: K1 + K2 + family kernels on the fragment batch
    if (n_frags >= (1ull << 32)) return ckm_fail(CKM_EINVAL, "more than 2^32 fragments in one batch");
    RC(run_device(c, (const uint8_t *)Q.frag_res.p, (const uint64_t *)Q.frag_off.p, (uint32_t)n_frags, n_aa,
                  std::max(max_read / 3u + 1u, 1u), CKM_WANT_HITS | CKM_WANT_CALLS | CKM_WANT_BEST));
    RC(family_device(c, (const uint64_t *)Q.frag_off.p, (uint32_t)n_frags, n_aa));
    RC(Q.best_frame.ensure(((size_t)n + 1) * 4));
    RC(Q.best_score.ensure(((size_t)n + 1) * 8));
    RC(Q.best_n.ensure(((size_t)n + 1) * 4));
    RC(Q.best_first.ensure(((size_t)n + 1) * 8));
    RC(Q.match_off.ensure(((size_t)n + 2) * 8));
    uint64_t n_matches = 0;
    if (n) {
        const unsigned blocks = (n + 255) / 256;
        fq_reduce_kernel<<<blocks, 256, 0, c->stream>>>((const uint64_t *)Q.frag_base.p, (const ckm_family_match_t *)c->fam.matches.p,
                                                        n, (int32_t *)Q.best_frame.p, (double *)Q.best_score.p,
                                                        (uint32_t *)Q.best_n.p, (uint64_t *)Q.best_first.p);
        c->launches++;
        RC(prefix_sum(c, (const uint32_t *)Q.best_n.p, n, (uint64_t *)Q.match_off.p));
        CU(cudaMemcpyAsync(&n_matches, (const uint64_t *)Q.match_off.p + n, 8, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        RC(Q.matches.ensure((n_matches + 1) * sizeof(ckm_fq_match_t)));
        fq_gather_kernel<<<blocks, 256, 0, c->stream>>>((const uint64_t *)Q.match_off.p, (const uint64_t *)Q.best_first.p,
                                                        (const uint64_t *)Q.frag_off.p, (const ckm_family_match_t *)c->fam.matches.p,
                                                        n, (ckm_fq_match_t *)Q.matches.p);
        c->launches++;
    } else {
        CU(cudaMemsetAsync(Q.match_off.p, 0, 8, c->stream));
    }
    RC(Q.h_best_frame.ensure(((size_t)n + 1) * 4));
    RC(Q.h_best_score.ensure(((size_t)n + 1) * 8));
    RC(Q.h_match_off.ensure(((size_t)n + 2) * 8));
    RC(Q.h_matches.ensure((n_matches + 1) * sizeof(ckm_fq_match_t)));
    if (n) {
        CU(cudaMemcpyAsync(Q.h_best_frame.p, Q.best_frame.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(Q.h_best_score.p, Q.best_score.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaMemcpyAsync(Q.h_match_off.p, Q.match_off.p, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    if (n_matches)
        CU(cudaMemcpyAsync(Q.h_matches.p, Q.matches.p, n_matches * sizeof(ckm_fq_match_t), cudaMemcpyDeviceToHost, c->stream));
    uint64_t totals[3] = {0, 0, 0};
    CU(cudaMemcpyAsync(totals, c->totals.p, 24, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    out->best_frame = (const int32_t *)Q.h_best_frame.p;
    out->best_score = (const double *)Q.h_best_score.p;
    out->match_offsets = (const uint64_t *)Q.h_match_off.p;
    out->matches = (const ckm_fq_match_t *)Q.h_matches.p;
    out->n_fragments = n_frags;
    out->n_probes = totals[0];
    return 0;
}

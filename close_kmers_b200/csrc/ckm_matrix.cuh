// /add postings and /matrix pairwise shared-k-mer counts on the GPU (add_request.cc:133, 164-170;
// kmer.cc:174-214; matrix_request.cc:78-95, 130-161).  Included at the end of ckm_api.cu.
//
// Postings: every hit occurrence of an /add-ed protein appends (k-mer, peg id).  Before the first /matrix
// after a change they are indexed WITHOUT sorting: distinct k-mers are inserted into an open-addressed table
// while their multiplicities are counted, a prefix sum turns counts into list offsets, and a second pass
// drops every peg id into its k-mer's list (order inside a list is irrelevant to the counts).  The index has
// the same {key+1, offset, length} slot format as the family table, so fam_lookup_kernel serves both.
//
// Matrix rows: one warp per protein of the row block.  Its hits' posting lists are walked with lanes across
// list entries; partners are accumulated in a per-warp open-addressed map keyed by peg id (shared memory, or
// a global scratch region when the row touches more than kFamSmemE postings) with atomic adds, then the map
// is compacted into COO entries.
#pragma once

namespace ckm {

__global__ void __launch_bounds__(256)
post_append_kernel(const uint64_t *__restrict__ offsets, const uint64_t *__restrict__ hit_keys,
                   const uint64_t *__restrict__ hit_off, const uint32_t *__restrict__ eids, uint32_t n, uint64_t n0,
                   uint64_t *__restrict__ pkeys, uint32_t *__restrict__ peids) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    const uint64_t base = offsets[w], o0 = hit_off[w];
    const uint32_t cnt = (uint32_t)(hit_off[w + 1] - o0), e = eids[w];
    for (uint32_t k = lane; k < cnt; k += 32) {
        pkeys[n0 + o0 + k] = hit_keys[base + k];
        peids[n0 + o0 + k] = e;
    }
}

__device__ __forceinline__ uint64_t post_slot(unsigned long long *tkeys, uint64_t mask, uint64_t key) {
    uint64_t s = fam_hash(key) & mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(&tkeys[s], 0ull, (unsigned long long)(key + 1));
        if (prev == 0ull || prev == key + 1) return s;
        s = (s + 1) & mask;
    }
}

__global__ void __launch_bounds__(256)
post_count_kernel(const uint64_t *__restrict__ pkeys, uint64_t n, unsigned long long *tkeys, uint64_t mask, uint32_t *tcnt) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    atomicAdd(&tcnt[post_slot(tkeys, mask, pkeys[i])], 1u);
}

__global__ void __launch_bounds__(256)
post_fill_kernel(const uint64_t *__restrict__ pkeys, const uint32_t *__restrict__ peids, uint64_t n, unsigned long long *tkeys,
                 uint64_t mask, const uint64_t *__restrict__ toff, uint32_t *tcur, uint32_t *__restrict__ ids) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t s = post_slot(tkeys, mask, pkeys[i]);
    ids[toff[s] + atomicAdd(&tcur[s], 1u)] = peids[i];
}

__global__ void __launch_bounds__(256)
post_pack_kernel(const unsigned long long *__restrict__ tkeys, const uint32_t *__restrict__ tcnt,
                 const uint64_t *__restrict__ toff, uint64_t cap, FamSlot *__restrict__ slots, unsigned long long *n_occupied) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // cap is a multiple of 16 >= 16, blocks are whole warps
    FamSlot f;
    f.key1 = s < cap ? tkeys[s] : 0ull;
    if (s < cap) {
        f.off = (uint32_t)toff[s];
        f.cnt = tcnt[s];
        slots[s] = f;
    }
    const unsigned occ = __popc(__ballot_sync(0xffffffffu, f.key1 != 0ull));
    if ((threadIdx.x & 31u) == 0 && occ) atomicAdd(n_occupied, (unsigned long long)occ);
}

// capacity of row i's output region: it has at most min(E_i, i) distinct partners
__global__ void __launch_bounds__(256)
matrix_rowcap_kernel(const uint32_t *__restrict__ E, uint32_t n_rows, uint32_t row_begin, uint32_t *__restrict__ rcap, bool filter) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    rcap[r] = filter ? min(E[r], row_begin + r) : E[r];
}

constexpr int kMatWarps = 8;
constexpr size_t kMatSmem = (size_t)kMatWarps * (2 * kFamSmemCap + 32 * kFamStage) * 4;

__global__ void __launch_bounds__(kMatWarps * 32)
matrix_row_kernel(const uint32_t *__restrict__ post_ids, const uint64_t *__restrict__ offsets /* row-block batch */,
                  const uint32_t *__restrict__ n_hits, const uint2 *__restrict__ hit_fam, const uint32_t *__restrict__ E,
                  const uint32_t *__restrict__ gcap, const uint64_t *__restrict__ gofs, uint32_t *__restrict__ gscratch,
                  const uint32_t *__restrict__ eids /* whole request */, const uint32_t *__restrict__ first_idx, uint32_t max_eid,
                  uint32_t n_rows, uint32_t row_begin, const uint64_t *__restrict__ rofs, ckm_pair_t *__restrict__ entries,
                  uint32_t *__restrict__ n_distinct, bool filter) {
    extern __shared__ __align__(16) uint32_t mat_smem[];
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    uint32_t *const my = mat_smem + (size_t)wib * (2 * kFamSmemCap + 32 * kFamStage);
    uint32_t(*const s_stage)[kFamStage] = reinterpret_cast<uint32_t(*)[kFamStage]>(my + 2 * kFamSmemCap);
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); r < n_rows; r += n_warps) {
        const uint32_t i = row_begin + r;  // index in the request
        const uint32_t me = filter ? eids[i] : i;  // unfiltered (LookupRequest::on_hit, lookup_request.cc:466-478): row index
        const uint64_t base = offsets[r];
        const uint32_t nh = n_hits[r];
        uint32_t *keys, *cnt, cap;
        if (gcap[r] == 0) {
            cap = 32u;  // at most E distinct partners, map kept at most half full
            while (cap < 2u * E[r]) cap <<= 1;
            keys = my;
            cnt = my + kFamSmemCap;
            for (uint32_t s = lane; s < cap; s += 32) { keys[s] = 0u; cnt[s] = 0u; }
        } else {  // pre-zeroed global scratch: keys[cap] then counts[cap]
            cap = gcap[r];
            keys = gscratch + gofs[r] * 2ull;
            cnt = keys + cap;
        }
        const uint32_t mask = cap - 1;
        __syncwarp();
        if (E[r] != 0) {
            for (uint32_t k0 = 0; k0 < nh; k0 += 32) {
                const uint2 mine = (k0 + lane < nh) ? hit_fam[base + k0 + lane] : make_uint2(0u, 0u);
                for (uint32_t t = 0; t < min(mine.y, (uint32_t)kFamStage); t++) s_stage[lane][t] = post_ids[mine.x + t];
                __syncwarp();
                const uint32_t lim = min(32u, nh - k0);
                for (uint32_t j = 0; j < lim; j++) {  // MatrixRequest::on_hit, matrix_request.cc:130-161
                    const uint32_t off_j = __shfl_sync(0xffffffffu, mine.x, j);
                    const uint32_t cnt_j = __shfl_sync(0xffffffffu, mine.y, j);
                    for (uint32_t t = lane; t < cnt_j; t += 32) {
                        const uint32_t e = t < (uint32_t)kFamStage ? s_stage[j][t] : post_ids[off_j + t];
                        // eid != id, and eid already in matrix_proteins_ (set at line 90 before protein i runs)
                        if (filter && (e == me || e > max_eid || first_idx[e] > i)) continue;
                        bool fresh;
                        atomicAdd(&cnt[map_slot(keys, mask, e + 1, &fresh)], 1u);
                    }
                }
                __syncwarp();
            }
        }
        __syncwarp();
        // compact the map into this row's output region
        ckm_pair_t *out = entries + rofs[r];
        uint32_t total = 0;
        for (uint32_t s0 = 0; s0 < cap; s0 += 32) {
            const uint32_t s = s0 + lane;
            const uint32_t key = keys[s];
            const bool q = key != 0u;
            const uint32_t m = __ballot_sync(0xffffffffu, q);
            if (q) {
                ckm_pair_t p;
                p.eid_i = me;
                p.eid_j = key - 1;
                p.count = cnt[s];
                out[total + __popc(m & ((1u << lane) - 1u))] = p;
            }
            total += __popc(m);
        }
        if (lane == 0) n_distinct[r] = total;
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256)
matrix_export_kernel(const uint64_t *__restrict__ rofs, const ckm_pair_t *__restrict__ entries,
                     const uint64_t *__restrict__ out_off, uint32_t n_rows, ckm_pair_t *__restrict__ out, bool sorted) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n_rows) return;
    const uint64_t o0 = out_off[r];
    const uint32_t cnt = (uint32_t)(out_off[r + 1] - o0);
    const ckm_pair_t *src = entries + rofs[r];
    for (uint32_t k = lane; k < cnt; k += 32) {
        uint32_t dst = k;
        if (sorted) {  // rank by partner id (distinct within a row)
            dst = 0;
            for (uint32_t j = 0; j < cnt; j++) dst += src[j].eid_j < src[k].eid_j;
        }
        out[o0 + dst] = src[k];
    }
}

}  // namespace ckm

// grow a device buffer keeping its first `keep` bytes
static int grow_preserve(ckm_ctx *c, DevBuf &b, size_t keep, size_t want) {
    if (want <= b.cap) return 0;
    DevBuf nb;
    RC(nb.ensure(want + want / 2));
    if (keep && b.p) CU(cudaMemcpyAsync(nb.p, b.p, keep, cudaMemcpyDeviceToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    b.release();
    b = nb;
    return 0;
}

// append (k-mer, eid) for every hit of the batch K1 just processed (hit_keys / n_hits on the device)
static int postings_append_device(ckm_ctx *c, ckm_ctx::Post &P, const uint32_t *eids, const uint64_t *d_off, uint32_t n) {
    if (n == 0) return 0;
    uint64_t totals[3];
    RC(ckm_read_totals(c, totals));
    const uint64_t nh = totals[1];
    RC(c->hit_off.ensure(((size_t)n + 1) * 8));
    RC(prefix_sum(c, (const uint32_t *)c->n_hits.p, n, (uint64_t *)c->hit_off.p));
    RC(grow_preserve(c, P.keys, P.n * 8, (P.n + nh + 1) * 8));
    RC(grow_preserve(c, P.eids, P.n * 4, (P.n + nh + 1) * 4));
    RC(P.d_eids.ensure(((size_t)n + 1) * 4));
    CU(cudaMemcpyAsync(P.d_eids.p, eids, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    post_append_kernel<<<(unsigned)(((uint64_t)n * 32 + 255) / 256), 256, 0, c->stream>>>(
        d_off, (const uint64_t *)c->hit_keys.p, (const uint64_t *)c->hit_off.p, (const uint32_t *)P.d_eids.p, n, P.n,
        (uint64_t *)P.keys.p, (uint32_t *)P.eids.p);
    c->launches++;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    P.n += nh;
    P.dirty = true;
    return 0;
}

extern "C" int ckm_postings_add(ckm_ctx *c, const uint32_t *eids, const char *residues, const uint64_t *offsets, uint32_t n) {
    if (!c || (n && !eids)) return ckm_fail(CKM_EINVAL, "NULL argument");
    uint64_t total = 0;
    uint32_t max_len = 0;
    RC(upload_batch(c, residues, offsets, n, &total, &max_len));
    RC(run_device(c, (const uint8_t *)c->in_res.p, (const uint64_t *)c->in_off.p, n, total, std::max(max_len, 1u), CKM_WANT_HITS));
    return postings_append_device(c, c->post, eids, (const uint64_t *)c->in_off.p, n);
}

// after a ckm_call_batch(..., flags containing CKM_WANT_HITS, ...) on the same sequences: reuse its hits
extern "C" int ckm_postings_append_last(ckm_ctx *c, const uint32_t *eids, uint32_t n) {
    if (!c || (n && !eids)) return ckm_fail(CKM_EINVAL, "NULL argument");
    if (n != c->cur_n || !(c->cur_flags & CKM_WANT_HITS) || c->cur_off != (const uint64_t *)c->in_off.p)
        return ckm_fail(CKM_ESTATE, "ckm_postings_append_last needs the preceding ckm_call_batch to have asked for hits");
    return postings_append_device(c, c->post, eids, (const uint64_t *)c->in_off.p, n);
}

extern "C" void ckm_postings_clear(ckm_ctx *c) {
    c->post.n = 0;
    c->post.dirty = true;
}
extern "C" uint64_t ckm_postings_count(const ckm_ctx *c) { return c->post.n; }

// One set of postings per KmerPegMapping: the server keeps one mapping per "/mapping/<key>" (krequest2.cc:440-456) and the
// root mapping under key 0.  Selecting parks the current set and brings in (or creates) the requested one.
extern "C" int ckm_postings_select(ckm_ctx *c, uint32_t key) {
    if (!c) return ckm_fail(CKM_EINVAL, "ctx is NULL");
    if (key == c->post_key) return 0;
    c->post_store[c->post_key] = c->post;  // buffers are plain handles: the struct moves by value
    auto it = c->post_store.find(key);
    if (it == c->post_store.end()) {
        c->post = ckm_ctx::Post();
    } else {
        c->post = it->second;
        c->post_store.erase(it);
    }
    c->post_key = key;
    return 0;
}

static int postings_index(ckm_ctx *c, ckm_ctx::Post &P) {
    if (!P.dirty) return 0;
    if (P.n >= (1ull << 32)) return ckm_fail(CKM_EINVAL, "more than 2^32 postings");
    // every key is a signature k-mer that was hit: there are at most num_sigs distinct ones
    uint64_t cap = 16;
    while (cap < 2 * std::min<uint64_t>(P.n, c->num_sigs)) cap <<= 1;
    P.mask = cap - 1;
    RC(P.occ.ensure(64));
    CU(cudaMemsetAsync(P.occ.p, 0, 8, c->stream));
    RC(P.tkeys.ensure(cap * 8));
    RC(P.tcnt.ensure((cap + 1) * 4));
    RC(P.tcur.ensure((cap + 1) * 4));
    RC(P.toff.ensure((cap + 2) * 8));
    RC(P.slots.ensure(cap * sizeof(FamSlot)));
    RC(P.ids.ensure((P.n + 1) * 4));
    CU(cudaMemsetAsync(P.tkeys.p, 0, cap * 8, c->stream));
    CU(cudaMemsetAsync(P.tcnt.p, 0, cap * 4, c->stream));
    CU(cudaMemsetAsync(P.tcur.p, 0, cap * 4, c->stream));
    if (P.n) {
        const unsigned pb = (unsigned)((P.n + 255) / 256);
        post_count_kernel<<<pb, 256, 0, c->stream>>>((const uint64_t *)P.keys.p, P.n, (unsigned long long *)P.tkeys.p, P.mask,
                                                    (uint32_t *)P.tcnt.p);
        c->launches++;
    }
    RC(prefix_sum(c, (const uint32_t *)P.tcnt.p, cap, (uint64_t *)P.toff.p));
    if (P.n) {
        const unsigned pb = (unsigned)((P.n + 255) / 256);
        post_fill_kernel<<<pb, 256, 0, c->stream>>>((const uint64_t *)P.keys.p, (const uint32_t *)P.eids.p, P.n,
                                                   (unsigned long long *)P.tkeys.p, P.mask, (const uint64_t *)P.toff.p,
                                                   (uint32_t *)P.tcur.p, (uint32_t *)P.ids.p);
        c->launches++;
    }
    post_pack_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>((const unsigned long long *)P.tkeys.p,
                                                                          (const uint32_t *)P.tcnt.p, (const uint64_t *)P.toff.p,
                                                                          cap, (FamSlot *)P.slots.p, (unsigned long long *)P.occ.p);
    c->launches++;
    CU(cudaMemcpyAsync(&P.n_keys, P.occ.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    P.dirty = false;
    return 0;
}

// rows [row_begin, row_end) of the request against the selected postings.  filter: the /matrix rules (partner != self and
// already a member of matrix_proteins_); without it every posting of every hit counts (the /lookup peg mode), eid_i is the
// row index, each row's pairs come out by ascending partner id and *h_off (if given) receives the CSR offsets.
static int matrix_rows_impl(ckm_ctx *c, const uint32_t *eids, const char *residues, const uint64_t *offsets, uint32_t n,
                            uint32_t row_begin, uint32_t row_end, bool filter, const ckm_pair_t **pairs, uint64_t *n_pairs,
                            const uint64_t **h_off) {
    CU(cudaSetDevice(c->device));
    ckm_ctx::Post &P = c->post;
    RC(postings_index(c, c->post));
    const uint32_t n_rows = row_end - row_begin;
    // matrix_proteins_ membership: first request index of every id (matrix_request.cc:88-90)
    uint32_t max_eid = 0;
    std::vector<uint32_t> first(1, 0xffffffffu);
    if (filter) {
        for (uint32_t i = 0; i < n; i++) max_eid = std::max(max_eid, eids[i]);
        first.assign((size_t)max_eid + 1, 0xffffffffu);
        for (uint32_t i = 0; i < n; i++)
            if (first[eids[i]] == 0xffffffffu) first[eids[i]] = i;
    }
    RC(P.d_first.ensure(first.size() * 4));
    RC(P.d_eids.ensure(((size_t)n + 1) * 4));
    // the row block's sequences are the batch; hits only (calls == 0, otu == 0: matrix_request.cc:92-94)
    uint64_t total = 0;
    uint32_t max_len = 0;
    RC(upload_batch(c, residues, offsets + row_begin, n_rows, &total, &max_len));
    CU(cudaMemcpyAsync(P.d_first.p, first.data(), first.size() * 4, cudaMemcpyHostToDevice, c->stream));
    if (filter) CU(cudaMemcpyAsync(P.d_eids.p, eids, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    RC(run_device(c, (const uint8_t *)c->in_res.p, (const uint64_t *)c->in_off.p, n_rows, total, std::max(max_len, 1u), CKM_WANT_HITS));
    ckm_ctx::Family &F = c->fam;  // per-batch lookup buffers are shared with the family path
    RC(F.hit_fam.ensure((total + 1) * sizeof(uint2)));
    RC(F.E.ensure(((size_t)n_rows + 1) * 4));
    RC(F.gcap.ensure(((size_t)n_rows + 1) * 4));
    RC(F.gofs.ensure(((size_t)n_rows + 2) * 8));
    RC(P.rcap.ensure(((size_t)n_rows + 1) * 4));
    RC(P.rofs.ensure(((size_t)n_rows + 2) * 8));
    RC(P.nd.ensure(((size_t)n_rows + 1) * 4));
    RC(P.out_off.ensure(((size_t)n_rows + 2) * 8));
    FamTables ft;
    memset(&ft, 0, sizeof ft);
    ft.table = (const FamSlot *)P.slots.p;
    ft.mask = P.mask;
    fam_lookup_kernel<<<(unsigned)(((uint64_t)n_rows * 32 + 255) / 256), 256, 0, c->stream>>>(
        ft, (const uint64_t *)c->in_off.p, (const uint64_t *)c->hit_keys.p, (const uint32_t *)c->n_hits.p, n_rows,
        (uint2 *)F.hit_fam.p, (uint32_t *)F.E.p, (uint32_t *)F.gcap.p);
    matrix_rowcap_kernel<<<(n_rows + 255) / 256, 256, 0, c->stream>>>((const uint32_t *)F.E.p, n_rows, row_begin, (uint32_t *)P.rcap.p, filter);
    c->launches += 2;
    RC(prefix_sum(c, (const uint32_t *)F.gcap.p, n_rows, (uint64_t *)F.gofs.p));
    RC(prefix_sum(c, (const uint32_t *)P.rcap.p, n_rows, (uint64_t *)P.rofs.p));
    uint64_t gtotal = 0, rtotal = 0, gtotal_entries = 0;
    if (c->matrix_tile_on_device) {  // the row block's work, for the caller's report: posting-list entries behind its hits
        RC(P.out_off.ensure(((size_t)n_rows + 2) * 8));
        RC(prefix_sum(c, (const uint32_t *)F.E.p, n_rows, (uint64_t *)P.out_off.p));
        CU(cudaMemcpyAsync(&gtotal_entries, (const uint64_t *)P.out_off.p + n_rows, 8, cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaMemcpyAsync(&gtotal, (const uint64_t *)F.gofs.p + n_rows, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(&rtotal, (const uint64_t *)P.rofs.p + n_rows, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (gtotal) {
        RC(F.gscratch.ensure(gtotal * 2 * 4));
        CU(cudaMemsetAsync(F.gscratch.p, 0, gtotal * 2 * 4, c->stream));
    }
    RC(P.entries.ensure((rtotal + 1) * sizeof(ckm_pair_t)));
    CU(cudaFuncSetAttribute(matrix_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMatSmem));
    const unsigned mb = (unsigned)std::min<uint64_t>(((uint64_t)n_rows + kMatWarps - 1) / kMatWarps, (uint64_t)c->sm_count * 8);
    matrix_row_kernel<<<mb, kMatWarps * 32, kMatSmem, c->stream>>>(
        (const uint32_t *)P.ids.p, (const uint64_t *)c->in_off.p, (const uint32_t *)c->n_hits.p, (const uint2 *)F.hit_fam.p,
        (const uint32_t *)F.E.p, (const uint32_t *)F.gcap.p, (const uint64_t *)F.gofs.p, (uint32_t *)F.gscratch.p,
        (const uint32_t *)P.d_eids.p, (const uint32_t *)P.d_first.p, max_eid, n_rows, row_begin, (const uint64_t *)P.rofs.p,
        (ckm_pair_t *)P.entries.p, (uint32_t *)P.nd.p, filter);
    c->launches++;
    RC(prefix_sum(c, (const uint32_t *)P.nd.p, n_rows, (uint64_t *)P.out_off.p));
    uint64_t np = 0;
    CU(cudaMemcpyAsync(&np, (const uint64_t *)P.out_off.p + n_rows, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    RC(P.out.ensure((np + 1) * sizeof(ckm_pair_t)));
    RC(P.h_out.ensure((np + 1) * sizeof(ckm_pair_t)));
    matrix_export_kernel<<<(unsigned)(((uint64_t)n_rows * 32 + 255) / 256), 256, 0, c->stream>>>(
        (const uint64_t *)P.rofs.p, (const ckm_pair_t *)P.entries.p, (const uint64_t *)P.out_off.p, n_rows, (ckm_pair_t *)P.out.p,
        !filter || c->matrix_tile_on_device);
    c->launches++;
    c->matrix_walked = gtotal_entries;
    if (np && !c->matrix_tile_on_device) CU(cudaMemcpyAsync(P.h_out.p, P.out.p, np * sizeof(ckm_pair_t), cudaMemcpyDeviceToHost, c->stream));
    if (h_off) {
        RC(P.h_off.ensure(((size_t)n_rows + 2) * 8));
        CU(cudaMemcpyAsync(P.h_off.p, P.out_off.p, ((size_t)n_rows + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        *h_off = (const uint64_t *)P.h_off.p;
    }
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    *pairs = (const ckm_pair_t *)P.h_out.p;
    *n_pairs = np;
    return 0;
}

// ---- multi-GPU /matrix: hit extraction sharded by protein block, postings exchanged device to device ----
// The (k-mer, peg id) pairs this ctx holds, in HBM: what a rank contributes to the all-gather of the postings.
extern "C" int ckm_postings_device(ckm_ctx *c, const uint64_t **d_keys, const uint32_t **d_eids, uint64_t *n) {
    if (!c || !d_keys || !d_eids || !n) return ckm_fail(CKM_EINVAL, "NULL argument");
    *d_keys = (const uint64_t *)c->post.keys.p;
    *d_eids = (const uint32_t *)c->post.eids.p;
    *n = c->post.n;
    return 0;
}
// Replace the ctx's postings by `n` pairs that already sit in HBM (the gathered postings of all ranks); device-to-device copy.
extern "C" int ckm_postings_import_device(ckm_ctx *c, const uint64_t *d_keys, const uint32_t *d_eids, uint64_t n) {
    if (!c || (n && (!d_keys || !d_eids))) return ckm_fail(CKM_EINVAL, "NULL argument");
    CU(cudaSetDevice(c->device));
    ckm_ctx::Post &P = c->post;
    RC(P.keys.ensure((n + 1) * 8));
    RC(P.eids.ensure((n + 1) * 4));
    if (n) {
        CU(cudaMemcpyAsync(P.keys.p, d_keys, n * 8, cudaMemcpyDeviceToDevice, c->stream));
        CU(cudaMemcpyAsync(P.eids.p, d_eids, n * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    P.n = n;
    P.dirty = true;
    return 0;
}
// ckm_matrix_rows that leaves its COO tile in HBM (valid until the next call on the ctx), every row's entries ordered by partner
// id: with ids that ascend in request order the tile is then already in the (eid_i, eid_j) order of the reference's std::map.
// *postings_walked = posting-list entries the rows' hits carry (the row block's work).
extern "C" int ckm_matrix_rows_device(ckm_ctx *c, const uint32_t *eids, const char *residues, const uint64_t *offsets, uint32_t n,
                                      uint32_t row_begin, uint32_t row_end, const ckm_pair_t **d_pairs, uint64_t *n_pairs,
                                      uint64_t *postings_walked) {
    if (!c || !d_pairs || !n_pairs || (n && (!eids || !offsets))) return ckm_fail(CKM_EINVAL, "NULL argument");
    *d_pairs = nullptr;
    *n_pairs = 0;
    if (postings_walked) *postings_walked = 0;
    if (row_end > n) row_end = n;
    if (row_begin >= row_end) return 0;
    const ckm_pair_t *unused = nullptr;
    c->matrix_tile_on_device = true;
    const int rc = matrix_rows_impl(c, eids, residues, offsets, n, row_begin, row_end, true, &unused, n_pairs, nullptr);
    c->matrix_tile_on_device = false;
    if (rc) return rc;
    *d_pairs = (const ckm_pair_t *)c->post.out.p;
    if (postings_walked) *postings_walked = c->matrix_walked;
    return 0;
}

extern "C" int ckm_matrix_rows(ckm_ctx *c, const uint32_t *eids, const char *residues, const uint64_t *offsets, uint32_t n,
                               uint32_t row_begin, uint32_t row_end, const ckm_pair_t **pairs, uint64_t *n_pairs) {
    if (!c || !pairs || !n_pairs || (n && (!eids || !offsets))) return ckm_fail(CKM_EINVAL, "NULL argument");
    *pairs = nullptr;
    *n_pairs = 0;
    if (row_end > n) row_end = n;
    if (row_begin >= row_end) return 0;
    return matrix_rows_impl(c, eids, residues, offsets, n, row_begin, row_end, true, pairs, n_pairs, nullptr);
}

// LookupRequest::on_hit without families (lookup_request.cc:466-478): seq_score_[eid].hit_count++ for every posting of
// every hit k-mer.  pairs[k] = {sequence index, peg id, hit_count}, CSR by sequence, ascending peg id.
extern "C" int ckm_postings_scores(ckm_ctx *c, const char *residues, const uint64_t *offsets, uint32_t n, const ckm_pair_t **pairs,
                                   const uint64_t **pair_offsets) {
    if (!c || !pairs || !pair_offsets || !offsets) return ckm_fail(CKM_EINVAL, "NULL argument");
    static const uint64_t zero = 0;
    *pairs = nullptr;
    *pair_offsets = &zero;
    if (n == 0) return 0;
    uint64_t np = 0;
    return matrix_rows_impl(c, nullptr, residues, offsets, n, 0, n, false, pairs, &np, pair_offsets);
}

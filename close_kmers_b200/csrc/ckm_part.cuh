// Partitioned probe path: the same results as probe_kernel + scan_kernel, with the table probes re-ordered so that
// they stream through L2 instead of being independent random HBM transactions.
//
// Why: independent random 16-byte reads over a table larger than L2 top out at ~37 G/s on a B200 whatever the load
// flavour (tools/gather_micro.cu): every probe costs one 128-byte DRAM line fetch.  With a big batch
// the probes are dense in the table (C2: one per 28 bytes, C3: one per 3.4 bytes), so visiting the table region by
// region turns them into L2 hits: the same microbenchmark reads 160 G probes/s when the probe stream is sorted into
// 32 MB table bins.  The price is two radix-partition passes (probe records into table bins, hit records back into
// position order), all of them streaming.
//
//   part_count_kernel    one warp per protein (tile_keys, same encode as probe_kernel): histogram of probes per table bin
//   part_scatter_kernel  same walk; 8-byte records (key | origin << 35) written bin by bin through a shared-memory
//                        write-combining stage (128-byte lines)
//   part_probe_kernel    the record stream in bin order; each record probes its slot (L2-resident bin); hits leave as
//                        16-byte records staged into position bins (origin >> dshift)
//   part_place_kernel    position bins in order: payload -> dense8[origin], bit -> valid[origin] (window stays in L2)
//   scan_kernel<.,DENSE> the ordered scoring scan reads the position-indexed hits
//
// Used for calls / best-call batches (no hit export, no OTU stats, order_constraint off, no protein long enough to
// saturate the 39 998-hit window) on a packed table, when the batch is dense enough in the table to pay off.
#pragma once

namespace ckm {

constexpr int kPartThreads = 256;
constexpr uint32_t kOriginBits = 29;     // residues per partitioned batch < 2^29 (key takes the low 35 bits of a record)
constexpr uint64_t kKeyMask = (1ull << 35) - 1;
constexpr int kMaxTableBins = 512;
constexpr int kMaxDestBins = 1024;
constexpr uint32_t kPlaceWindow = 1u << 19;  // positions per position bin: its validity bitmap (64 KB) lives in shared memory

struct PartGeom {
    uint32_t tshift, n_tbins;  // table bin = slot >> tshift
    uint32_t dshift, n_dbins;  // position bin = origin >> dshift
};

struct __align__(16) PartHit {
    uint32_t origin;  // global residue index of the k-mer's first residue
    uint32_t fI;
    uint32_t wt_bits;
    uint32_t pad;
};

// Block-level radix partition of one tile of records, without staging the records themselves:
//   1. every thread takes a rank inside its record's bin (shared-memory atomic on the tile histogram),
//   2. one thread per non-empty bin reserves that many slots of the bin's global region (one global atomic per bin
//      per tile) and clears the histogram,
//   3. every thread writes its record to region base + reservation + rank.
// Records of one bin written by one tile are contiguous (tens of bytes); the scattered 8/16-byte stores merge in L2
// before they are written back.  tile_reserve() contains the two block barriers.
__device__ __forceinline__ void tile_reserve(uint32_t n_bins, uint32_t *s_hist, uint32_t *s_res, unsigned long long *__restrict__ cursor) {
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < n_bins; b += blockDim.x) {
        const uint32_t c = s_hist[b];
        if (c) {
            s_hist[b] = 0;
            s_res[b] = (uint32_t)atomicAdd(&cursor[b], (unsigned long long)c);
        }
    }
    __syncthreads();
}

// SCATTER = false: count probes per table bin; true: write the probe records (key | origin << 35) bin by bin
template <bool SCATTER>
__global__ void __launch_bounds__(kPartThreads)
part_encode_kernel(TableView tv, PartGeom pg, const uint8_t *__restrict__ residues, const uint64_t *__restrict__ offsets, uint32_t n,
                   uint32_t *__restrict__ hist, uint64_t *__restrict__ recs, const uint64_t *__restrict__ tbase,
                   unsigned long long *__restrict__ tcursor, unsigned long long *__restrict__ totals) {
    __shared__ uint32_t s_hist[kMaxTableBins];
    __shared__ uint32_t s_res[kMaxTableBins];
    __shared__ uint8_t lut[256];
    for (uint32_t b = threadIdx.x; b < pg.n_tbins; b += blockDim.x) s_hist[b] = 0;
    fill_aa_lut(lut);
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // this warp's current protein
    uint32_t my_probes = 0;
    // per-protein walking state (warp-uniform)
    uint64_t base = 0;
    uint32_t len = 0, nwin = 0, t0 = 0, nwords = 0, sh = 0;
    const uint32_t *wb = nullptr;
    bool open = false;
    for (;;) {
        // advance to the next warp step that has windows
        while (!open && i < n) {
            base = __ldg(offsets + i);
            len = (uint32_t)(__ldg(offsets + i + 1) - base);
            if (len > CKM_KMER_SIZE) {
                nwin = len - CKM_KMER_SIZE;
                const uint8_t *p0 = residues + base;
                const uint32_t s = (uint32_t)(reinterpret_cast<uintptr_t>(p0) & 3u);
                wb = reinterpret_cast<const uint32_t *>(p0 - s);
                nwords = (len + s + 3u) >> 2;
                sh = 8u * s;
                t0 = 0;
                open = true;
            } else {
                i += n_warps;
            }
        }
        if (SCATTER) {
            if (!__syncthreads_or(open ? 1 : 0)) break;  // block-synchronous tiles: all warps step together
        } else if (!open) {
            break;
        }
        TileKeys tk;
        tk.act = 0;
        uint64_t origin0 = 0;
        if (open) {
            tk = tile_keys(lut, wb, nwords, sh, t0, lane, len, nwin);
            origin0 = base + t0 + 4u * lane;
            t0 += kTile;
            if (t0 >= nwin) {  // nwin may have shrunk (embedded NUL)
                open = false;
                i += n_warps;
            }
        }
        uint32_t bin[4], rank[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            bin[j] = (uint32_t)(fast_mod(tk.key[j], tv.num_sigs, tv.magic) >> pg.tshift);
            if (tk.act & (1u << j)) rank[j] = atomicAdd(&s_hist[bin[j]], 1u);
        }
        my_probes += __popc(tk.act);
        if (SCATTER) {
            tile_reserve(pg.n_tbins, s_hist, s_res, tcursor);
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (tk.act & (1u << j)) recs[tbase[bin[j]] + s_res[bin[j]] + rank[j]] = tk.key[j] | ((origin0 + j) << 35);
        }
    }
    if (!SCATTER) {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < pg.n_tbins; b += blockDim.x)
            if (s_hist[b]) atomicAdd(&hist[b], s_hist[b]);
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) my_probes += __shfl_down_sync(0xffffffffu, my_probes, d);
        if (lane == 0 && my_probes) atomicAdd(totals + 0, (unsigned long long)my_probes);
    }
}

constexpr int kProbeRecs = 4;  // records per thread per chunk: four independent slot loads in flight
constexpr uint32_t kProbeChunk = kPartThreads * kProbeRecs;

// the record stream, in table-bin order: persistent blocks take chunks from a global counter, so the grid sweeps the
// bins together and the slots it touches stay within an L2-sized window; hits leave partitioned by position bin
__global__ void __launch_bounds__(kPartThreads)
part_probe_kernel(TableView tv, PartGeom pg, const uint64_t *__restrict__ recs, uint64_t n_recs, unsigned long long *next_chunk,
                  PartHit *__restrict__ hits_out, unsigned long long *__restrict__ dcursor, unsigned long long *__restrict__ totals) {
    __shared__ uint32_t s_hist[kMaxDestBins];
    __shared__ uint32_t s_res[kMaxDestBins];
    __shared__ unsigned long long s_chunk;
    for (uint32_t b = threadIdx.x; b < pg.n_dbins; b += blockDim.x) s_hist[b] = 0;
    uint32_t my_hits = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(next_chunk, 1ull);
        __syncthreads();
        const uint64_t start = s_chunk * kProbeChunk;
        if (start >= n_recs) break;
        uint64_t key[kProbeRecs], h[kProbeRecs];
        uint32_t origin[kProbeRecs];
        uint4 v[kProbeRecs];
        uint32_t act = 0;
#pragma unroll
        for (int u = 0; u < kProbeRecs; u++) {
            const uint64_t r = start + (uint64_t)u * kPartThreads + threadIdx.x;
            if (r < n_recs) {
                const uint64_t rec = __ldg(recs + r);
                key[u] = rec & kKeyMask;
                origin[u] = (uint32_t)(rec >> 35);
                h[u] = fast_mod(key[u], tv.num_sigs, tv.magic);
                act |= 1u << u;
            }
        }
#pragma unroll
        for (int u = 0; u < kProbeRecs; u++)
            if (act & (1u << u)) v[u] = SlotIO<true>::load(tv.slots, h[u]);
        uint32_t hm = 0, rank[kProbeRecs];
        SlotFields f[kProbeRecs];
#pragma unroll
        for (int u = 0; u < kProbeRecs; u++) {
            if (act & (1u << u)) {
                int r = SlotIO<true>::test(v[u], key[u], f[u]);
                uint64_t guard = 0;
                while (r == 0) {  // linear probing, kguts.cc:589
                    h[u] = (h[u] + 1 == tv.num_sigs) ? 0 : h[u] + 1;
                    if (++guard >= tv.num_sigs) { r = -1; break; }
                    v[u] = SlotIO<true>::load(tv.slots, h[u]);
                    r = SlotIO<true>::test(v[u], key[u], f[u]);
                }
                if (r > 0) {
                    hm |= 1u << u;
                    rank[u] = atomicAdd(&s_hist[origin[u] >> pg.dshift], 1u);
                }
            }
        }
        my_hits += __popc(hm);
        tile_reserve(pg.n_dbins, s_hist, s_res, dcursor);
#pragma unroll
        for (int u = 0; u < kProbeRecs; u++) {
            if (hm & (1u << u)) {
                const uint32_t db = origin[u] >> pg.dshift;
                PartHit ph;
                ph.origin = origin[u];
                ph.fI = f[u].fI;
                ph.wt_bits = __float_as_uint(f[u].wt);
                ph.pad = 0;
                hits_out[((uint64_t)db << pg.dshift) + s_res[db] + rank[u]] = ph;
            }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) my_hits += __shfl_down_sync(0xffffffffu, my_hits, d);
    if ((threadIdx.x & 31u) == 0 && my_hits) atomicAdd(totals + 1, (unsigned long long)my_hits);
}

// One block per position bin: its hit payloads land at their residue index (dense8) and its validity bits are
// collected in shared memory and written out once -- no global atomics, no global clearing pass.
constexpr int kPlaceThreads = 512;
__global__ void __launch_bounds__(kPlaceThreads)
part_place_kernel(PartGeom pg, const PartHit *__restrict__ hits_in, const unsigned long long *__restrict__ dcursor,
                  uint2 *__restrict__ dense8, uint32_t *__restrict__ valid, uint64_t total) {
    extern __shared__ uint32_t s_bits[];  // (1 << dshift) / 32 words
    const uint32_t bin = blockIdx.x;
    const uint32_t words = (1u << pg.dshift) >> 5;
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) s_bits[w] = 0u;
    __syncthreads();
    const uint64_t cnt = dcursor[bin];
    const PartHit *src = hits_in + ((uint64_t)bin << pg.dshift);
    const uint32_t mask = (1u << pg.dshift) - 1u;
    for (uint64_t k0 = 0; k0 < cnt; k0 += 4ull * kPlaceThreads) {
        PartHit ph[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t k = k0 + (uint64_t)u * kPlaceThreads + threadIdx.x;
            if (k < cnt) ph[u] = src[k];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t k = k0 + (uint64_t)u * kPlaceThreads + threadIdx.x;
            if (k < cnt) {
                dense8[ph[u].origin] = make_uint2(ph[u].fI, ph[u].wt_bits);
                atomicOr(&s_bits[(ph[u].origin & mask) >> 5], 1u << (ph[u].origin & 31u));
            }
        }
    }
    __syncthreads();
    const uint64_t w0 = ((uint64_t)bin << pg.dshift) >> 5, wmax = (total >> 5) + 1;
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x)
        if (w0 + w <= wmax) valid[w0 + w] = s_bits[w];
}

}  // namespace ckm

// ---------------------------------------------------------------------------------------------------
// host side (included by ckm_api.cu after RunPlan / prefix_sum)
// ---------------------------------------------------------------------------------------------------
static bool part_geometry(const ckm_ctx *c, uint64_t total, ckm::PartGeom *pg) {
    using namespace ckm;
    uint32_t tshift = c->part_tshift ? c->part_tshift : 21;  // 2^21 slots x 16 B = 32 MB of table per bin
    while ((((c->num_sigs - 1) >> tshift) + 1) > (uint64_t)kMaxTableBins) tshift++;
    uint32_t dshift = c->part_dshift ? c->part_dshift : 19;  // 2^19 positions: 64 KB of validity bits per block, 4 MB of dense8
    while (((total >> dshift) + 1) > (uint64_t)kMaxDestBins) dshift++;
    if (dshift > 19) return false;  // the placement kernel keeps a bin's bitmap in shared memory
    pg->tshift = tshift;
    pg->n_tbins = (uint32_t)(((c->num_sigs - 1) >> tshift) + 1);
    pg->dshift = dshift;
    pg->n_dbins = (uint32_t)((total >> dshift) + 1);
    return total < (1ull << kOriginBits);
}

// is the partitioned path applicable (and, unless forced, worthwhile) for this batch?
static bool use_partitioned(const ckm_ctx *c, uint32_t n, uint64_t total, uint32_t max_len, uint32_t flags) {
    if (c->part_mode == 0 || n == 0) return false;
    if (c->slot_bytes != ckm::kPackedSlotBytes) return false;
    if (flags == 0 || (flags & ~(CKM_WANT_BEST | CKM_WANT_CALLS))) return false;  // no hit export / OTU stats
    if (c->prm.order_constraint != 0 || max_len == 0 || max_len > ckm::kHitCap + CKM_KMER_SIZE) return false;
    if (total >= (1ull << ckm::kOriginBits)) return false;
    if (c->part_mode == 1) return true;  // forced (tests)
    // auto: the table must not fit L2, and the batch must be dense in it (at least one probe per 64 B of table),
    // otherwise each table line is touched about once anyway and the two extra passes are pure overhead
    const uint64_t table_bytes = c->num_sigs * (uint64_t)c->slot_bytes;
    return table_bytes > (uint64_t)c->l2_bytes && total * 64 >= table_bytes;
}

static int run_partitioned(ckm_ctx *c, const uint8_t *d_res, const uint64_t *d_off, uint32_t n, uint64_t total, uint32_t flags,
                           const RunPlan &plan) {
    using namespace ckm;
    ckm_ctx::Part &P = c->part;
    PartGeom pg;
    if (!part_geometry(c, total, &pg)) return ckm_fail(CKM_EINVAL, "batch too large for the partitioned path");
    cudaStream_t st = c->stream;
    TableView tv;
    tv.slots = c->table.p;
    tv.num_sigs = c->num_sigs;
    tv.magic = c->magic;
    tv.occupied = nullptr;
    tv.tuning = 0;
    RC(P.hist.ensure(((size_t)pg.n_tbins + 1) * 4));
    RC(P.tbase.ensure(((size_t)pg.n_tbins + 2) * 8));
    RC(P.tcursor.ensure(((size_t)pg.n_tbins + 1) * 8));
    RC(P.dcursor.ensure(((size_t)pg.n_dbins + 1) * 8));
    RC(P.counter.ensure(64));
    RC(P.dense8.ensure((total + 64) * 8));
    RC(P.valid.ensure((((size_t)pg.n_dbins << pg.dshift) / 32 + 4) * 4));
    RC(c->hits.ensure(((size_t)pg.n_dbins << pg.dshift) * sizeof(PartHit)));  // position-bin regions reuse the hit buffer
    CU(cudaMemsetAsync(P.hist.p, 0, ((size_t)pg.n_tbins + 1) * 4, st));
    CU(cudaMemsetAsync(P.tcursor.p, 0, ((size_t)pg.n_tbins + 1) * 8, st));
    CU(cudaMemsetAsync(P.dcursor.p, 0, ((size_t)pg.n_dbins + 1) * 8, st));
    CU(cudaMemsetAsync(P.counter.p, 0, 64, st));

    ckm_ctx::ProfEv pe;
    memset(&pe, 0, sizeof pe);
    const bool prof = c->profiling;
    if (prof) {
        for (int k = 0; k < 6; k++) CU(cudaEventCreate(&pe.ev[k]));
        CU(cudaEventRecord(pe.ev[0], st));
    }
    const unsigned eblocks = (unsigned)std::min<uint64_t>(((uint64_t)n + 7) / 8, (uint64_t)c->sm_count * 16);
    // 1. probes per table bin
    part_encode_kernel<false><<<eblocks, kPartThreads, 0, st>>>(tv, pg, d_res, d_off, n, (uint32_t *)P.hist.p, nullptr, nullptr,
                                                                       nullptr, (unsigned long long *)c->totals.p);
    c->launches++;
    RC(prefix_sum(c, (const uint32_t *)P.hist.p, pg.n_tbins, (uint64_t *)P.tbase.p));
    uint64_t n_recs = 0;
    CU(cudaMemcpyAsync(&n_recs, (const uint64_t *)P.tbase.p + pg.n_tbins, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    RC(P.recs.ensure((n_recs + 64) * 8));
    if (prof) CU(cudaEventRecord(pe.ev[1], st));
    // 2. probe records, table bin by table bin
    part_encode_kernel<true><<<eblocks, kPartThreads, 0, st>>>(tv, pg, d_res, d_off, n, nullptr, (uint64_t *)P.recs.p,
                                                                      (const uint64_t *)P.tbase.p, (unsigned long long *)P.tcursor.p,
                                                                      (unsigned long long *)c->totals.p);
    c->launches++;
    if (prof) CU(cudaEventRecord(pe.ev[2], st));
    // 3. the probes themselves, in bin order; the grid is exactly what is resident so that it sweeps the bins together
    int per_sm = 1;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, part_probe_kernel, kPartThreads, 0));
    per_sm = std::max(1, std::min(per_sm, 4));
    part_probe_kernel<<<(unsigned)(c->sm_count * per_sm), kPartThreads, 0, st>>>(
        tv, pg, (const uint64_t *)P.recs.p, n_recs, (unsigned long long *)P.counter.p, (PartHit *)c->hits.p,
        (unsigned long long *)P.dcursor.p, (unsigned long long *)c->totals.p);
    c->launches++;
    if (prof) CU(cudaEventRecord(pe.ev[3], st));
    // 4. hits back to position order
    {
        const size_t bits_smem = ((size_t)1 << pg.dshift) / 8;
        CU(cudaFuncSetAttribute(part_place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bits_smem));
        part_place_kernel<<<pg.n_dbins, kPlaceThreads, bits_smem, st>>>(pg, (const PartHit *)c->hits.p,
                                                                        (const unsigned long long *)P.dcursor.p, (uint2 *)P.dense8.p,
                                                                        (uint32_t *)P.valid.p, total);
        c->launches++;
    }
    if (prof) CU(cudaEventRecord(pe.ev[4], st));
    // 5. ordered scoring scan (+ find_best_call) over the position-indexed hits
    {
        ScanArgs a;
        memset(&a, 0, sizeof a);
        a.offsets = d_off;
        a.calls = (ckm_call_t *)c->calls.p;
        a.calls_work = (flags & CKM_WANT_BEST) ? (ckm_call_t *)c->calls_work.p : nullptr;
        a.n_calls = (uint32_t *)c->n_calls.p;
        a.best = (flags & CKM_WANT_BEST) ? (ckm_best_t *)c->best.p : nullptr;
        a.totals = (unsigned long long *)c->totals.p;
        a.valid = (const uint32_t *)P.valid.p;
        a.dense8 = (const uint2 *)P.dense8.p;
        a.n = n;
        a.index_base = 0;
        a.prm = c->prm;
        scan_kernel<false, true><<<(n + kScanThreads - 1) / kScanThreads, kScanThreads, 0, st>>>(a);
        c->launches++;
    }
    if (prof) {
        CU(cudaEventRecord(pe.ev[5], st));
        pe.n_ev = 6;
        c->prof.push_back(pe);
    }
    (void)plan;
    CU(cudaGetLastError());
    return 0;
}

// The k-mer -> family table built on the GPU from the proteins of families.nr ("next" row N2).  Included after
// ckm_matrix.cuh (it reuses the postings append / index machinery) at the end of ckm_api.cu.
#pragma once

// ---------------------------------------------------------------------------------------------------
// Family tables from families.nr proteins: NRLoader::thread_load (nr_loader.cc:131-202) runs
// process_aa_seq on every protein of a family with a hit callback that queues (k-mer, family id);
// KmerInserter::thread_main (kmer_inserter.cc:36-58) applies KmerPegMapping::add_fam_mapping (kmer.cc:244-268), whose
// fam_map_insert (216-230) keeps each family id once per k-mer.  Here: the hit kernel over the batch, (k-mer, family)
// pairs appended like postings, duplicates removed through a device hash set, lists built by the postings indexer,
// and the result installed as the family table that fam_lookup_kernel / fam_vote_kernel read.
// ---------------------------------------------------------------------------------------------------
namespace ckm {
__global__ void __launch_bounds__(256)
pair_dedupe_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ fams, uint64_t n, unsigned long long *set,
                   uint64_t mask, uint64_t *__restrict__ out_keys, uint32_t *__restrict__ out_fams, unsigned long long *n_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long packed = (keys[i] | ((unsigned long long)fams[i] << 35)) + 1ull;  // k-mer < 2^35, family < 2^29
    uint64_t s = fam_hash(packed) & mask;
    for (;;) {
        const unsigned long long prev = atomicCAS(&set[s], 0ull, packed);
        if (prev == packed) return;  // seen before: fam_map_insert finds it and does nothing
        if (prev == 0ull) {
            const unsigned long long o = atomicAdd(n_out, 1ull);
            out_keys[o] = keys[i];
            out_fams[o] = fams[i];
            return;
        }
        s = (s + 1) & mask;
    }
}
}  // namespace ckm

// drop repeated (k-mer, family) pairs from the collected list; idempotent, so it can run whenever the list has grown
static int famnr_compact(ckm_ctx *c) {
    ckm_ctx::Post &P = c->famnr;
    uint64_t cap = 16;
    while (cap < 2 * P.n) cap <<= 1;
    DevBuf set, ukeys, ufams, cnt;
    RC(set.ensure(cap * 8));
    RC(ukeys.ensure((P.n + 1) * 8));
    RC(ufams.ensure((P.n + 1) * 4));
    RC(cnt.ensure(64));
    CU(cudaMemsetAsync(set.p, 0, cap * 8, c->stream));
    CU(cudaMemsetAsync(cnt.p, 0, 8, c->stream));
    if (P.n) {
        ckm::pair_dedupe_kernel<<<(unsigned)((P.n + 255) / 256), 256, 0, c->stream>>>(
            (const uint64_t *)P.keys.p, (const uint32_t *)P.eids.p, P.n, (unsigned long long *)set.p, cap - 1, (uint64_t *)ukeys.p,
            (uint32_t *)ufams.p, (unsigned long long *)cnt.p);
        c->launches++;
    }
    uint64_t n_unique = 0;
    CU(cudaMemcpyAsync(&n_unique, cnt.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaGetLastError());
    set.release();
    cnt.release();
    P.keys.release();
    P.eids.release();
    P.keys = ukeys;
    P.eids = ufams;
    P.n = n_unique;
    P.dirty = true;
    c->famnr_last_unique = n_unique;
    return 0;
}

extern "C" int ckm_family_nr_begin(ckm_ctx *c) {
    if (!c) return ckm_fail(CKM_EINVAL, "ctx is NULL");
    c->famnr.n = 0;
    c->famnr.dirty = true;
    c->famnr_last_unique = 0;
    if (const char *e = getenv("CKM_FAMNR_COMPACT_AT")) c->famnr_compact_at = strtoull(e, nullptr, 10);
    return 0;
}

extern "C" int ckm_family_nr_add(ckm_ctx *c, const uint32_t *fam_ids, const char *residues, const uint64_t *offsets, uint32_t n) {
    if (!c || (n && !fam_ids)) return ckm_fail(CKM_EINVAL, "NULL argument");
    for (uint32_t i = 0; i < n; i++)
        if (fam_ids[i] != 0xffffffffu && fam_ids[i] >= (1u << 29)) return ckm_fail(CKM_EINVAL, "family id %u does not fit 29 bits", fam_ids[i]);
    // "NO FAM FOR id" (nr_loader.cc:154-160): thread_load RETURNS there, so the first protein without a family ends the
    // chunk -- the sequences after it in the same call are not loaded either.  Kept as is.
    for (uint32_t i = 0; i < n; i++)
        if (fam_ids[i] == 0xffffffffu) {
            n = i;
            break;
        }
    if (n == 0) return 0;
    uint64_t total = 0;
    uint32_t max_len = 0;
    RC(upload_batch(c, residues, offsets, n, &total, &max_len));
    // process_aa_seq(id, seq, 0, hit_cb, 0): hits only (nr_loader.cc:172)
    RC(run_device(c, (const uint8_t *)c->in_res.p, (const uint64_t *)c->in_off.p, n, total, std::max(max_len, 1u), CKM_WANT_HITS));
    RC(postings_append_device(c, c->famnr, fam_ids, (const uint64_t *)c->in_off.p, n));
    // members of one family share most of their k-mers: keep the list near its distinct size instead of its raw size
    if (c->famnr.n >= std::max<uint64_t>(c->famnr_compact_at, 2 * c->famnr_last_unique)) RC(famnr_compact(c));
    return 0;
}

extern "C" int ckm_family_nr_finish(ckm_ctx *c, uint32_t n_families, const char *const *pgf, const char *const *plf,
                                    const char *const *function, uint64_t *n_kmers_out, uint64_t *n_entries_out) {
    if (!c) return ckm_fail(CKM_EINVAL, "ctx is NULL");
    if (c->shares_tables) return ckm_fail(CKM_ESTATE, "a clone reads its parent's family tables and cannot load its own");
    CU(cudaSetDevice(c->device));
    ckm_ctx::Post &P = c->famnr;
    // 1. (k-mer, family) pairs, each once
    RC(famnr_compact(c));
    const uint64_t n_unique = P.n;
    // 2. group by k-mer with the postings indexer (count -> prefix sum -> fill)
    RC(postings_index(c, P));
    // 3. install as the family table (same slot format), with the family metadata interned like ckm_family_load
    ckm_ctx::Family &F = c->fam;
    RC(family_install_metadata(c, n_families, pgf, plf, function));
    F.table.release();
    F.ids.release();
    F.table = P.slots;
    F.ids = P.ids;
    F.mask = P.mask;
    P.slots = DevBuf();
    P.ids = DevBuf();
    F.loaded = true;
    if (n_entries_out) *n_entries_out = n_unique;
    if (n_kmers_out) *n_kmers_out = P.n_keys;
    P.n = 0;
    DevBuf *work[] = {&P.keys, &P.eids, &P.tkeys, &P.tcnt, &P.tcur, &P.toff};
    for (auto b : work) b->release();
    return 0;
}

// copy the installed k-mer -> family lists back to the host as CSR (for inspection / tests); arrays sized by the caller
// from ckm_family_nr_finish's counts: kmers[n_kmers], offsets[n_kmers+1], ids[n_entries]; lists are unordered
extern "C" int ckm_family_export(ckm_ctx *c, uint64_t n_kmers, uint64_t n_entries, uint64_t *kmers, uint64_t *offsets, uint32_t *ids) {
    if (!c || !c->fam.loaded) return ckm_fail(CKM_ESTATE, "no family table loaded");
    CU(cudaSetDevice(c->device));
    const ckm_ctx::Family &F = c->fam;
    std::vector<ckm::FamSlot> slots((size_t)(F.mask + 1));
    CU(cudaMemcpy(slots.data(), F.table.p, slots.size() * sizeof(ckm::FamSlot), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> all(n_entries + 1);
    if (n_entries) CU(cudaMemcpy(all.data(), F.ids.p, n_entries * 4, cudaMemcpyDeviceToHost));
    uint64_t k = 0, e = 0;
    offsets[0] = 0;
    for (const auto &sl : slots) {
        if (!sl.key1) continue;
        if (k >= n_kmers || e + sl.cnt > n_entries) return ckm_fail(CKM_EINVAL, "export arrays too small");
        kmers[k] = sl.key1 - 1;
        memcpy(ids + e, all.data() + sl.off, (size_t)sl.cnt * 4);
        e += sl.cnt;
        offsets[++k] = e;
    }
    if (k != n_kmers) return ckm_fail(CKM_EINVAL, "table holds %llu k-mers, caller expected %llu", (unsigned long long)k, (unsigned long long)n_kmers);
    return 0;
}

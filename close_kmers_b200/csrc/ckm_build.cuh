// Image builder on the GPU ("next" row N4): write_hashtable (build_signature_kmers.cc:860-898) = KmerGuts(dir, nbuckets) +
// insert_kmer for every kept k-mer + save_kmer_hash_table (kguts.cc:77-115, 166-171, 188-234), and the per-signature
// weight (build_signature_kmers.cc:841-853).  Included at the end of ckm_api.cu.
//
// The reference inserts sequentially with first-come-first-served linear probing, so the FILE BYTES depend on insertion
// order.  They are reproduced in parallel with priority linear probing, priority = insertion index: an inserter that
// meets a slot held by a later key takes the slot and carries the evicted key onwards; one that meets an earlier key moves
// on.  The resulting table satisfies "every slot between a key's home and its position holds an earlier key", which the
// sequential table also satisfies and which has exactly one solution (the table is under half full, so the probe circle is
// cut by empty slots, and those are the same whatever the order) -- hence identical bytes, whatever the thread schedule.
#pragma once

namespace ckm {

constexpr uint32_t kBuildEmpty = 0xffffffffu;

__global__ void __launch_bounds__(256)
build_insert_kernel(const uint64_t *__restrict__ keys, uint32_t n, uint64_t nbuckets, uint64_t magic, uint32_t *owner) {
    uint32_t cur = blockIdx.x * blockDim.x + threadIdx.x;
    if (cur >= n) return;
    uint64_t key = keys[cur];
    if (key > CKM_MAX_ENCODED) return;  // kguts.cc:206-210: not inserted
    uint64_t s = fast_mod(key, nbuckets, magic);
    for (;;) {
        const uint32_t old = atomicMin(&owner[s], cur);
        if (old == kBuildEmpty) return;
        if (old > cur) {  // we were inserted first: the slot is ours, the evicted key continues from the next slot
            cur = old;
        }
        s = (s + 1 == nbuckets) ? 0 : s + 1;
    }
}

// sig_kmer_t (kmer_image.h:17-23) per bucket, byte for byte what save_kmer_hash_table writes: empty buckets carry
// MAX_ENCODED+1 and zeros (kguts.cc:93, 106-107)
__global__ void __launch_bounds__(256)
build_emit_kernel(const uint32_t *__restrict__ owner, uint64_t nbuckets, const uint64_t *__restrict__ keys, const int32_t *__restrict__ fI,
                  const int32_t *__restrict__ oI, const uint16_t *__restrict__ avg, const float *__restrict__ wt, uint64_t *__restrict__ raw) {
    const uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nbuckets) return;
    const uint32_t i = owner[s];
    uint64_t w0 = CKM_MAX_ENCODED + 1, w1 = 0, w2 = 0;
    if (i != kBuildEmpty) {
        w0 = keys[i];
        w1 = (uint64_t)(uint32_t)oI[i] | ((uint64_t)avg[i] << 32);                       // otu_index @8, avg_from_end @12, pad
        w2 = (uint64_t)(uint32_t)fI[i] | ((uint64_t)__float_as_uint(wt[i]) << 32);       // function_index @16, function_wt @20
    }
    raw[3 * s] = w0;
    raw[3 * s + 1] = w1;
    raw[3 * s + 2] = w2;
}

}  // namespace ckm

// builds the 24-byte slots of the image on the current device; *raw_out receives a device buffer of nbuckets slots
static int build_raw_device(cudaStream_t stream, uint64_t nbuckets, uint64_t n, const uint64_t *keys, const int32_t *fI, const int32_t *oI,
                            const uint16_t *avg, const float *wt, DevBuf *raw_out) {
    if (nbuckets == 0) return ckm_fail(CKM_EINVAL, "bucket count must be positive");
    if (n >= 0xffffffffull) return ckm_fail(CKM_EINVAL, "more than 2^32-2 k-mers");
    if (n && (!keys || !fI || !oI || !avg || !wt)) return ckm_fail(CKM_EINVAL, "NULL argument");
    uint64_t loaded = 0;
    for (uint64_t i = 0; i < n; i++) loaded += keys[i] <= CKM_MAX_ENCODED;
    if (loaded && (long long)loaded >= (long long)nbuckets / 2)  // kguts.cc:213-216 (the reference exits)
        return ckm_fail(CKM_EINVAL, "Your Kmer hash is half-full; use a larger bucket count");
    DevBuf d_keys, d_fI, d_oI, d_avg, d_wt, owner;
    int rc = 0;
    auto up = [&](DevBuf &b, const void *src, size_t bytes) {
        if (rc) return;
        rc = b.ensure(bytes + 16);
        if (!rc && bytes && cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess)
            rc = ckm_fail(CKM_ECUDA, "upload of the k-mer attributes failed");
    };
    up(d_keys, keys, n * 8);
    up(d_fI, fI, n * 4);
    up(d_oI, oI, n * 4);
    up(d_avg, avg, n * 2);
    up(d_wt, wt, n * 4);
    if (!rc) rc = owner.ensure(nbuckets * 4 + 16);
    if (!rc) rc = raw_out->ensure(nbuckets * kRawSlotBytes + 64);
    if (!rc && cudaMemsetAsync(owner.p, 0xff, nbuckets * 4, stream) != cudaSuccess) rc = ckm_fail(CKM_ECUDA, "memset failed");
    if (!rc) {
        const uint64_t magic = (uint64_t)((((unsigned __int128)1) << 64) / nbuckets);
        if (n)
            ckm::build_insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((const uint64_t *)d_keys.p, (uint32_t)n, nbuckets, magic,
                                                                                    (uint32_t *)owner.p);
        ckm::build_emit_kernel<<<(unsigned)((nbuckets + 255) / 256), 256, 0, stream>>>(
            (const uint32_t *)owner.p, nbuckets, (const uint64_t *)d_keys.p, (const int32_t *)d_fI.p, (const int32_t *)d_oI.p,
            (const uint16_t *)d_avg.p, (const float *)d_wt.p, (uint64_t *)raw_out->p);
        if (cudaStreamSynchronize(stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = ckm_fail(CKM_ECUDA, "image build kernels failed");
    }
    DevBuf *tmp[] = {&d_keys, &d_fI, &d_oI, &d_avg, &d_wt, &owner};
    for (auto b : tmp) b->release();
    if (rc) raw_out->release();
    return rc;
}

extern "C" int ckm_image_build_device(int device, uint64_t nbuckets, uint64_t n, const uint64_t *keys, const int32_t *fI,
                                      const int32_t *oI, const uint16_t *avg, const float *wt, void *image_out, size_t image_bytes) {
    const size_t need = sizeof(ckm_image_header_t) + (size_t)nbuckets * sizeof(ckm_sig_kmer_t);
    if (!image_out || image_bytes != need) return ckm_fail(CKM_EINVAL, "image buffer must be exactly %zu bytes", need);
    RC(select_device(device));
    cudaStream_t stream;
    CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    DevBuf raw;
    int rc = build_raw_device(stream, nbuckets, n, keys, fI, oI, avg, wt, &raw);
    if (!rc) {
        ckm_image_header_t *h = (ckm_image_header_t *)image_out;
        h->num_sigs = nbuckets;
        h->entry_size = sizeof(ckm_sig_kmer_t);
        h->version = 1;
        const size_t chunk = (size_t)1 << 28, bytes = (size_t)nbuckets * kRawSlotBytes;
        for (size_t o = 0; o < bytes && !rc; o += chunk)
            if (cudaMemcpyAsync((uint8_t *)(h + 1) + o, (uint8_t *)raw.p + o, std::min(chunk, bytes - o), cudaMemcpyDeviceToHost, stream) != cudaSuccess)
                rc = ckm_fail(CKM_ECUDA, "download of the image failed");
        if (!rc && cudaStreamSynchronize(stream) != cudaSuccess) rc = ckm_fail(CKM_ECUDA, "download of the image failed");
    }
    raw.release();
    cudaStreamDestroy(stream);
    return rc;
}

// build the table on the device and open a context on it without materialising the file
extern "C" int ckm_open_built(int device, uint64_t nbuckets, uint64_t n, const uint64_t *keys, const int32_t *fI, const int32_t *oI,
                              const uint16_t *avg, const float *wt, const char *const *function_names, int32_t n_functions,
                              const char *const *otu_names, int32_t n_otus, ckm_ctx **out) {
    if (!out) return ckm_fail(CKM_EINVAL, "out is NULL");
    *out = nullptr;
    ckm_ctx *c = nullptr;
    RC(ctx_create(device, &c));
    DevBuf raw;
    int rc = build_raw_device(c->stream, nbuckets, n, keys, fI, oI, avg, wt, &raw);
    if (!rc) rc = install_table(c, raw, nbuckets);
    if (rc) {
        ckm_close(c);
        return rc;
    }
    for (int32_t i = 0; i < n_functions; i++) c->functions.emplace_back(function_names[i]);
    for (int32_t i = 0; i < n_otus; i++) c->otu_names.emplace_back(otu_names[i]);
    *out = c;
    return 0;
}

// compute_weight_of_signature, build_signature_kmers.cc:841-853.  The counts are held as floats.  The first quotient is
// double arithmetic (the 1.0 literals promote); the second is FLOAT arithmetic, and its logarithm is the double function
// all the same: the reference includes <cmath> only, under which an unqualified log(float) is ::log(double) (checked with
// the reference's compiler and headers; oracle/ref_driver.cc static_asserts it for its restatement).  Spelled out here
// because this translation unit sees CUDA's global float overloads.  Host code, the same libm: the same bits.
extern "C" float ckm_signature_weight(float NSF, float KS, float NSi, float NFj, float NSiFj) {
    const float second = (NSF - NFj + KS) / (NFj + KS);
    return (float)(::log((double)((NSiFj + 1.0) / (NSi - NSiFj + 1.0))) + ::log((double)second));
}

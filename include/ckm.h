/*
 * ckm.h -- C ABI of the B200-native signature-k-mer calling path (libckm.so).
 *
 * This is the drop-in boundary for the hot path of olsonanl/close_kmers: the calls that the
 * reference's request handlers make, one sequence at a time, into a per-thread KmerGuts
 * (query_request.cc:103-152, add_request.cc:116-170, matrix_request.cc:82-94,
 * fq_process_request.cc:298-365, family_mapper.cc:46-205) are replaced by ONE batch call per body
 * chunk.  Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * Every entry point returns 0 on success or a negative CKM_E* code; ckm_last_error() gives the text.
 * The library never calls exit() (the reference does: kmer_image.cc:46-104, kguts.cc:550-553) and has
 * NO CPU fallback: without a usable sm_100 device every compute entry point fails with CKM_ECUDA.
 *
 * Threading: one ckm_ctx per GPU per host thread (the reference keeps one KmerGuts per worker thread,
 * threadpool.h:42).  Calls on one ctx are serialised on the ctx's CUDA stream.  Result pointers
 * returned through ckm_batch_out_t & co. point into ctx-owned pinned host memory and stay valid until
 * the next compute call on the same ctx.
 */
#ifndef CKM_H
#define CKM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* kmer_params.h:5,12,18,20 */
#define CKM_KMER_SIZE 8
#define CKM_CORE 1280000000ULL          /* 20^7 */
#define CKM_MAX_ENCODED 25600000000ULL  /* 20^8; which_kmer > MAX_ENCODED marks an empty slot */
#define CKM_MAX_HITS_PER_SEQ 40000

enum {
    CKM_OK = 0,
    CKM_EINVAL = -1,  /* bad argument */
    CKM_EIO = -2,     /* file missing / unreadable */
    CKM_EFORMAT = -3, /* image fails the reference's validation (kmer_image.cc:87-105) or index file not dense */
    CKM_ECUDA = -4,   /* CUDA error or no device; there is no CPU fallback */
    CKM_ENOMEM = -5,
    CKM_ESTATE = -6   /* call order violated (e.g. family batch before ckm_family_load) */
};

typedef struct ckm_ctx ckm_ctx;

/* ---- on-disk / in-memory image format, read unchanged (kmer_image.h:11-23) ---------------------- */
typedef struct {
    uint64_t num_sigs;   /* number of hash buckets */
    uint64_t entry_size; /* must be 24 */
    int64_t version;     /* must be 1 */
} ckm_image_header_t;

typedef struct {
    uint64_t which_kmer; /* > CKM_MAX_ENCODED => empty */
    int32_t otu_index;
    uint16_t avg_from_end;
    uint16_t pad_;
    int32_t function_index;
    float function_wt;
} ckm_sig_kmer_t; /* 24 bytes == sizeof(sig_kmer_t) */

/* ---- result records ------------------------------------------------------------------------------ */
/* KmerCall, kguts.h:166-183 */
typedef struct {
    uint32_t start;
    uint32_t end;
    int32_t count;
    uint32_t function_index;
    float weighted_hits;
} ckm_call_t; /* 20 bytes */

/* hit_in_sequence_t, kguts.h:228-233: a copy of the table slot plus the k-mer's offset in the protein */
typedef struct {
    uint64_t which_kmer;
    uint32_t offset;
    int32_t otu_index;
    int32_t function_index;
    float function_wt;
    uint16_t avg_from_end;
    uint16_t pad_;
    uint32_t pad2_;
} ckm_hit_t; /* 32 bytes */

/* one entry of KmerOtuStats::otu_map (kguts.h:191), emitted in ascending otu_index (std::map order) */
typedef struct {
    int32_t otu_index;
    int32_t count;
} ckm_otu_t;

/* outputs of KmerGuts::find_best_call, kguts.cc:1008-1199, with names left as indices */
#define CKM_BEST_HAS_CALLS 1u /* calls.size() > 0, i.e. score_offset was assigned (kguts.cc:1015-1018) */
#define CKM_BEST_AMBIG 2u     /* function is "F1 ?? F2" built from function_index_a / _b (kguts.cc:1176-1196) */
typedef struct {
    int32_t function_index; /* -1 when no confident call */
    int32_t ambig_a;        /* vec[0].first when CKM_BEST_AMBIG, else -1 */
    int32_t ambig_b;        /* vec[1].first when CKM_BEST_AMBIG, else -1 */
    uint32_t flags;
    float score;
    float weighted_score;
    float score_offset; /* 0 when !CKM_BEST_HAS_CALLS (the reference leaves the caller's variable untouched) */
} ckm_best_t;     /* 28 bytes */

/* what to compute / copy back; mirrors which arguments the handlers pass as non-null to process_aa_seq */
#define CKM_WANT_CALLS 1u /* vector<KmerCall> */
#define CKM_WANT_HITS 2u  /* hit callback list (details=1, /add, /matrix) */
#define CKM_WANT_OTU 4u   /* KmerOtuStats */
#define CKM_WANT_BEST 8u  /* find_best_call on the calls */

typedef struct {
    uint32_t n;                   /* sequences in the batch */
    const uint64_t *call_offsets; /* n+1, CSR into calls; NULL unless CKM_WANT_CALLS */
    const ckm_call_t *calls;
    const uint64_t *hit_offsets; /* n+1; NULL unless CKM_WANT_HITS */
    const ckm_hit_t *hits;
    const uint64_t *otu_offsets; /* n+1; NULL unless CKM_WANT_OTU */
    const ckm_otu_t *otus;
    const ckm_best_t *best; /* n; NULL unless CKM_WANT_BEST */
    uint64_t n_probes;      /* table probes issued for this batch (valid windows) */
    uint64_t n_hits;        /* table hits */
} ckm_batch_out_t;

/* ---- lifecycle ------------------------------------------------------------------------------------ */

/* Replaces KmerImage(dir) + KmerGuts(dir, image) (kmer_image.cc:41-108, kguts.cc:34-58, 659-679):
 * maps <kmer_dir>/kmer.table.mem_map, applies the reference's three validations, uploads the table to
 * HBM, loads <kmer_dir>/function.index and otu.index. */
int ckm_open(const char *kmer_dir, int device, ckm_ctx **out);

/* Same from an image already in memory in the file format (what KmerImage::image() points at).
 * function/otu names may be NULL (count 0) when only indices are needed. */
int ckm_open_image(const void *image, size_t image_bytes, int device, const char *const *function_names,
                   int32_t n_functions, const char *const *otu_names, int32_t n_otus, ckm_ctx **out);

void ckm_close(ckm_ctx *ctx);
/* A second context over the SAME tables: the reference builds one KmerGuts per worker thread over one shared KmerImage
 * (threadpool.cc:33, kguts.h:312).  The clone has its own stream, parameters, postings and result buffers and reads the
 * parent's signature table and family tables (as loaded at the time of the call) in place; nothing is copied.  Calls on
 * different contexts may run concurrently from different threads.  Close clones before the parent.  A clone must not load
 * family tables itself. */
int ckm_clone(ckm_ctx *parent, ckm_ctx **out);

/* thread-local text of the last error (also valid when ctx creation failed) */
const char *ckm_last_error(void);

/* KmerGuts::function_at_index, kguts.h:361-366: "INVALID_OFFSET" when out of range */
const char *ckm_function_at_index(const ckm_ctx *ctx, int32_t i);
const char *ckm_otu_at_index(const ckm_ctx *ctx, int32_t i);
int32_t ckm_function_count(const ckm_ctx *ctx);
int32_t ckm_otu_count(const ckm_ctx *ctx);
uint64_t ckm_num_sigs(const ckm_ctx *ctx);
/* 16 = packed sector-friendly slots, 24 = verbatim slots (chosen at load, results identical) */
/* Buckets of the table as it sits in HBM: the image's count, or -- the default for images whose fields fit the 16-byte slots -- the
 * power of two at or above it of the library's own open-addressed table, filled from the image at load with a cheaper hash
 * (CKM_REFERENCE_HASH=1 keeps the image's order and key % num_sigs).  ckm_num_sigs reports the image's count. */
uint64_t ckm_table_buckets(const ckm_ctx *ctx);
int ckm_table_slot_bytes(const ckm_ctx *ctx);
/* cudaLimitMaxL2FetchGranularity in effect on the ctx's device (the loader asks for 32-byte sectors) */
int ckm_l2_fetch_granularity(const ckm_ctx *ctx);
/* 1 when the L2-resident slot-occupancy bitmap is in use (tables larger than L2; CKM_OCCUPANCY_BITMAP=0/1 overrides) */
int ckm_has_occupancy_bitmap(const ckm_ctx *ctx);
/* Neighbour-ordered copy of the table (built at load for tables larger than L2; CKM_CHAIN=0/1 overrides; results are
 * identical with and without it).  info[0] = entries (0 = not built), info[1] = chains, info[2] = build time in
 * microseconds, info[3] = hits of the last batch that were answered from the copy instead of a hash probe. */
int ckm_chain_info(ckm_ctx *ctx, uint64_t info[4]);
/* A/B switches of K1 (results unaffected): CKM_TUNE_PLAIN_PROBE = plain hash probing (probe_kernel) although the neighbour
 * copy exists; CKM_TUNE_UNFUSED = K1 writes hit records and scan_kernel runs even when the request could be served by the
 * fused K1 (ckm_warp_scan.cuh); CKM_TUNE_NO_FALLBACK = never suspend the neighbour copy automatically.  A library built with
 * -DCKM_EXPERIMENTS understands further bits (cache policies, block shapes, the walking probe_chain_kernel): see
 * csrc/ckm_api.cu. */
#define CKM_TUNE_PLAIN_PROBE 32u
#define CKM_TUNE_UNFUSED 0x100000u
#define CKM_TUNE_NO_FALLBACK 0x200000u
void ckm_set_tuning(ckm_ctx *ctx, uint32_t bits);
int ckm_experiments_enabled(void); /* 1 when built with -DCKM_EXPERIMENTS */
/* The automatic fall-back of K1: a batch that went through the neighbour copy but had fewer than 12 % of its probes answered
 * from it suspends the copy for 16 batches (doubling up to 1024 while retries keep failing), after which one batch tries it
 * again.  state[0] = suspended now, state[1] = batches until the retry, state[2] = suspensions so far.  Results are the same
 * on either path. */
int ckm_copy_state(const ckm_ctx *ctx, uint32_t state[3]);
/* 1 when the last batch of this ctx was served by probe_pc_kernel (scoring scan inside K1, no hit records in HBM) */
int ckm_last_batch_was_fused(const ckm_ctx *ctx);

/* ---- parameters (KmerGuts::set_default_parameters / set_parameters, kguts.cc:236-268) -------------- */
void ckm_set_default_params(ckm_ctx *ctx); /* order_constraint 0, min_hits 5, min_weighted_hits 0, max_gap 200 */
int ckm_set_params(ckm_ctx *ctx, int order_constraint, int min_hits, int min_weighted_hits, int max_gap);
void ckm_get_params(const ckm_ctx *ctx, int *order_constraint, int *min_hits, int *min_weighted_hits, int *max_gap);

/* ---- static helpers (KmerGuts::encoded_aa_kmer / decoded_kmer, kguts.cc:457-483) ------------------- */
uint64_t ckm_encoded_aa_kmer(const char *p);              /* MAX_ENCODED+1 if any invalid char */
void ckm_decoded_kmer(uint64_t encoded, char decoded[9]); /* NUL-terminated */

/* ---- the hot path: process_aa_seq[_hits] (+ find_best_call) over a batch --------------------------- */

/* residues: the n amino-acid strings concatenated (no separators); offsets[i]..offsets[i+1] is
 * sequence i.  Host pointers; the call does H2D, the kernels and D2H of what `flags` asks for. */
int ckm_call_batch(ckm_ctx *ctx, const char *residues, const uint64_t *offsets, uint32_t n, uint32_t flags,
                   ckm_batch_out_t *out);

/* The same call for packed residues (csrc/ckm_packed.cuh): sequence i is the 32-bit words [word_offsets[i], word_offsets[i+1])
 * of `packed`, SEVEN residues to a word as the digits of a base-22 number -- residue 7 w + k = (word_w / 22^k) % 22: 0..19 =
 * ACDEFGHIKLMNPQRSTVWY, 20 = any other character, 21 = end of the sequence (what an embedded NUL is to the reference,
 * kguts.cc:791; the unused digits of a sequence's last word hold it too).  Results are those of ckm_call_batch on the unpacked
 * strings; hit offsets count residues.  Moves 0.58 of the bytes over PCIe (4.57 bits per residue).  A parser can emit this form
 * in the pass it makes over every byte anyway; ckm_pack_residues converts an ASCII batch (`packed` must hold the sum of
 * ckm_packed_words(len_i) words). */
int ckm_call_batch_packed(ckm_ctx *ctx, const uint32_t *packed, const uint64_t *word_offsets, uint32_t n, uint32_t flags,
                          ckm_batch_out_t *out);
uint64_t ckm_packed_words(uint64_t n_residues); /* ceil(n / 7) */
int ckm_pack_residues(const char *residues, const uint64_t *offsets, uint32_t n, uint32_t *packed, uint64_t packed_capacity_words,
                      uint64_t *word_offsets /* n + 1 */);

/* Device-resident form for pipelines that already hold the batch in HBM: d_residues must have 32
 * readable bytes after offsets[n] (the library's own uploads zero them) and offsets[0] must be 0; max_len is the
 * longest sequence (0 = unknown, which selects the general scan kernel).  An understated max_len is detected on
 * the device and reported by the next ckm_read_totals (CKM_EINVAL).  Results stay on the device
 * (ckm_device_results); the call only enqueues work on ckm_stream(ctx). */
int ckm_call_batch_device(ckm_ctx *ctx, const void *d_residues, const uint64_t *d_offsets, uint32_t n,
                          uint64_t total_residues, uint32_t max_len, uint32_t flags);

typedef struct {
    uint32_t n;
    const uint32_t *d_n_hits;      /* n   : hits per sequence */
    const uint32_t *d_n_calls;     /* n   : calls per sequence */
    const ckm_call_t *d_calls;     /* sequence i's calls start at d_calls[offsets[i] / max(1,min_hits) + i] */
    int32_t min_hits_for_call_base;
    const ckm_best_t *d_best;      /* n */
    const uint64_t *d_totals;      /* [0] probes, [1] hits, [2] calls */
} ckm_device_out_t;
int ckm_device_results(ckm_ctx *ctx, ckm_device_out_t *out);
/* synchronises and copies {probes, hits, calls} of the last batch to the host */
int ckm_read_totals(ckm_ctx *ctx, uint64_t totals[3]);

/* ---- family voting: FamilyMapper::find_best_family_match (family_mapper.cc:46-205, 287-330) -------------- */

/* The read-only family side tables that NRLoader / load_families leave in KmerPegMapping (kmer.h:118-127):
 * kmer_to_family_id_ as a CSR (kmers[k] -> fam_ids[fam_offsets[k] .. fam_offsets[k+1]), lists deduped as
 * kmer.cc:216-230 leaves them) and family_data_ (pgf, plf, function per encoded family id 0..n_families-1).
 * Uploads device images of both; may be called again to replace them. */
int ckm_family_load(ckm_ctx *ctx, uint64_t n_kmers, const uint64_t *kmers, const uint64_t *fam_offsets,
                    const uint32_t *fam_ids, uint32_t n_families, const char *const *pgf, const char *const *plf,
                    const char *const *function);

/* The same tables built on the GPU from the proteins of families.nr, replacing NRLoader::thread_load (nr_loader.cc:131-202),
 * KmerInserter (kmer_inserter.cc:36-58) and add_fam_mapping / fam_map_insert (kmer.cc:216-268): every signature k-mer hit
 * of a protein is mapped to the protein's family, each family once per k-mer.  fam_ids[i] = encoded family id of sequence
 * i (< 2^29), or 0xFFFFFFFF for a protein without a family.  One add() call is one chunk (seq_list_t) of the loader:
 * like thread_load, whose "NO FAM FOR id" branch returns (nr_loader.cc:154-160), the first protein without a family ends
 * the chunk and the sequences after it in that call are not loaded.  finish() installs the result exactly as
 * ckm_family_load would, and reports the table's size. */
int ckm_family_nr_begin(ckm_ctx *ctx);
int ckm_family_nr_add(ckm_ctx *ctx, const uint32_t *fam_ids, const char *residues, const uint64_t *offsets, uint32_t n);
int ckm_family_nr_finish(ckm_ctx *ctx, uint32_t n_families, const char *const *pgf, const char *const *plf,
                         const char *const *function, uint64_t *n_kmers, uint64_t *n_entries);
/* the installed k-mer -> family lists as CSR on the host (lists unordered): kmers[n_kmers], offsets[n_kmers+1], ids[n_entries] */
int ckm_family_export(ckm_ctx *ctx, uint64_t n_kmers, uint64_t n_entries, uint64_t *kmers, uint64_t *offsets, uint32_t *ids);

/* best_match_t (family_mapper.h:20-28) with names left as ids */
typedef struct {
    int32_t gfam;           /* id of the best PGF (ckm_family_pgf_name), -1 if none */
    int32_t lfam;           /* encoded family id whose PLF is the best (ckm_family_plf_name), -1 if none */
    float gfam_score;       /* rolled-up weighted_total of the best PGF */
    float lfam_score;       /* weighted_total of the best family */
    float score;            /* best_call_score of find_best_call */
    int32_t function_index; /* confident best call, else -1 ("hypothetical protein", family_mapper.cc:103-123) */
} ckm_family_match_t;

/* find_best_family_match for every sequence of the batch: hits -> per-family (hit_total, weighted_total +=
 * 1/|families of the k-mer|, in hit order) -> find_best_call -> families with hit_total >= 3 whose function
 * equals the called function -> best PLF, best rolled-up PGF.  Exactly tied maxima resolve to the smallest
 * id (the reference resolves them by std::unordered_map iteration order); PGF roll-ups are f32 sums in
 * ascending hash-slot order, within 1e-6 relative of the reference's unordered sums. */
int ckm_family_batch(ckm_ctx *ctx, const char *residues, const uint64_t *offsets, uint32_t n,
                     const ckm_family_match_t **matches);

/* LookupRequest's seq_score_ in family mode (lookup_request.h:26-45, lookup_request.cc:441-464): for every sequence, every
 * family its hits touch with hit_count (== hit_total for the vector flavour of family_counts_t) and weighted_total (f32 sum
 * of 1/|families of the k-mer| in hit order), ascending family id; plus find_best_call of the sequence and the
 * ckm_family_batch match.  Arrays are owned by ctx and valid until its next call. */
typedef struct {
    uint32_t id;
    uint32_t hit_count;
    float weighted_total;
} ckm_score_t;
typedef struct {
    uint32_t n;
    const ckm_score_t *scores;
    const uint64_t *score_offsets; /* n + 1 */
    const ckm_best_t *best;
    const ckm_family_match_t *matches;
} ckm_family_scores_t;
int ckm_family_scores(ckm_ctx *ctx, const char *residues, const uint64_t *offsets, uint32_t n, ckm_family_scores_t *out);

const char *ckm_family_pgf_name(const ckm_ctx *ctx, int32_t gfam);   /* "" if out of range */
const char *ckm_family_plf_name(const ckm_ctx *ctx, int32_t lfam);
/* best_match_t::function for a match: function.index name or "hypothetical protein" */
const char *ckm_family_function_name(const ckm_ctx *ctx, const ckm_family_match_t *m);

/* ---- fastq reads: 6-frame translation + calling + family voting + best frame ---------------------------
 * FqProcessRequest::on_parsed_seq (fq_process_request.cc:298-365) over DNASequence::get_possible_proteins
 * (dna_seq.cc:9-47, dna_seq.h:28-111) and TranslationTable code 11 (trans_table.cc:8-84). */
typedef struct {
    uint32_t length;        /* prot.length() of the fragment (> 10) */
    ckm_family_match_t m;   /* find_best_family_match of the fragment */
} ckm_fq_match_t;           /* 28 bytes */

typedef struct {
    uint32_t n;                    /* reads */
    const int32_t *best_frame;     /* n: 1,2,3,-1,-2,-3, or 0 when best_score == 0 (no output line) */
    const double *best_score;      /* n: sum of the winning frame's fragment scores up to the winning fragment */
    const uint64_t *match_offsets; /* n+1: CSR into matches */
    const ckm_fq_match_t *matches; /* the winning frame's processed fragments, in order, up to the winning one */
    uint64_t n_fragments;          /* fragments longer than 10 aa over all reads and frames */
    uint64_t n_probes;
} ckm_fq_out_t;

/* bases: the n reads concatenated; offsets[i]..offsets[i+1] is read i (IUPAC letters, either case, U = T).
 * Needs ckm_family_load.  Every fragment is called with the ctx's current parameters (the reference's fq
 * path never calls set_parameters: fq_process_request.cc:241). */
int ckm_fq_batch(ckm_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n, ckm_fq_out_t *out);

/* D1/D2 alone: the protein fragments (> min_len aa) of all six frames, as a protein batch.  frag_frame_offsets
 * has 6n+1 entries: fragments of (read r, frame slot s) are [ffo[6r+s], ffo[6r+s+1]) with slots ordered
 * {1,2,3,-1,-2,-3}.  Host pointers into ctx-owned pinned memory, valid until the next call. */
typedef struct {
    uint32_t n_reads;
    uint64_t n_fragments;
    const uint64_t *frag_frame_offsets; /* 6n+1 */
    const uint64_t *frag_offsets;       /* n_fragments+1, into residues */
    const char *residues;
} ckm_fq_fragments_t;
int ckm_fq_translate(ckm_ctx *ctx, const char *bases, const uint64_t *offsets, uint32_t n, uint32_t min_len,
                     ckm_fq_fragments_t *out);

/* ---- /add postings and the /matrix pairwise shared-k-mer counts ------------------------------------------
 * AddRequest (add_request.cc:133, 164-170) pushes one kmer_to_id_ entry per hit occurrence (kmer.cc:174-214);
 * MatrixRequest (matrix_request.cc:78-95, 130-161) counts, for the proteins of one request in order,
 * distance[(eid_i, e)]++ for every hit of protein i and every posting e of the hit's k-mer with e != eid_i
 * and e already seen in this request (matrix_proteins_ is set before protein i is processed). */

/* append the postings of a batch: eids[i] is KmerPegMapping::encode_id(id_i), assigned by the caller */
int ckm_postings_add(ckm_ctx *ctx, const uint32_t *eids, const char *residues, const uint64_t *offsets, uint32_t n);
/* same, reusing the hits of the ckm_call_batch that just ran on these n sequences with CKM_WANT_HITS set
 * (AddRequest computes calls, OTU stats and hits in one process_aa_seq_hits call: add_request.cc:133) */
int ckm_postings_append_last(ckm_ctx *ctx, const uint32_t *eids, uint32_t n);
void ckm_postings_clear(ckm_ctx *ctx);
uint64_t ckm_postings_count(const ckm_ctx *ctx); /* (k-mer, peg) entries held */
/* one set of postings per KmerPegMapping (the server's "/mapping/<key>" routes, krequest2.cc:440-456); key 0 is selected
 * initially.  The other ckm_postings_* calls and ckm_matrix_rows act on the selected set. */
int ckm_postings_select(ckm_ctx *ctx, uint32_t key);

typedef struct {
    uint32_t eid_i; /* the protein being processed */
    uint32_t eid_j; /* a previously seen protein of the same request sharing signature k-mers */
    uint64_t count; /* sum over hits of i of the multiplicity of eid_j in the k-mer's postings */
} ckm_pair_t;       /* 16 bytes */

/* Rows [row_begin, row_end) of the request's strictly-lower-triangular count matrix, as unordered COO
 * entries (the handler layer orders them by (eid_i, eid_j) like the reference's std::map and merges rows of
 * a repeated id).  The whole request (all n proteins) must be passed so that membership is known; only the
 * selected rows are probed -- this is the unit of row-block sharding across GPUs. */
int ckm_matrix_rows(ckm_ctx *ctx, const uint32_t *eids, const char *residues, const uint64_t *offsets, uint32_t n,
                    uint32_t row_begin, uint32_t row_end, const ckm_pair_t **pairs, uint64_t *n_pairs);

/* Multi-GPU /matrix (row blocks over the ranks; the caller's collective moves device memory, e.g. ncclAllGather): a rank
 * extracts the hits of ITS protein block with ckm_postings_add, publishes them with ckm_postings_device, every rank installs the
 * gathered postings of all ranks with ckm_postings_import_device and computes its row block with ckm_matrix_rows_device, whose
 * tile stays in HBM with every row's entries ordered by partner id (*d_pairs is valid until the next call on the ctx;
 * *postings_walked = posting-list entries behind the rows' hits).  close_kmers_b200/parallel.py drives this over torch.distributed. */
int ckm_postings_device(ckm_ctx *ctx, const uint64_t **d_keys, const uint32_t **d_eids, uint64_t *n);
int ckm_postings_import_device(ckm_ctx *ctx, const uint64_t *d_keys, const uint32_t *d_eids, uint64_t n);
int ckm_matrix_rows_device(ckm_ctx *ctx, const uint32_t *eids, const char *residues, const uint64_t *offsets, uint32_t n,
                           uint32_t row_begin, uint32_t row_end, const ckm_pair_t **d_pairs, uint64_t *n_pairs,
                           uint64_t *postings_walked);

/* LookupRequest::on_hit without families (lookup_request.cc:466-478): for every sequence of the batch, the pegs of the
 * selected postings that share a hit k-mer with it and how many postings did (seq_score_[eid].hit_count).  pairs[k] =
 * {eid_i = sequence index in the batch, eid_j = peg id, count}; pair_offsets has n + 1 entries; pegs ascending per sequence. */
int ckm_postings_scores(ckm_ctx *ctx, const char *residues, const uint64_t *offsets, uint32_t n, const ckm_pair_t **pairs,
                        const uint64_t **pair_offsets);

/* ---- image builder: KmerGuts(dir, nbuckets) + insert_kmer + save_kmer_hash_table (kguts.cc:77-115, 188-234).
 * Host-side; writes the reference's file bytes (header + nbuckets slots) into image_out, which must be
 * exactly 24 + 24*nbuckets bytes.  keys > MAX_ENCODED are skipped like kguts.cc:206-210. */
int ckm_image_build(uint64_t nbuckets, uint64_t n, const uint64_t *keys, const int32_t *function_index,
                    const int32_t *otu_index, const uint16_t *avg_from_end, const float *function_wt, void *image_out,
                    size_t image_bytes);

/* The same image built on the GPU ("next" row N4; write_hashtable, build_signature_kmers.cc:860-898).  Byte-identical to
 * ckm_image_build -- and to the reference's sequential insert_kmer loop -- by priority linear probing with priority =
 * insertion index (see csrc/ckm_build.cuh).  n < 2^32 - 1. */
int ckm_image_build_device(int device, uint64_t nbuckets, uint64_t n, const uint64_t *keys, const int32_t *function_index,
                           const int32_t *otu_index, const uint16_t *avg_from_end, const float *function_wt, void *image_out,
                           size_t image_bytes);
/* ... and opened directly, without the file or its host copy: the context ckm_open_image would give for that image. */
int ckm_open_built(int device, uint64_t nbuckets, uint64_t n, const uint64_t *keys, const int32_t *function_index,
                   const int32_t *otu_index, const uint16_t *avg_from_end, const float *function_wt,
                   const char *const *function_names, int32_t n_functions, const char *const *otu_names, int32_t n_otus,
                   ckm_ctx **out);
/* compute_weight_of_signature (build_signature_kmers.cc:841-853): NSF = sequences with a signature, KS = distinct
 * signatures, NSi = sequences containing this k-mer, NFj = sequences with its function, NSiFj = sequences with both. */
float ckm_signature_weight(float NSF, float KS, float NSi, float NFj, float NSiFj);

/* Calibration for roofline reports: independent random `bytes`-sized (16 or 32) reads over the resident
 * table, `unroll` (1, 4 or 8) in flight per thread, `rounds` rounds, blocks_per_sm x SM-count blocks of 256
 * threads.  Returns accesses per second -- the gather ceiling this GPU sustains at this table size. */
int ckm_calibrate_gather(ckm_ctx *ctx, int bytes, int unroll, uint32_t rounds, int blocks_per_sm, double *accesses_per_s,
                         double *ms);

/* page-locked host memory for batch assembly (fast H2D) */
int ckm_host_alloc(void **p, size_t bytes);
void ckm_host_free(void *p);

/* the CUDA stream (cudaStream_t) all work of this ctx is launched on, for event timing by the caller */
void *ckm_stream(ckm_ctx *ctx);
/* number of kernel launches issued by this ctx so far */
uint64_t ckm_launch_count(const ckm_ctx *ctx);
int ckm_synchronize(ckm_ctx *ctx);

/* Per-kernel device timing for roofline reports: when enabled, every batch records CUDA events around
 * the probe kernel (K1) and the scan kernel (K2) on ckm_stream(ctx); ckm_profile_read synchronises,
 * returns the summed durations (ms) and the number of batches since the last read, and resets. */
void ckm_profile_enable(ckm_ctx *ctx, int on);
int ckm_profile_read(ckm_ctx *ctx, double *probe_ms, double *scan_ms, uint64_t *batches);

#ifdef __cplusplus
}
#endif
#endif /* CKM_H */

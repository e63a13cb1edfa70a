/*
 * ckm_handlers.h -- request-handler surface above the compute ABI (ckm.h).
 *
 * These entry points reproduce, for ONE body chunk of already-parsed sequences, the response text the
 * reference's handlers write for it: the worker-lambda bodies of QueryRequest (query_request.cc:79-152),
 * AddRequest (add_request.cc:102-170), MatrixRequest (matrix_request.cc:78-95, 163-189) and
 * FqProcessRequest (fq_process_request.cc:298-365).  Socket / HTTP plumbing (Boost.Asio) is out of scope;
 * a front end hands over (ids, sequences) and writes the returned text to its client.
 *
 * Text is returned as a malloc'ed NUL-terminated string to be released with ckm_free_text().
 */
#ifndef CKM_HANDLERS_H
#define CKM_HANDLERS_H
#include "ckm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* POST /query (query_request.cc:103-152).  details != 0 adds HIT lines; find_best_call != 0 switches to
 * "<id>\t<function>\t<score>\t<weighted score>" lines for sequences with a non-empty best function. */
int ckm_query_text(ckm_ctx *ctx, const char *const *ids, const char *residues, const uint64_t *offsets, uint32_t n,
                   int details, int find_best_call, char **text);

/* KmerPegMapping's peg-id table (kmer.h:109-116, kmer.cc:272-295): ids are assigned in order of first
 * encode_id.  Host-side state shared by /add and /matrix requests of one server. */
typedef struct ckm_mapping ckm_mapping;
ckm_mapping *ckm_mapping_new(void);
void ckm_mapping_free(ckm_mapping *m);
uint32_t ckm_mapping_encode_id(ckm_mapping *m, const char *peg);
/* assign_new_peg_id (kmer.h:109-116), what load_families uses: always a fresh id; the name then maps to the newest one */
uint32_t ckm_mapping_assign_new_id(ckm_mapping *m, const char *peg);
const char *ckm_mapping_decode_id(const ckm_mapping *m, uint32_t id); /* "" when unknown */

/* POST /add (add_request.cc:102-170): unless `silent`, "PROTEIN-ID", "CALL" lines, "OTU-COUNTS" and a
 * "BEST-CALL\t<id>\t<function>\t<score>\t<weighted>\t<offset>" line per sequence ("hypothetical protein" for an
 * empty or ambiguous call; <offset> is printed as 0 where the reference prints an uninitialised float, i.e.
 * when there are no calls); then every hit is appended to the ctx's k-mer -> peg postings. */
int ckm_add_text(ckm_ctx *ctx, ckm_mapping *m, const char *const *ids, const char *residues, const uint64_t *offsets,
                 uint32_t n, int silent, char **text);

/* POST /matrix (matrix_request.cc:78-95, 163-189) for a request that fits one chunk: rows
 * "<peg1>\t<peg2>\t<count>\t<count/(len1+len2)>\n" ordered by (encoded id 1, encoded id 2).  The HTTP status
 * lines of process_results are the front end's business and are not included. */
int ckm_matrix_text(ckm_ctx *ctx, ckm_mapping *m, const char *const *ids, const char *residues, const uint64_t *offsets,
                    uint32_t n, char **text);

/* the ordering / merging step of ckm_matrix_text on its own, for callers that computed row blocks on several
 * GPUs: sorts by (eid_i, eid_j) and sums entries with equal keys, in place; returns the new count */
uint64_t ckm_matrix_merge_pairs(ckm_pair_t *pairs, uint64_t n_pairs);

/* POST /fq_lookup (fq_process_request.cc:298-365): one line per read with a positive best score,
 * "<id>\t<frame>\t<best score>\t<len>\t<gfam>\t<gscore>\t<lfam>\t<lscore>\t<function>\t<score>[\t<len>...]*\n";
 * reads with an empty id are skipped.  Needs ckm_family_load. */
int ckm_fq_text(ckm_ctx *ctx, const char *const *ids, const char *bases, const uint64_t *offsets, uint32_t n, char **text);

/* POST /lookup (lookup_request.cc:132-400).  Options are the LookupRequest members its constructor reads from the request
 * parameters ("kmer_hit_threhsold" [sic], find_best_match, find_reps, allow_ambiguous_functions, target_genus via
 * genus_map_).  fams[f] is family_data_[f] (kmer.h:58-68).
 *   family_mode && find_best_match: one line per sequence "<id>\t<best PGF>\t<score>\t<best PLF of the target genus>\t<score>\t
 *     <function>\t<best call score>\t<weighted>\n" (201-327);
 *   family_mode: "<id>\n", then per family, best weighted_total first and stopping at the first one under the hit threshold,
 *     "<hits>\t<hits>\t<weighted>\t<pgf>\t<plf>\t<total_size>\t<count>\t<hits/total_size>\t<function>\n" (+ "///\n" with find_reps),
 *     then "//\n" (329-377);
 *   otherwise (peg mode): "<id>\n", "<peg>\t<hit count>\n" per peg of the mapping's postings -- only when the threshold is 0,
 *     because that mode never increments hit_total (466-478) -- then "//\n".
 * Where the reference leaves an order to std::unordered_map iteration or to std::sort on equal keys (tied weighted_total,
 * the order of the PGF roll-up's f32 additions), ids ascend here. */
typedef struct {
    const char *pgf, *plf, *function;
    uint64_t genus_id, total_size;
    uint16_t count;
} ckm_family_data_t;
typedef struct {
    int family_mode;
    unsigned int kmer_hit_threshold; /* default 3 */
    int find_best_match, find_reps, allow_ambiguous_functions;
    uint64_t target_genus_id;
} ckm_lookup_options_t;
int ckm_lookup_text(ckm_ctx *ctx, ckm_mapping *m, const ckm_family_data_t *fams, uint32_t n_fams, const ckm_lookup_options_t *opt,
                    const char *const *ids, const char *residues, const uint64_t *offsets, uint32_t n, char **text);

/* FamilyMapper::find_all_matches (family_mapper.cc:207-285; its only caller is the reference's test_family_mapper.cc:128) for
 * every sequence of a chunk: "<id>\n", then one line per family the sequence's hits touched -- hit count, hit total, weighted
 * total, PGF, PLF, total size, count, hit count / total size, function -- best weighted total first, up to the first family
 * with fewer than three hits (kmer_hit_threshold_, family_mapper.cc:7), then "//\n".  The same listing as POST /lookup in
 * family mode without find_best_match (lookup_request.cc:329-377); families with exactly equal weighted totals come out by
 * ascending id here (std::sort over unordered_map order in the reference). */
int ckm_family_all_matches_text(ckm_ctx *ctx, const ckm_family_data_t *fams, uint32_t n_fams, const char *const *ids,
                                const char *residues, const uint64_t *offsets, uint32_t n, char **text);

/* FamilyMapper::find_best_family_match as text, one "<gfam>\t<gscore>\t<lfam>\t<lscore>\t<function>\t<score>\n"
 * line per sequence (operator<< of best_match_t, family_mapper.h:70-75) */
int ckm_family_text(ckm_ctx *ctx, const char *residues, const uint64_t *offsets, uint32_t n, char **text);

/* KmerGuts::format_call / format_hit / format_otu_stats (kguts.cc:939-973) for callers that assemble their own
 * responses.  `otus` are the ascending-otu_index pairs of one sequence; the function applies
 * KmerOtuStats::finalize's count-descending std::sort (kguts.h:214-218) and prints the top five. */
char *ckm_format_call(const ckm_ctx *ctx, const ckm_call_t *call);
char *ckm_format_hit(const ckm_ctx *ctx, const ckm_hit_t *hit);
char *ckm_format_otu_stats(const char *id, uint64_t seq_len, const ckm_otu_t *otus, uint64_t n_otus);

/* the `function` string find_best_call returns for a ckm_best_t (kguts.cc:1160, 1176-1196) */
char *ckm_best_function(const ckm_ctx *ctx, const ckm_best_t *best);

void ckm_free_text(char *text);

#ifdef __cplusplus
}
#endif
#endif

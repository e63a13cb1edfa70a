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

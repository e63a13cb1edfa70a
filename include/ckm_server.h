/*
 * ckm_server.h -- the data formats and the request front end either side of the hot path.
 *
 *   - streaming FASTA / FASTQ parsers with the reference's state machines (fasta_parser.h:40-140,
 *     fastq_parser.h:41-147), producing the flat (ids, residues, offsets) batches the compute ABI takes;
 *   - the HTTP-ish request head parsing and routing of KmerRequest2 (krequest2.cc:26-33, 87-159, 160-245, 247-486);
 *   - ckm_kser_main: a dependency-free replacement for the kser binary (kser.cc:37-341, kserver.cc:13-216) that serves
 *     the same routes with the same bytes on the wire, every body chunk going through the GPU handlers of ckm_handlers.h.
 *
 * Everything here is host code; it needs no GPU until ckm_kser_main opens the engine.
 */
#ifndef CKM_SERVER_H
#define CKM_SERVER_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- sequence parsers -------------------------------------------------------------------------------------------- */
#define CKM_FORMAT_FASTA 0 /* FastaParser: '>' id [blank def] \n data lines; '\r' ignored; data = letters and '*' */
#define CKM_FORMAT_FASTQ 1 /* FastqParser: '@' id [blank def] \n one data line \n '+' line \n quality line \n */

typedef struct ckm_seq_parser ckm_seq_parser;

typedef struct {
    uint32_t n;             /* sequences completed since the previous take */
    const char *const *ids; /* n NUL-terminated ids */
    const char *residues;   /* concatenated sequences */
    const uint64_t *offsets; /* n + 1 */
    uint64_t n_errors;      /* "Error found: ..." events so far (the reference prints them and carries on) */
} ckm_seq_batch_t;

ckm_seq_parser *ckm_seq_parser_new(int format);
void ckm_seq_parser_free(ckm_seq_parser *p);
/* parse_char over data[0..n): parser state carries over between calls, like the handlers' parser_ member */
void ckm_seq_parser_feed(ckm_seq_parser *p, const char *data, size_t n);
/* parse_complete (fasta_parser.cc:30-36, fastq_parser.cc:30-36): emits the sequence in progress -- also when nothing
 * was parsed, which is how an empty body yields one ("", "") work item in the reference */
void ckm_seq_parser_complete(ckm_seq_parser *p);
/* residue bytes of completed, not yet taken sequences */
uint64_t ckm_seq_parser_pending(const ckm_seq_parser *p);
/* hand over the completed sequences; pointers stay valid until the next feed / complete / take / free on p */
void ckm_seq_parser_take(ckm_seq_parser *p, ckm_seq_batch_t *out);
/* text of the most recent error event ("Missing >", "Bad data character 'x'", ...) with line and id, "" if none */
const char *ckm_seq_parser_last_error(const ckm_seq_parser *p);

/* ---- request head ------------------------------------------------------------------------------------------------ */
/* Parses a request head (request line + header lines, up to and including the empty line) the way
 * KmerRequest2::read_initial_line / read_headers / process_request do and returns a malloc'ed description (release with
 * ckm_free_text): one "key=value" line each for type, path, raw parameters, fragment, version, every parameter
 * ("param.<k>") and header ("header.<k>", key lower-cased) in map order, and "decision=<...>" -- what process_request
 * does next: "invalid" (request line does not match request_regex), "501 chunked", "100-continue" is listed as
 * "continue=1", "respond <code> <status>", "get <route>", or "post <action> key=<mapping key> length=<n>". */
char *ckm_http_describe(const char *head, size_t n);

/* ---- the server -------------------------------------------------------------------------------------------------- */
/* kser [options] listen-port kmer-data-dir   (options: see INTEGRATION.md / --help).  Returns the process exit code. */
int ckm_kser_main(int argc, char **argv);

#ifdef __cplusplus
}
#endif
#endif

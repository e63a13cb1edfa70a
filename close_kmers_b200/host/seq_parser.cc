// Streaming FASTA / FASTQ parsers with the reference's state machines, emitting flat batches.
//
// The reference feeds its parsers one character at a time (query_request.cc:52-56) and collects std::string pairs; here a
// block is scanned line-wise where the state allows it (identifier and data lines are the bulk of a body), and completed
// sequences land directly in the (ids, residues, offsets) layout ckm_call_batch takes.  What is accepted, skipped or
// reported is the same, state by state: fasta_parser.h:40-140, fastq_parser.h:41-147.
#include "seq_parser.h"

using ckm_parse::is_alpha;
using ckm_parse::is_blank;

void ckm_seq_parser::feed_fasta(const char *d, size_t n) {
    size_t i = 0;
    while (i < n) {
        if (state == s_data) {
            // bulk of the body: letters and '*' up to the end of the line
            const size_t j = i + take_run<true>(d + i, n - i);
            if (j > i) {
                i = j;
                continue;
            }
        } else if (state == s_defline) {
            const char *nl = (const char *)memchr(d + i, '\n', n - i);  // the definition is not kept
            if (!nl) return;
            i = (size_t)(nl - d);
        }
        const unsigned char ch = (unsigned char)d[i];
        i++;
        if (ch == '\n') line_number++;
        if (ch == '\r') continue;
        switch (state) {
        case s_start:
            if (ch != '>') error("Missing >");
            else state = s_id;
            break;
        case s_id:
            if (is_blank(ch)) state = s_defline;
            else if (ch == '\n') state = s_data;
            else cur_id.push_back((char)ch);
            break;
        case s_defline:
            if (ch == '\n') state = s_data;
            break;
        case s_data:
            if (ch == '\n') state = s_id_or_data;
            else error(std::string("Bad data character '") + (char)ch + "'");  // letters and '*' were taken above
            break;
        case s_id_or_data:
            if (ch == '>') {
                emit();
                state = s_id;
            } else if (ch == '\n') {
            } else if (is_alpha(ch)) {
                residues.push_back((char)ch);
                state = s_data;
            } else {
                error(std::string("Bad id or data character '") + (char)ch + "'");
            }
            break;
        default:
            break;
        }
    }
}

void ckm_seq_parser::feed_fastq(const char *d, size_t n) {
    size_t i = 0;
    while (i < n) {
        if (state == s_data) {
            const size_t j = i + take_run<false>(d + i, n - i);
            if (j > i) {
                i = j;
                continue;
            }
        } else if (state == s_defline || state == s_plus_line || state == s_qual) {
            const char *nl = (const char *)memchr(d + i, '\n', n - i);  // nothing on these lines is kept
            if (!nl) return;
            i = (size_t)(nl - d);
        }
        const unsigned char ch = (unsigned char)d[i];
        i++;
        if (ch == '\n') line_number++;
        switch (state) {
        case s_start:
            if (ch == '>') error("Starts with >. Is this a fasta file not a fastq file?");
            else if (ch != '@') error("Missing @");
            else state = s_id;
            break;
        case s_id:
            if (is_blank(ch)) state = s_defline;
            else if (ch == '\n') state = s_data;
            else cur_id.push_back((char)ch);
            break;
        case s_defline:
            if (ch == '\n') state = s_data;
            break;
        case s_data:
            if (ch == '\n') state = s_plus_start;
            else error(std::string("Bad data character '") + (char)ch + "'");
            break;
        case s_plus_start:
            if (ch != '+') error("Missing +");
            else state = s_plus_line;
            break;
        case s_plus_line:
            if (ch == '\n') state = s_qual;
            break;
        case s_qual:
            if (ch == '\n') {
                emit();
                state = s_start;
            }
            break;
        default:
            break;
        }
    }
}

extern "C" ckm_seq_parser *ckm_seq_parser_new(int format) {
    if (format != CKM_FORMAT_FASTA && format != CKM_FORMAT_FASTQ) return nullptr;
    ckm_seq_parser *p = new ckm_seq_parser();
    p->format = format;
    return p;
}
extern "C" void ckm_seq_parser_free(ckm_seq_parser *p) { delete p; }
extern "C" void ckm_seq_parser_feed(ckm_seq_parser *p, const char *data, size_t n) {
    if (!p || !n) return;
    if (p->format == CKM_FORMAT_FASTA) p->feed_fasta(data, n);
    else p->feed_fastq(data, n);
}
extern "C" void ckm_seq_parser_complete(ckm_seq_parser *p) {
    if (p) p->emit();
}
extern "C" uint64_t ckm_seq_parser_pending(const ckm_seq_parser *p) { return p ? p->offsets.back() : 0; }
extern "C" void ckm_seq_parser_take(ckm_seq_parser *p, ckm_seq_batch_t *out) {
    p->out_ids.swap(p->ids);
    p->out_residues.swap(p->residues);
    p->out_offsets.swap(p->offsets);
    p->ids.clear();
    const size_t completed = (size_t)p->out_offsets.back();
    p->residues.assign(p->out_residues, completed, std::string::npos);  // the sequence in progress stays with the parser
    p->out_residues.resize(completed);
    p->offsets.assign(1, 0);
    p->out_ptrs.clear();
    for (const auto &s : p->out_ids) p->out_ptrs.push_back(s.c_str());
    out->n = (uint32_t)p->out_ids.size();
    out->ids = p->out_ptrs.data();
    out->residues = p->out_residues.data();
    out->offsets = p->out_offsets.data();
    out->n_errors = p->n_errors;
}
extern "C" const char *ckm_seq_parser_last_error(const ckm_seq_parser *p) { return p ? p->last_error.c_str() : ""; }

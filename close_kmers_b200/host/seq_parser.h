// The parser object behind ckm_seq_parser_* (include/ckm_server.h), for host code that wants to take ownership of the
// parsed batches (the server hands them to a worker thread) instead of borrowing pointers.
#ifndef CKM_HOST_SEQ_PARSER_H
#define CKM_HOST_SEQ_PARSER_H
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ckm_server.h"

namespace ckm_parse {
inline bool is_alpha(unsigned char c) { return (unsigned)((c | 0x20u) - 'a') < 26u; }  // isalpha in the "C" locale
inline bool is_blank(unsigned char c) { return c == ' ' || c == '\t'; }
}  // namespace ckm_parse

struct ckm_seq_parser {
    enum State { s_start, s_id, s_defline, s_data, s_id_or_data, s_plus_start, s_plus_line, s_qual };
    int format;
    State state = s_start;
    int line_number = 1;
    std::string cur_id;  // the sequence in progress is the tail of `residues`, past offsets.back()
    uint64_t n_errors = 0;
    std::string last_error;

    // completed sequences not yet taken, and the batch most recently handed out
    std::vector<std::string> ids, out_ids;
    std::string residues, out_residues;
    std::vector<uint64_t> offsets{0}, out_offsets;
    std::vector<const char *> out_ptrs;

    void emit() {  // call_callback + reset (fasta_parser.h:112-116, 158-164)
        ids.push_back(cur_id);
        offsets.push_back(residues.size());
        cur_id.clear();
    }
    // Appends the longest prefix of d[0..n) made of sequence characters and returns its length.  Whole lines are taken
    // with one memchr + one vectorisable validity pass + one copy.
    template <bool STAR>
    size_t take_run(const char *d, size_t n) {
        const char *nl = (const char *)memchr(d, '\n', n);
        const size_t end = nl ? (size_t)(nl - d) : n;
        unsigned bad = 0;
        for (size_t k = 0; k < end; k++) {
            const unsigned char c = (unsigned char)d[k];
            bad |= !(((unsigned)((c | 0x20u) - 'a') < 26u) | (STAR & (c == '*')));
        }
        size_t take = end;  // up to the newline (or the end of the block): the state machine takes it from there
        if (bad) {          // stop at the first character that is not sequence data
            take = 0;
            while (ckm_parse::is_alpha((unsigned char)d[take]) || (STAR && d[take] == '*')) take++;
        }
        residues.append(d, take);
        return take;
    }
    void error(const std::string &what) {
        n_errors++;
        last_error = "Error found: " + what + " at line " + std::to_string(line_number) + " id='" + cur_id + "'";
    }
    void feed_fasta(const char *d, size_t n);
    void feed_fastq(const char *d, size_t n);
    // move the completed sequences out (the sequence in progress stays): the caller owns the three containers
    void take_owned(std::vector<std::string> &o_ids, std::string &o_residues, std::vector<uint64_t> &o_offsets) {
        o_ids.clear();
        o_ids.swap(ids);
        o_residues.clear();
        o_residues.swap(residues);
        o_offsets.clear();
        o_offsets.swap(offsets);
        const size_t completed = (size_t)o_offsets.back();
        residues.assign(o_residues, completed, std::string::npos);
        o_residues.resize(completed);
        offsets.assign(1, 0);
    }
};

#endif

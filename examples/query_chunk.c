/* Minimal C caller of the ABI: what QueryRequest's worker lambda does for one body chunk (query_request.cc:79-152).
 *   cc -Iinclude examples/query_chunk.c -Lclose_kmers_b200 -lckm -Wl,-rpath,$PWD/close_kmers_b200 -o query_chunk
 *   ./query_chunk <kmer_dir> */
#include <stdio.h>
#include <string.h>

#include "ckm.h"
#include "ckm_handlers.h"

int main(int argc, char **argv) {
    if (argc < 2) {
        fprintf(stderr, "usage: %s <kmer_dir>\n", argv[0]);
        return 2;
    }
    ckm_ctx *ctx = NULL;
    if (ckm_open(argv[1], 0, &ctx) != CKM_OK) {
        fprintf(stderr, "ckm_open: %s\n", ckm_last_error());
        return 1;
    }
    const char *ids[2] = {"fig|83333.1.peg.1", "fig|83333.1.peg.2"};
    const char *seqs[2] = {"MKVLAAGIVGLCAAGHRPKNAEAERLATELGLEYRHIDDYLSHRLPRNLGI", "MSTNPKPQRKTKRNTNRRPQDVKFPGG"};
    char residues[256];
    uint64_t offsets[3] = {0, 0, 0};
    for (int i = 0; i < 2; i++) {
        memcpy(residues + offsets[i], seqs[i], strlen(seqs[i]));
        offsets[i + 1] = offsets[i] + strlen(seqs[i]);
    }
    ckm_set_default_params(ctx); /* set_parameters() resets to the defaults on every request (kguts.cc:244-247) */
    char *text = NULL;
    if (ckm_query_text(ctx, ids, residues, offsets, 2, /*details=*/0, /*find_best_call=*/0, &text) != CKM_OK) {
        fprintf(stderr, "ckm_query_text: %s\n", ckm_last_error());
        ckm_close(ctx);
        return 1;
    }
    fputs(text, stdout);
    ckm_free_text(text);
    ckm_close(ctx);
    return 0;
}

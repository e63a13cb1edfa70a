/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  See ckm_oracle.h for the rules and the parity status
 * (PINNED against oracle/_ref, the reference's own object code, and tests/golden/).
 *
 * Plain-C restatement of the reference's CPU algorithm.  Every function cites the reference lines it
 * follows (paths relative to /root/reference).
 */
#define _GNU_SOURCE
#include "ckm_oracle.h"

#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

struct orc_table {
    const ckm_image_header_t *hdr;
    const ckm_sig_kmer_t *slots;
    uint64_t num_sigs;
    void *map_base; /* non-NULL when we mmap'ed it */
    size_t map_len;
};

/* ---- T2: kmer_image.cc:41-108 (the three validations at 87, 95, 101) ---- */
int orc_open_image(const void *image, size_t bytes, orc_table **out) {
    if (!image || bytes < sizeof(ckm_image_header_t)) return CKM_EFORMAT;
    const ckm_image_header_t *h = (const ckm_image_header_t *)image;
    if (bytes != sizeof(ckm_sig_kmer_t) * h->num_sigs + sizeof(ckm_image_header_t)) return CKM_EFORMAT;
    if (h->version != 1) return CKM_EFORMAT;
    if (h->entry_size != sizeof(ckm_sig_kmer_t)) return CKM_EFORMAT;
    orc_table *t = (orc_table *)calloc(1, sizeof *t);
    t->hdr = h;
    t->slots = (const ckm_sig_kmer_t *)(h + 1);
    t->num_sigs = h->num_sigs;
    *out = t;
    return 0;
}

int orc_open(const char *dir, orc_table **out) {
    char path[4096];
    snprintf(path, sizeof path, "%s/kmer.table.mem_map", dir);
    int fd = open(path, O_RDONLY);
    if (fd < 0) return CKM_EIO;
    struct stat sb;
    if (fstat(fd, &sb) < 0) { close(fd); return CKM_EIO; }
    void *p = mmap(0, (size_t)sb.st_size, PROT_READ, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return CKM_EIO;
    int rc = orc_open_image(p, (size_t)sb.st_size, out);
    if (rc) { munmap(p, (size_t)sb.st_size); return rc; }
    (*out)->map_base = p;
    (*out)->map_len = (size_t)sb.st_size;
    return 0;
}

void orc_close(orc_table *t) {
    if (!t) return;
    if (t->map_base) munmap(t->map_base, t->map_len);
    free(t);
}

uint64_t orc_num_sigs(const orc_table *t) { return t->num_sigs; }

/* ---- Q1 defaults: kguts.cc:236-242 ---- */
void orc_default_params(orc_params_t *p) {
    p->order_constraint = 0;
    p->min_hits = 5;
    p->min_weighted_hits = 0;
    p->max_gap = 200;
}

/* ---- E1: kguts.cc:273-339 (switch over the 20 uppercase letters, default 20) ---- */
static const char PROT_ALPHA[21] = "ACDEFGHIKLMNPQRSTVWY"; /* kguts.cc:30-32 */
uint8_t orc_to_amino_acid_off(char c) {
    for (int i = 0; i < 20; i++)
        if (PROT_ALPHA[i] == c) return (uint8_t)i;
    return 20;
}

/* E1, table variant: kmer_encoder.cc:7-13 fills entries 0..254 with 20 and leaves [255] untouched;
 * the restatement returns 20 there (any value >= 20 is "invalid" to every caller: kmer_encoder.h:46). */
uint8_t orc_encoder_aa_to_offset(uint8_t c) { return orc_to_amino_acid_off((char)c); }

/* ---- E2: kguts.cc:438-455 (Horner base 20, r_0 most significant), 457-471, 473-483 ---- */
static uint64_t encoded_kmer(const uint8_t *p) {
    uint64_t k = p[0];
    for (int i = 1; i <= CKM_KMER_SIZE - 1; i++) k = k * 20 + p[i];
    return k;
}
uint64_t orc_encoded_aa_kmer(const char *p) {
    uint8_t off[CKM_KMER_SIZE];
    for (int j = 0; j < CKM_KMER_SIZE; j++) {
        off[j] = orc_to_amino_acid_off(p[j]);
        if (off[j] >= 20) return CKM_MAX_ENCODED + 1;
    }
    return encoded_kmer(off);
}
void orc_decoded_kmer(uint64_t k, char out[9]) {
    out[CKM_KMER_SIZE] = 0;
    for (int i = CKM_KMER_SIZE - 1; i >= 0; i--) {
        out[i] = PROT_ALPHA[k % 20];
        k /= 20;
    }
}

/* ---- P1: kguts.cc:585-602 (key % size_hash, linear probe, stop at match or empty) ---- */
int64_t orc_lookup_hash_entry(const orc_table *t, uint64_t key) {
    uint64_t h = key % t->num_sigs;
    while (t->slots[h].which_kmer != key && t->slots[h].which_kmer <= CKM_MAX_ENCODED) h = (h + 1) % t->num_sigs;
    return t->slots[h].which_kmer > CKM_MAX_ENCODED ? -1 : (int64_t)h;
}

/* ---- growable arrays ---- */
#define VEC_DECL(name, T) typedef struct { T *p; uint64_t n, cap; } name
VEC_DECL(vec_u64, uint64_t);
VEC_DECL(vec_call, ckm_call_t);
VEC_DECL(vec_hit, ckm_hit_t);
VEC_DECL(vec_otu, ckm_otu_t);
VEC_DECL(vec_best, ckm_best_t);
#define VEC_PUSH(v, x) do { if ((v).n == (v).cap) { (v).cap = (v).cap ? (v).cap * 2 : 64; \
        (v).p = realloc((v).p, (v).cap * sizeof *(v).p); } (v).p[(v).n++] = (x); } while (0)

typedef struct { /* KmerHit, kguts.h:154-163 */
    uint32_t oI, pos, fI;
    uint16_t avg;
    float wt;
} hit_t;

typedef struct {
    orc_out_t pub;
    vec_u64 call_off, hit_off, otu_off;
    vec_call calls;
    vec_hit hits;
    vec_otu otus;
    vec_best best;
} out_impl_t;

/* engine state of one KmerGuts (kguts.h:263-293) */
typedef struct {
    hit_t *hits; /* MAX_HITS_PER_SEQ */
    int num_hits;
    uint32_t current_fI;
    orc_params_t prm;
    /* per-sequence sinks */
    int want_calls, want_otu;
    vec_call *calls;
    vec_otu otu; /* std::map<int,int> kept sorted by key */
} guts_t;

static void otu_inc(guts_t *g, int32_t oI) {
    uint64_t i = 0;
    while (i < g->otu.n && g->otu.p[i].otu_index < oI) i++;
    if (i < g->otu.n && g->otu.p[i].otu_index == oI) { g->otu.p[i].count++; return; }
    ckm_otu_t e = {oI, 1};
    VEC_PUSH(g->otu, e);
    for (uint64_t j = g->otu.n - 1; j > i; j--) g->otu.p[j] = g->otu.p[j - 1];
    g->otu.p[i] = e;
}

/* ---- S2: kguts.cc:734-781 ---- */
static void process_set_of_hits(guts_t *g) {
    if (!g->want_otu && !g->want_calls) return; /* 737-738: state is NOT reset */
    int fI_count = 0, last_hit = 0;
    float weighted_hits = 0;
    for (int i = 0; i < g->num_hits; i++) {
        if (g->hits[i].fI == g->current_fI) {
            last_hit = i;
            fI_count++;
            weighted_hits += g->hits[i].wt;
        }
    }
    if (fI_count >= g->prm.min_hits && weighted_hits >= g->prm.min_weighted_hits) {
        if (g->want_calls) {
            ckm_call_t c = {g->hits[0].pos, g->hits[last_hit].pos + (CKM_KMER_SIZE - 1), fI_count, g->current_fI, weighted_hits};
            VEC_PUSH(*g->calls, c);
        }
        if (g->want_otu)
            for (int i = 0; i <= last_hit; i++)
                if (g->hits[i].fI == g->current_fI) otu_inc(g, (int32_t)g->hits[i].oI);
    }
    /* 772-780.  The reference reads hits[num_hits-2] even when num_hits < 2 (undefined behaviour,
     * only reachable with min_hits < 2); the restatement defines that case as "no carry". */
    if (g->num_hits >= 2 && g->hits[g->num_hits - 2].fI != g->current_fI &&
        g->hits[g->num_hits - 2].fI == g->hits[g->num_hits - 1].fI) {
        g->current_fI = g->hits[g->num_hits - 1].fI;
        g->hits[0] = g->hits[g->num_hits - 2];
        g->hits[1] = g->hits[g->num_hits - 1];
        g->num_hits = 2;
    } else {
        g->num_hits = 0;
    }
}

/* ---- E3: kguts.cc:682-732 ---- */
static void advance_past_ambig(const uint8_t **p, const uint8_t *bound) {
    int bad = 1;
    while (*p < bound && bad == 1) {
        bad = 0;
        for (int k = 7; k >= 0; k--) {
            if ((*p)[k] == 20) {
                bad = 1;
                *p += k + 1;
                break;
            }
        }
    }
}

/* ---- S1: kguts.cc:783-877, pointer walk and rolling update restated as written ---- */
static void gather_hits(const orc_table *t, guts_t *g, const uint8_t *pIseq, size_t len, vec_hit *hit_sink,
                        uint64_t *n_probes) {
    const uint8_t *p = pIseq;
    /* bound = pIseq + strlen - KMER_SIZE; for len < 8 the reference forms a pointer before the buffer
     * and every `p < bound` is false */
    if (len < CKM_KMER_SIZE) {
        if (g->num_hits >= g->prm.min_hits) process_set_of_hits(g);
        g->num_hits = 0;
        return;
    }
    const uint8_t *bound = pIseq + len - CKM_KMER_SIZE;
    advance_past_ambig(&p, bound);
    uint64_t encodedK = 0;
    if (p < bound) encodedK = encoded_kmer(p);
    while (p < bound) {
        int64_t where = orc_lookup_hash_entry(t, encodedK);
        (*n_probes)++;
        uint32_t pLoc = (uint32_t)(p - pIseq);
        if (where >= 0) {
            const ckm_sig_kmer_t *e = &t->slots[where];
            uint16_t avg_off_end = e->avg_from_end;
            uint32_t fI = (uint32_t)e->function_index;
            int32_t oI = e->otu_index;
            float f_wt = e->function_wt;
            if (hit_sink) { /* 814-815 */
                ckm_hit_t r;
                memset(&r, 0, sizeof r);
                r.which_kmer = e->which_kmer;
                r.offset = pLoc;
                r.otu_index = e->otu_index;
                r.function_index = e->function_index;
                r.function_wt = e->function_wt;
                r.avg_from_end = e->avg_from_end;
                VEC_PUSH(*hit_sink, r);
            }
            /* 821-831: unsigned int + int -> unsigned arithmetic */
            if (g->num_hits > 0 && (uint32_t)(g->hits[g->num_hits - 1].pos + (uint32_t)g->prm.max_gap) < pLoc) {
                if (g->num_hits >= g->prm.min_hits)
                    process_set_of_hits(g);
                else
                    g->num_hits = 0;
            }
            if (g->num_hits == 0) g->current_fI = fI; /* 833-836 */
            /* 838-842: (pLoc - last.pos) is unsigned int, (last.avg - avg) is int -> unsigned int
             * difference, converted to long for labs() */
            int ok = !g->prm.order_constraint || g->num_hits == 0;
            if (!ok) {
                const hit_t *l = &g->hits[g->num_hits - 1];
                uint32_t d = (pLoc - l->pos) - (uint32_t)((int)l->avg - (int)avg_off_end);
                ok = fI == l->fI && labs((long)d) <= 20;
            }
            if (ok) {
                hit_t *h = &g->hits[g->num_hits];
                h->oI = (uint32_t)oI;
                h->fI = fI;
                h->pos = pLoc;
                h->avg = avg_off_end;
                h->wt = f_wt;
                if (g->num_hits < CKM_MAX_HITS_PER_SEQ - 2) g->num_hits++; /* 850-851 */
                if (g->num_hits > 1 && g->current_fI != fI && g->hits[g->num_hits - 2].fI == g->hits[g->num_hits - 1].fI)
                    process_set_of_hits(g); /* 852-856 */
            }
        }
        p++;
        if (p < bound) { /* 859-871 */
            if (p[CKM_KMER_SIZE - 1] < 20) {
                encodedK = ((encodedK % CKM_CORE) * 20) + p[CKM_KMER_SIZE - 1];
            } else {
                p += CKM_KMER_SIZE;
                advance_past_ambig(&p, bound);
                if (p < bound) encodedK = encoded_kmer(p);
            }
        }
    }
    if (g->num_hits >= g->prm.min_hits) process_set_of_hits(g); /* 873-876 */
    g->num_hits = 0;
}

/* ---- B1: kguts.cc:1008-1199 ---- */
typedef struct {
    int fI;
    int count;
    float weighted;
} fscore_t;

/* libstdc++ std::partial_sort(first, first+2, last, weighted-desc) transcribed for a 2-element heap:
 * bits/stl_algo.h __heap_select + bits/stl_heap.h __make_heap/__adjust_heap/__push_heap/__sort_heap
 * (g++ 13).  The heap top (v[0]) is the SMALLER weight of the two kept. */
static void adjust_heap2(fscore_t *v, fscore_t value) {
    v[0] = v[1];
    if (v[0].weighted > value.weighted) { /* __push_heap: comp(first+parent, value) */
        v[1] = v[0];
        v[0] = value;
    } else {
        v[1] = value;
    }
}
static void partial_sort2(fscore_t *v, uint64_t n) {
    adjust_heap2(v, v[0]); /* __make_heap, len 2, parent 0 */
    for (uint64_t i = 2; i < n; i++) {
        if (v[i].weighted > v[0].weighted) { /* __pop_heap(first, middle, i) */
            fscore_t value = v[i];
            v[i] = v[0];
            adjust_heap2(v, value);
        }
    }
    fscore_t t0 = v[0]; /* __sort_heap on 2 elements swaps them */
    v[0] = v[1];
    v[1] = t0;
}

void orc_find_best_call(const ckm_call_t *calls, uint64_t n, ckm_best_t *out) {
    memset(out, 0, sizeof *out);
    out->function_index = -1;
    out->ambig_a = out->ambig_b = -1;
    if (n == 0) return; /* 1015-1018 */
    out->flags |= CKM_BEST_HAS_CALLS;
    ckm_call_t *collapsed = malloc(n * sizeof *collapsed), *merged = malloc(n * sizeof *merged);
    uint64_t nc = 0, nm = 0;
    /* 1023-1040: collapse adjacent calls with the same function */
    for (uint64_t i = 0; i < n;) {
        ckm_call_t cur = calls[i++];
        while (i < n && cur.function_index == calls[i].function_index) {
            cur.end = calls[i].end;
            cur.count += calls[i].count;
            cur.weighted_hits += calls[i].weighted_hits;
            i++;
        }
        collapsed[nc++] = cur;
    }
    /* 1063-1086: F1 | small F2 | F1 sandwich merge */
    for (uint64_t i = 0; i < nc;) {
        ckm_call_t cur = collapsed[i++];
        while (i < nc && i + 1 < nc && cur.function_index == collapsed[i + 1].function_index && collapsed[i].count < 5 &&
               cur.count + collapsed[i + 1].count >= 10) {
            cur.end = collapsed[i + 1].end;
            cur.count += collapsed[i + 1].count;
            cur.weighted_hits += collapsed[i + 1].weighted_hits;
            i += 2;
        }
        merged[nm++] = cur;
    }
    /* 1108-1128: std::map<int,FuncScore> (key is int: function_index converted) -> vector ascending */
    fscore_t *vec = malloc(nm * sizeof *vec);
    uint64_t nv = 0;
    for (uint64_t i = 0; i < nm; i++) {
        int fI = (int)merged[i].function_index;
        uint64_t j = 0;
        while (j < nv && vec[j].fI < fI) j++;
        if (j < nv && vec[j].fI == fI) {
            vec[j].count += merged[i].count;
            vec[j].weighted += merged[i].weighted_hits;
        } else {
            for (uint64_t k = nv; k > j; k--) vec[k] = vec[k - 1];
            vec[j].fI = fI;
            vec[j].count = merged[i].count;
            vec[j].weighted = merged[i].weighted_hits;
            nv++;
        }
    }
    if (nv > 1) partial_sort2(vec, nv); /* 1134-1139 */
    float score_offset = nv == 1 ? (float)vec[0].count : (float)(vec[0].count - vec[1].count); /* 1149-1152 */
    out->score_offset = score_offset;
    if (score_offset >= 5.0f) { /* 1156-1163 */
        out->function_index = vec[0].fI;
        out->score = (float)vec[0].count;
        out->weighted_score = vec[0].weighted;
    } else if (nv >= 2) { /* 1174-1196 */
        if (nv == 2) {
            out->flags |= CKM_BEST_AMBIG;
            out->ambig_a = vec[0].fI;
            out->ambig_b = vec[1].fI;
            out->score = (float)vec[0].count;
        } else {
            float pair_offset = (float)(vec[1].count - vec[2].count);
            if (pair_offset > 5.0f) {
                out->flags |= CKM_BEST_AMBIG;
                out->ambig_a = vec[0].fI;
                out->ambig_b = vec[1].fI;
                out->score = (float)vec[0].count;
                out->score_offset = pair_offset;
                out->weighted_score = vec[0].weighted;
            }
        }
    }
    free(collapsed);
    free(merged);
    free(vec);
}

/* ---- S4: kguts.cc:888-908 over a batch, results flattened like the product ---- */
static guts_t *guts_new(const orc_params_t *p) {
    guts_t *g = calloc(1, sizeof *g);
    g->hits = malloc(sizeof(hit_t) * CKM_MAX_HITS_PER_SEQ);
    g->prm = *p;
    return g;
}
static void guts_free(guts_t *g) {
    free(g->hits);
    free(g->otu.p);
    free(g);
}

orc_out_t *orc_call_batch(const orc_table *t, const orc_params_t *p, const char *residues, const uint64_t *offsets,
                          uint32_t n, uint32_t flags) {
    out_impl_t *o = calloc(1, sizeof *o);
    guts_t *g = guts_new(p);
    uint8_t *pIseq = NULL;
    size_t cap = 0;
    vec_call seq_calls = {0};
    uint64_t zero = 0, n_probes = 0;
    VEC_PUSH(o->call_off, zero);
    VEC_PUSH(o->hit_off, zero);
    VEC_PUSH(o->otu_off, zero);
    /* a run with calls==NULL and otu==NULL (hits only) makes process_set_of_hits a no-op: 737-738 */
    g->want_calls = (flags & (CKM_WANT_CALLS | CKM_WANT_BEST)) != 0;
    g->want_otu = (flags & CKM_WANT_OTU) != 0;
    g->calls = &seq_calls;
    for (uint32_t i = 0; i < n; i++) {
        size_t len = (size_t)(offsets[i + 1] - offsets[i]);
        const char *s = residues + offsets[i];
        if (len + 16 > cap) {
            cap = (len + 16) * 2;
            pIseq = realloc(pIseq, cap);
        }
        for (size_t k = 0; k < len; k++) pIseq[k] = orc_to_amino_acid_off(s[k]); /* 901-902 */
        /* the reference uses strlen(pseq): an embedded NUL truncates the scan (gather_hits 791) */
        size_t slen = strnlen(s, len);
        seq_calls.n = 0;
        g->otu.n = 0;
        g->num_hits = 0;
        gather_hits(t, g, pIseq, slen, (flags & CKM_WANT_HITS) ? &o->hits : NULL, &n_probes);
        if (flags & CKM_WANT_CALLS) {
            for (uint64_t k = 0; k < seq_calls.n; k++) VEC_PUSH(o->calls, seq_calls.p[k]);
            VEC_PUSH(o->call_off, o->calls.n);
        }
        if (flags & CKM_WANT_HITS) VEC_PUSH(o->hit_off, o->hits.n);
        if (flags & CKM_WANT_OTU) {
            for (uint64_t k = 0; k < g->otu.n; k++) VEC_PUSH(o->otus, g->otu.p[k]);
            VEC_PUSH(o->otu_off, o->otus.n);
        }
        if (flags & CKM_WANT_BEST) {
            ckm_best_t b;
            orc_find_best_call(seq_calls.p, seq_calls.n, &b);
            VEC_PUSH(o->best, b);
        }
    }
    o->pub.o.n = n;
    if (flags & CKM_WANT_CALLS) { o->pub.o.call_offsets = o->call_off.p; o->pub.o.calls = o->calls.p; }
    if (flags & CKM_WANT_HITS) { o->pub.o.hit_offsets = o->hit_off.p; o->pub.o.hits = o->hits.p; }
    if (flags & CKM_WANT_OTU) { o->pub.o.otu_offsets = o->otu_off.p; o->pub.o.otus = o->otus.p; }
    if (flags & CKM_WANT_BEST) o->pub.o.best = o->best.p;
    o->pub.o.n_probes = n_probes;
    o->pub.o.n_hits = o->hits.n;
    free(pIseq);
    free(seq_calls.p);
    guts_free(g);
    return &o->pub;
}

void orc_out_free(orc_out_t *pub) {
    out_impl_t *o = (out_impl_t *)pub;
    if (!o) return;
    free(o->call_off.p); free(o->hit_off.p); free(o->otu_off.p);
    free(o->calls.p); free(o->hits.p); free(o->otus.p); free(o->best.p);
    free(o);
}

/* ---- F1/F2: family voting ---- */
struct orc_family {
    uint64_t n_kmers;
    uint64_t *kmers;    /* sorted */
    uint64_t *perm_off; /* n_kmers: offset of the list of sorted k-mer i */
    uint32_t *perm_cnt;
    uint32_t *fam_ids;
    uint32_t n_fams, n_pgf, n_functions, hypo_sid;
    uint32_t *fam_func_sid, *fam_pgf, *func_sid;
};

typedef struct { uint64_t k, off; uint32_t cnt; } kent_t;
static int kent_cmp(const void *a, const void *b) {
    const kent_t *x = a, *y = b;
    return x->k < y->k ? -1 : x->k > y->k;
}
static void *dupmem(const void *p, size_t n) {
    void *q = malloc(n ? n : 1);
    if (n) memcpy(q, p, n);
    return q;
}

orc_family *orc_family_new(uint64_t n_kmers, const uint64_t *kmers, const uint64_t *fam_off, const uint32_t *fam_ids,
                           uint32_t n_fams, const uint32_t *fam_func_sid, const uint32_t *fam_pgf, uint32_t n_pgf,
                           uint32_t n_functions, const uint32_t *func_sid, uint32_t hypo_sid) {
    orc_family *f = calloc(1, sizeof *f);
    kent_t *e = malloc(sizeof *e * (n_kmers ? n_kmers : 1));
    for (uint64_t i = 0; i < n_kmers; i++) e[i] = (kent_t){kmers[i], fam_off[i], (uint32_t)(fam_off[i + 1] - fam_off[i])};
    qsort(e, n_kmers, sizeof *e, kent_cmp);
    f->n_kmers = n_kmers;
    f->kmers = malloc(8 * (n_kmers ? n_kmers : 1));
    f->perm_off = malloc(8 * (n_kmers ? n_kmers : 1));
    f->perm_cnt = malloc(4 * (n_kmers ? n_kmers : 1));
    for (uint64_t i = 0; i < n_kmers; i++) { f->kmers[i] = e[i].k; f->perm_off[i] = e[i].off; f->perm_cnt[i] = e[i].cnt; }
    free(e);
    f->fam_ids = dupmem(fam_ids, 4 * (n_kmers ? fam_off[n_kmers] : 0));
    f->n_fams = n_fams; f->n_pgf = n_pgf; f->n_functions = n_functions; f->hypo_sid = hypo_sid;
    f->fam_func_sid = dupmem(fam_func_sid, 4 * (size_t)n_fams);
    f->fam_pgf = dupmem(fam_pgf, 4 * (size_t)n_fams);
    f->func_sid = dupmem(func_sid, 4 * (size_t)n_functions);
    return f;
}

void orc_family_free(orc_family *f) {
    if (!f) return;
    free(f->kmers); free(f->perm_off); free(f->perm_cnt); free(f->fam_ids);
    free(f->fam_func_sid); free(f->fam_pgf); free(f->func_sid); free(f);
}

static int64_t fam_find(const orc_family *f, uint64_t k) {
    uint64_t lo = 0, hi = f->n_kmers;
    while (lo < hi) {
        uint64_t mid = (lo + hi) / 2;
        if (f->kmers[mid] < k) lo = mid + 1; else hi = mid;
    }
    return (lo < f->n_kmers && f->kmers[lo] == k) ? (int64_t)lo : -1;
}

void orc_family_batch(const orc_table *t, const orc_params_t *p, const orc_family *f, const char *residues,
                      const uint64_t *offsets, uint32_t n, ckm_family_match_t *out) {
    /* sequence_accumulated_score_t per family (family_mapper.h:33-49), dense + touched list */
    uint32_t *hit_total = calloc(f->n_fams ? f->n_fams : 1, 4);
    float *weighted = calloc(f->n_fams ? f->n_fams : 1, 4);
    uint8_t *seen = calloc(f->n_fams ? f->n_fams : 1, 1);
    float *rollup = calloc(f->n_pgf ? f->n_pgf : 1, 4);
    uint8_t *pseen = calloc(f->n_pgf ? f->n_pgf : 1, 1);
    uint32_t flags = CKM_WANT_CALLS | CKM_WANT_HITS | CKM_WANT_BEST;
    for (uint32_t i = 0; i < n; i++) {
        /* ingest_protein (46-63): process_aa_seq with calls + on_hit callback, no OTU stats */
        uint64_t off2[2] = {0, offsets[i + 1] - offsets[i]};
        orc_out_t *o = orc_call_batch(t, p, residues + offsets[i], off2, 1, flags);
        for (uint64_t h = 0; h < o->o.hit_offsets[1]; h++) { /* on_hit, 287-330 (family mode) */
            int64_t ki = fam_find(f, o->o.hits[h].which_kmer);
            if (ki < 0) continue;
            uint32_t cnt = f->perm_cnt[ki];
            float weight = 1.0f / (float)cnt;
            for (uint32_t e = 0; e < cnt; e++) {
                uint32_t fam = f->fam_ids[f->perm_off[ki] + e];
                if (fam >= f->n_fams) continue; /* no family_data_ entry: skipped at 146-148 */
                seen[fam] = 1;
                hit_total[fam]++;
                weighted[fam] += weight;
            }
        }
        const ckm_best_t *b = &o->o.best[0];
        /* 98-123: empty or ambiguous function -> "hypothetical protein" (allow_ambiguous_functions_ is false) */
        uint32_t sid = f->hypo_sid;
        int32_t fidx = -1;
        if (b->function_index >= 0 && (uint32_t)b->function_index < f->n_functions) {
            sid = f->func_sid[b->function_index];
            fidx = b->function_index;
        }
        ckm_family_match_t m = {-1, -1, 0.0f, 0.0f, b->score, fidx};
        /* 137-178, visiting families in ascending id instead of unordered_map order */
        for (uint32_t fam = 0; fam < f->n_fams; fam++) {
            if (!seen[fam]) continue;
            if (hit_total[fam] >= 3 && f->fam_func_sid[fam] == sid) {
                rollup[f->fam_pgf[fam]] += weighted[fam];
                pseen[f->fam_pgf[fam]] = 1;
                if (weighted[fam] > m.lfam_score) {
                    m.lfam_score = weighted[fam];
                    m.lfam = (int32_t)fam;
                }
            }
        }
        for (uint32_t g = 0; g < f->n_pgf; g++) { /* 187-197 */
            if (!pseen[g]) continue;
            if (rollup[g] > m.gfam_score) {
                m.gfam_score = rollup[g];
                m.gfam = (int32_t)g;
            }
            rollup[g] = 0;
            pseen[g] = 0;
        }
        for (uint32_t fam = 0; fam < f->n_fams; fam++)
            if (seen[fam]) { seen[fam] = 0; hit_total[fam] = 0; weighted[fam] = 0; }
        out[i] = m;
        orc_out_free(o);
    }
    free(hit_total); free(weighted); free(seen); free(rollup); free(pseen);
}

/* LookupRequest::on_hit in family mode (lookup_request.cc:441-464): per sequence, every family touched with hit_count and
 * weighted_total (f32 sum of 1/|list| in hit order), ascending family id; *scores / *score_off are malloc'ed. */
void orc_family_scores(const orc_table *t, const orc_params_t *p, const orc_family *f, const char *residues,
                       const uint64_t *offsets, uint32_t n, ckm_score_t **scores, uint64_t **score_off) {
    uint32_t *hit_count = calloc(f->n_fams ? f->n_fams : 1, 4);
    float *weighted = calloc(f->n_fams ? f->n_fams : 1, 4);
    uint64_t cap = 1024, ns = 0;
    ckm_score_t *out = malloc(cap * sizeof *out);
    uint64_t *off = malloc(((size_t)n + 1) * 8);
    off[0] = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint64_t off2[2] = {0, offsets[i + 1] - offsets[i]};
        orc_out_t *o = orc_call_batch(t, p, residues + offsets[i], off2, 1, CKM_WANT_HITS);
        for (uint64_t h = 0; h < o->o.hit_offsets[1]; h++) {
            int64_t ki = fam_find(f, o->o.hits[h].which_kmer);
            if (ki < 0) continue;
            uint32_t cnt = f->perm_cnt[ki];
            float weight = 1.0f / (float)cnt;
            for (uint32_t e = 0; e < cnt; e++) {
                uint32_t fam = f->fam_ids[f->perm_off[ki] + e];
                if (fam >= f->n_fams) continue;
                hit_count[fam]++;
                weighted[fam] += weight;
            }
        }
        for (uint32_t fam = 0; fam < f->n_fams; fam++) {
            if (!hit_count[fam]) continue;
            if (ns == cap) out = realloc(out, (cap *= 2) * sizeof *out);
            out[ns++] = (ckm_score_t){fam, hit_count[fam], weighted[fam]};
            hit_count[fam] = 0;
            weighted[fam] = 0;
        }
        off[i + 1] = ns;
        orc_out_free(o);
    }
    free(hit_count); free(weighted);
    *scores = out;
    *score_off = off;
}

/* ---- D1: TranslationTable, trans_table.cc:8-63.  NCBI genetic code 11 listed in TCAG order; the table is
 * re-indexed by encode_triple (A=0,C=1,G=2,T/U=3 -> 16*b1+4*b2+b3), slot 64 = 'X' for ambiguous codons ---- */
static const char NCBI11_AAS[65] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
static char AA11[65];
static int aa11_ready = 0;
static int encode_char(char c) { /* trans_table.h:47-68 */
    switch (c) {
        case 'a': case 'A': return 0;
        case 'c': case 'C': return 1;
        case 'g': case 'G': return 2;
        case 't': case 'u': case 'T': case 'U': return 3;
        default: return 4;
    }
}
static void aa11_init(void) {
    if (aa11_ready) return;
    const char order[4] = {'T', 'C', 'A', 'G'};
    for (int pos = 0; pos < 64; pos++) {
        int e1 = encode_char(order[pos / 16]), e2 = encode_char(order[(pos / 4) % 4]), e3 = encode_char(order[pos % 4]);
        AA11[e1 * 16 + e2 * 4 + e3] = NCBI11_AAS[pos];
    }
    AA11[64] = 'X';
    aa11_ready = 1;
}
/* dna_seq.h:28-111 */
static char complement(char c) {
    switch (c) {
        case 'a': return 't'; case 'A': return 'T';
        case 'c': return 'g'; case 'C': return 'G';
        case 'g': return 'c'; case 'G': return 'C';
        case 't': case 'u': return 'a';
        case 'T': case 'U': return 'A';
        case 'm': return 'k'; case 'M': return 'K';
        case 'r': return 'y'; case 'R': return 'Y';
        case 'w': return 'w'; case 'W': return 'W';
        case 's': return 'S'; case 'S': return 'S';
        case 'y': return 'r'; case 'Y': return 'R';
        case 'k': return 'm'; case 'K': return 'M';
        case 'b': return 'v'; case 'B': return 'V';
        case 'd': return 'h'; case 'D': return 'H';
        case 'h': return 'd'; case 'H': return 'D';
        case 'v': return 'b'; case 'V': return 'B';
        case 'n': return 'n'; case 'N': return 'N';
        default: return c;
    }
}
/* D2: get_translated_frame (dna_seq.cc:25-37) over seq() or reverse_seq() (39-47), translate (trans_table.cc:65-84) */
size_t orc_translate_frame(const char *dna, size_t len, int frame, char *out) {
    aa11_init();
    size_t off = (size_t)(frame < 0 ? -frame : frame) - 1, n = 0;
    for (size_t i = off; i + 3 <= len; i += 3) {
        char c[3];
        for (int k = 0; k < 3; k++) c[k] = frame > 0 ? dna[i + k] : complement(dna[len - 1 - (i + k)]);
        int e1 = encode_char(c[0]), e2 = encode_char(c[1]), e3 = encode_char(c[2]);
        out[n++] = AA11[(e1 < 4 && e2 < 4 && e3 < 4) ? e1 * 16 + e2 * 4 + e3 : 64];
    }
    out[n] = 0;
    return n;
}

typedef struct {
    orc_fq_out_t pub;
    int32_t *best_frame;
    double *best_score;
    uint64_t *match_off;
    ckm_fq_match_t *matches;
} fq_impl_t;

/* D3: fq_process_request.cc:298-365.  boost::split(..., "*", token_compress_on) (dna_seq.cc:17) yields the
 * maximal stop-free runs (plus empty tokens that the length filter drops). */
orc_fq_out_t *orc_fq_batch(const orc_table *t, const orc_params_t *p, const orc_family *f, const char *bases,
                           const uint64_t *offsets, uint32_t n) {
    static const int FRAMES[6] = {1, 2, 3, -1, -2, -3};
    fq_impl_t *o = calloc(1, sizeof *o);
    o->best_frame = calloc(n ? n : 1, sizeof(int32_t));
    o->best_score = calloc(n ? n : 1, sizeof(double));
    o->match_off = calloc((size_t)n + 1, 8);
    uint64_t cap = 64, nm = 0, nfrag = 0;
    o->matches = malloc(cap * sizeof *o->matches);
    ckm_fq_match_t *cur = NULL, *best = NULL;
    size_t cur_cap = 0;
    for (uint32_t r = 0; r < n; r++) {
        const char *dna = bases + offsets[r];
        size_t len = (size_t)(offsets[r + 1] - offsets[r]);
        char *prot = malloc(len / 3 + 2);
        if (len / 3 + 1 > cur_cap) {
            cur_cap = len / 3 + 1;
            cur = realloc(cur, cur_cap * sizeof *cur);
            best = realloc(best, cur_cap * sizeof *best);
        }
        double best_score = 0.0;
        int best_frame = 0;
        size_t best_n = 0;
        for (int fi = 0; fi < 6; fi++) {
            size_t na = orc_translate_frame(dna, len, FRAMES[fi], prot);
            double score = 0.0;
            size_t ncur = 0;
            for (size_t s = 0; s <= na;) {
                size_t e = s;
                while (e < na && prot[e] != '*') e++;
                if (e - s > 10) { /* prot.length() > 10 */
                    uint64_t off2[2] = {0, e - s};
                    ckm_family_match_t m;
                    orc_family_batch(t, p, f, prot + s, off2, 1, &m);
                    cur[ncur].length = (uint32_t)(e - s);
                    cur[ncur].m = m;
                    ncur++;
                    nfrag++;
                    score += m.score;
                }
                if (score > best_score) { /* 340-346, inside the fragment loop */
                    best_score = score;
                    best_frame = FRAMES[fi];
                    best_n = ncur;
                    memcpy(best, cur, ncur * sizeof *cur);
                }
                s = e + 1;
            }
        }
        o->best_frame[r] = best_score > 0.0 ? best_frame : 0;
        o->best_score[r] = best_score;
        if (best_score > 0.0) {
            while (nm + best_n > cap) { cap *= 2; o->matches = realloc(o->matches, cap * sizeof *o->matches); }
            memcpy(o->matches + nm, best, best_n * sizeof *best);
            nm += best_n;
        }
        o->match_off[r + 1] = nm;
        free(prot);
    }
    free(cur);
    free(best);
    o->pub.o.n = n;
    o->pub.o.best_frame = o->best_frame;
    o->pub.o.best_score = o->best_score;
    o->pub.o.match_offsets = o->match_off;
    o->pub.o.matches = o->matches;
    o->pub.o.n_fragments = nfrag;
    return &o->pub;
}

void orc_fq_out_free(orc_fq_out_t *pub) {
    fq_impl_t *o = (fq_impl_t *)pub;
    if (!o) return;
    free(o->best_frame); free(o->best_score); free(o->match_off); free(o->matches); free(o);
}

/* ---- M1: postings + pairwise counts ---- */
typedef struct { uint64_t k; uint32_t e; } post_t;
struct orc_postings {
    post_t *p;
    uint64_t n, cap;
    int sorted;
};
orc_postings *orc_postings_new(void) { return calloc(1, sizeof(orc_postings)); }
void orc_postings_free(orc_postings *p) { if (p) { free(p->p); free(p); } }
uint64_t orc_postings_count(const orc_postings *p) { return p->n; }
void orc_free(void *p) { free(p); }

/* add_request.cc:133 (process_aa_seq_hits) + 164-170: one add_mapping(enc_id, hit.which_kmer) per hit */
void orc_postings_add(const orc_table *t, const orc_params_t *prm, orc_postings *p, const uint32_t *eids,
                      const char *residues, const uint64_t *offsets, uint32_t n) {
    orc_out_t *o = orc_call_batch(t, prm, residues, offsets, n, CKM_WANT_HITS);
    for (uint32_t i = 0; i < n; i++)
        for (uint64_t h = o->o.hit_offsets[i]; h < o->o.hit_offsets[i + 1]; h++) {
            post_t e = {o->o.hits[h].which_kmer, eids[i]};
            VEC_PUSH(*p, e);
        }
    p->sorted = 0;
    orc_out_free(o);
}
/* the postings as parallel arrays, and back: what a rank contributes to / receives from the all-gather of the multi-GPU /matrix */
void orc_postings_export(const orc_postings *p, uint64_t *keys, uint32_t *eids) {
    for (uint64_t i = 0; i < p->n; i++) { keys[i] = p->p[i].k; eids[i] = p->p[i].e; }
}
void orc_postings_import(orc_postings *p, const uint64_t *keys, const uint32_t *eids, uint64_t n) {
    p->n = 0;
    for (uint64_t i = 0; i < n; i++) { post_t e = {keys[i], eids[i]}; VEC_PUSH(*p, e); }
    p->sorted = 0;
}
static int post_cmp(const void *a, const void *b) {
    const post_t *x = a, *y = b;
    return x->k < y->k ? -1 : x->k > y->k;
}
static int u64_cmp(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* ---- N2: family tables from families.nr proteins.  nr_loader.cc:131-202 (thread_load, family mode) queues
 * (hit.which_kmer, fam_id) for every hit of process_aa_seq(id, seq, 0, hit_cb, 0); kmer_inserter.cc:36-58 applies
 * add_fam_mapping (kmer.cc:244-268), whose fam_map_insert (216-230) keeps a family once per k-mer.  A sequence with
 * no family (fam id 0xFFFFFFFF) ends its chunk (the return at nr_loader.cc:159).  The pairs accumulate in an
 * orc_postings; orc_family_nr_table returns them as CSR sorted by (k-mer, family id), each pair once. ---- */
void orc_family_nr_add(const orc_table *t, const orc_params_t *prm, orc_postings *p, const uint32_t *fam_ids,
                       const char *residues, const uint64_t *offsets, uint32_t n) {
    uint32_t m = 0;
    while (m < n && fam_ids[m] != 0xffffffffu) m++;
    orc_postings_add(t, prm, p, fam_ids, residues, offsets, m);
}
static int post_cmp2(const void *a, const void *b) {
    const post_t *x = a, *y = b;
    if (x->k != y->k) return x->k < y->k ? -1 : 1;
    return x->e < y->e ? -1 : x->e > y->e;
}
/* malloc'ed outputs: kmers[*n_kmers], fam_off[*n_kmers + 1], ids[*n_entries] */
void orc_family_nr_table(orc_postings *p, uint64_t *n_kmers, uint64_t *n_entries, uint64_t **kmers, uint64_t **fam_off,
                         uint32_t **ids) {
    qsort(p->p, p->n, sizeof(post_t), post_cmp2);
    p->sorted = 0;
    uint64_t *K = malloc((p->n + 1) * 8), *O = malloc((p->n + 2) * 8);
    uint32_t *I = malloc((p->n + 1) * 4);
    uint64_t nk = 0, ne = 0;
    for (uint64_t i = 0; i < p->n; i++) {
        if (i && p->p[i].k == p->p[i - 1].k && p->p[i].e == p->p[i - 1].e) continue;
        if (!i || p->p[i].k != p->p[i - 1].k) {
            K[nk] = p->p[i].k;
            O[nk++] = ne;
        }
        I[ne++] = p->p[i].e;
    }
    O[nk] = ne;
    *n_kmers = nk;
    *n_entries = ne;
    *kmers = K;
    *fam_off = O;
    *ids = I;
}

ckm_pair_t *orc_matrix_rows(const orc_table *t, const orc_params_t *prm, orc_postings *p, const uint32_t *eids,
                            const char *residues, const uint64_t *offsets, uint32_t n, uint32_t row_begin,
                            uint32_t row_end, uint64_t *n_pairs) {
    if (!p->sorted) { qsort(p->p, p->n, sizeof *p->p, post_cmp); p->sorted = 1; }
    /* matrix_proteins_: an id is a member from the first time it is set (matrix_request.cc:90) */
    uint32_t max_e = 0;
    for (uint32_t i = 0; i < n; i++) if (eids[i] > max_e) max_e = eids[i];
    uint32_t *first = malloc(((size_t)max_e + 1) * 4);
    for (uint32_t e = 0; e <= max_e; e++) first[e] = 0xffffffffu;
    for (uint32_t i = 0; i < n; i++) if (first[eids[i]] == 0xffffffffu) first[eids[i]] = i;
    struct { uint64_t *p; uint64_t n, cap; } contrib = {0};
    if (row_end > n) row_end = n;
    for (uint32_t i = row_begin; i < row_end; i++) {
        uint64_t off2[2] = {0, offsets[i + 1] - offsets[i]};
        orc_out_t *o = orc_call_batch(t, prm, residues + offsets[i], off2, 1, CKM_WANT_HITS); /* calls=0, otu=0: 92-94 */
        for (uint64_t h = 0; h < o->o.hit_offsets[1]; h++) { /* on_hit, 130-161 */
            uint64_t k = o->o.hits[h].which_kmer, lo = 0, hi = p->n;
            while (lo < hi) { uint64_t mid = (lo + hi) / 2; if (p->p[mid].k < k) lo = mid + 1; else hi = mid; }
            for (; lo < p->n && p->p[lo].k == k; lo++) {
                uint32_t e = p->p[lo].e;
                if (e != eids[i] && e <= max_e && first[e] <= i) {
                    uint64_t key = ((uint64_t)eids[i] << 32) | e;
                    VEC_PUSH(contrib, key);
                }
            }
        }
        orc_out_free(o);
    }
    qsort(contrib.p, contrib.n, 8, u64_cmp);
    ckm_pair_t *out = malloc((contrib.n ? contrib.n : 1) * sizeof *out);
    uint64_t np = 0;
    for (uint64_t x = 0; x < contrib.n;) {
        uint64_t y = x;
        while (y < contrib.n && contrib.p[y] == contrib.p[x]) y++;
        out[np++] = (ckm_pair_t){(uint32_t)(contrib.p[x] >> 32), (uint32_t)contrib.p[x], y - x};
        x = y;
    }
    free(contrib.p);
    free(first);
    *n_pairs = np;
    return out;
}

/* ---- CPU-baseline timing loop (bench.py cpu_baseline "port" leg) ---- */
typedef struct {
    const orc_table *t;
    const orc_params_t *p;
    const char *residues;
    const uint64_t *offsets;
    uint32_t lo, hi;
    int want_best;
    uint64_t ncalls;
} bench_arg_t;

static void *bench_worker(void *av) {
    bench_arg_t *a = av;
    guts_t *g = guts_new(a->p);
    vec_call seq_calls = {0};
    g->want_calls = 1;
    g->calls = &seq_calls;
    uint8_t *pIseq = NULL;
    size_t cap = 0;
    uint64_t probes = 0, c = 0;
    for (uint32_t i = a->lo; i < a->hi; i++) {
        size_t len = (size_t)(a->offsets[i + 1] - a->offsets[i]);
        const char *s = a->residues + a->offsets[i];
        if (len + 16 > cap) { cap = (len + 16) * 2; pIseq = realloc(pIseq, cap); }
        for (size_t k = 0; k < len; k++) pIseq[k] = orc_to_amino_acid_off(s[k]);
        seq_calls.n = 0;
        g->num_hits = 0;
        gather_hits(a->t, g, pIseq, len, NULL, &probes);
        c += seq_calls.n;
        if (a->want_best) {
            ckm_best_t b;
            orc_find_best_call(seq_calls.p, seq_calls.n, &b);
            c += b.function_index >= 0;
        }
    }
    a->ncalls = c;
    free(pIseq);
    free(seq_calls.p);
    guts_free(g);
    return NULL;
}

double orc_bench_calls(const orc_table *t, const orc_params_t *p, const char *residues, const uint64_t *offsets,
                       uint32_t n, int want_best, int threads, uint64_t *total_calls) {
    if (threads < 1) threads = 1;
    pthread_t *th = malloc(sizeof *th * threads);
    bench_arg_t *args = calloc(threads, sizeof *args);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < threads; k++) {
        args[k] = (bench_arg_t){t, p, residues, offsets, (uint32_t)((uint64_t)n * k / threads),
                                (uint32_t)((uint64_t)n * (k + 1) / threads), want_best, 0};
        pthread_create(&th[k], NULL, bench_worker, &args[k]);
    }
    uint64_t tot = 0;
    for (int k = 0; k < threads; k++) {
        pthread_join(th[k], NULL);
        tot += args[k].ncalls;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (total_calls) *total_calls = tot;
    free(th);
    free(args);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

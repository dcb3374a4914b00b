/*
 * oxli_oracle.c -- CPU restatement of the oxli k-mer counting hot path.
 * TEST INFRASTRUCTURE ONLY (see oxli_oracle.h).  Parity status: PINNED.
 *
 * Each function cites the reference lines (relative to /root/reference) whose
 * behaviour it restates.  The hash itself is the public MurmurHash3_x64_128
 * algorithm (crate murmurhash3 0.0.5, reached through sourmash 0.23.0
 * `_hash_murmur`; neither crate is vendored in the reference tree).
 */
#include "oxli_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------ murmur */

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

static inline uint64_t load_le64(const uint8_t *p) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}

void oxo_murmur3_x64_128(const uint8_t *data, size_t len, uint32_t seed, uint64_t out[2]) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    const size_t nblocks = len / 16;
    for (size_t b = 0; b < nblocks; ++b) {
        uint64_t k1 = load_le64(data + 16 * b), k2 = load_le64(data + 16 * b + 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + 16 * nblocks;
    const size_t rem = len & 15;
    uint64_t k1 = 0, k2 = 0;
    for (size_t i = rem; i > 8; --i) k2 = (k2 << 8) | tail[i - 1];
    if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
    for (size_t i = (rem < 8 ? rem : 8); i > 0; --i) k1 = (k1 << 8) | tail[i - 1];
    if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

/* SMHasher VerificationTest: hash keys {0}, {0,1}, ... with seed 256-i, then
 * hash the concatenated results with seed 0; first 4 bytes LE. */
uint32_t oxo_murmur3_smhasher_verification(void) {
    uint8_t key[256], hashes[256 * 16], final[16];
    uint64_t h[2];
    for (int i = 0; i < 256; ++i) {
        key[i] = (uint8_t)i;
        oxo_murmur3_x64_128(key, (size_t)i, (uint32_t)(256 - i), h);
        for (int b = 0; b < 8; ++b) {
            hashes[i * 16 + b] = (uint8_t)(h[0] >> (8 * b));
            hashes[i * 16 + 8 + b] = (uint8_t)(h[1] >> (8 * b));
        }
    }
    oxo_murmur3_x64_128(hashes, sizeof hashes, 0, h);
    for (int b = 0; b < 8; ++b) final[b] = (uint8_t)(h[0] >> (8 * b));
    return (uint32_t)final[0] | ((uint32_t)final[1] << 8) | ((uint32_t)final[2] << 16) |
           ((uint32_t)final[3] << 24);
}

/* ------------------------------------------------------- sequence -> hashes */

static inline uint8_t upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
static inline int is_acgt(uint8_t c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
static inline uint8_t complement(uint8_t c) {
    switch (c) {
    case 'A': return 'T';
    case 'T': return 'A';
    case 'C': return 'G';
    case 'G': return 'C';
    default: return c; /* only ever read inside all-ACGT windows */
    }
}

/* One window: the canonical k-mer is the bytewise-smaller of fw and rc
 * (sourmash SeqToHashes::next: `std::cmp::min(kmer, krc)`), hashed with seed
 * 42, keeping h1 (`_hash_murmur`).  src/lib.rs:69-79 and 576-584 pass
 * HashFunctions::Murmur64Dna, seed 42. */
static inline uint64_t window_hash(const uint8_t *fw, const uint8_t *rc, uint32_t k) {
    uint64_t h[2];
    oxo_murmur3_x64_128(memcmp(fw, rc, k) <= 0 ? fw : rc, k, 42, h);
    return h[0];
}

int oxo_hash_kmer(const uint8_t *kmer, size_t len, uint32_t ksize, uint64_t *out) {
    /* src/lib.rs:66-67: length check first */
    if (len != ksize) return OXO_ERR_WRONG_KSIZE;
    uint8_t fw[256], rc[256];
    for (uint32_t i = 0; i < ksize; ++i) {
        fw[i] = upper(kmer[i]);
        if (!is_acgt(fw[i])) return OXO_ERR_BAD_KMER; /* force=false -> Err (src/lib.rs:79) */
    }
    for (uint32_t i = 0; i < ksize; ++i) rc[i] = complement(fw[ksize - 1 - i]);
    *out = window_hash(fw, rc, ksize);
    return OXO_OK;
}

/* Shared driver.  Mirrors the reference flow of src/lib.rs:576-600: upper-case
 * copy, whole-sequence reverse complement, then one iterator step per window
 * left to right.  `sink` semantics chosen by the callers below. */
typedef struct {
    uint8_t *up, *rc;
    size_t cap;
} scratch_t;

static int scratch_reserve(scratch_t *s, size_t len) {
    if (len <= s->cap) return 0;
    size_t ncap = len + len / 2 + 64;
    uint8_t *a = (uint8_t *)realloc(s->up, ncap);
    if (!a) return -1;
    s->up = a;
    uint8_t *b = (uint8_t *)realloc(s->rc, ncap);
    if (!b) return -1;
    s->rc = b;
    s->cap = ncap;
    return 0;
}

static void prepare(scratch_t *s, const uint8_t *seq, size_t len) {
    for (size_t i = 0; i < len; ++i) s->up[i] = upper(seq[i]);
    for (size_t i = 0; i < len; ++i) s->rc[i] = complement(s->up[len - 1 - i]);
}

void oxo_hash_windows(const uint8_t *seq, size_t len, uint32_t ksize, uint64_t *out) {
    if (ksize == 0 || len < ksize) return;
    scratch_t s = {0, 0, 0};
    if (scratch_reserve(&s, len)) return;
    prepare(&s, seq, len);
    int64_t last_bad = -1;
    for (uint32_t j = 0; j + 1 < ksize; ++j)
        if (!is_acgt(s.up[j])) last_bad = j;
    for (size_t i = 0; i + ksize <= len; ++i) {
        if (!is_acgt(s.up[i + ksize - 1])) last_bad = (int64_t)(i + ksize - 1);
        out[i] = (last_bad >= (int64_t)i) ? 0 : window_hash(s.up + i, s.rc + (len - ksize - i), ksize);
    }
    free(s.up);
    free(s.rc);
}

/* -------------------------------------------------------------------- table */

struct oxo_table {
    uint64_t *keys, *vals;
    uint8_t *state; /* 0 empty, 1 live, 2 deleted */
    uint64_t cap, live, used; /* used = live + deleted */
};

static inline uint64_t mix(uint64_t x) {
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

static int table_alloc(oxo_table *t, uint64_t cap) {
    t->keys = (uint64_t *)malloc(cap * sizeof(uint64_t));
    t->vals = (uint64_t *)malloc(cap * sizeof(uint64_t));
    t->state = (uint8_t *)calloc(cap, 1);
    t->cap = cap; t->live = 0; t->used = 0;
    return (t->keys && t->vals && t->state) ? 0 : -1;
}

oxo_table *oxo_table_new(void) {
    oxo_table *t = (oxo_table *)calloc(1, sizeof *t);
    if (!t) return NULL;
    if (table_alloc(t, 1024)) { oxo_table_free(t); return NULL; }
    return t;
}

void oxo_table_free(oxo_table *t) {
    if (!t) return;
    free(t->keys); free(t->vals); free(t->state); free(t);
}

void oxo_table_clear(oxo_table *t) {
    memset(t->state, 0, t->cap);
    t->live = t->used = 0;
}

uint64_t oxo_table_len(const oxo_table *t) { return t->live; }

static uint64_t *find_or_insert(oxo_table *t, uint64_t h, int *fresh);

static void table_grow(oxo_table *t) {
    oxo_table old = *t;
    uint64_t ncap = old.cap;
    while (old.live * 2 >= ncap) ncap *= 2; /* also purges tombstones at same cap */
    if (table_alloc(t, ncap)) abort();
    for (uint64_t i = 0; i < old.cap; ++i)
        if (old.state[i] == 1) { int f; *find_or_insert(t, old.keys[i], &f) = old.vals[i]; }
    free(old.keys); free(old.vals); free(old.state);
}

static uint64_t *find_or_insert(oxo_table *t, uint64_t h, int *fresh) {
    if ((t->used + 1) * 10 > t->cap * 7) table_grow(t);
    const uint64_t mask = t->cap - 1;
    uint64_t i = mix(h) & mask, tomb = UINT64_MAX;
    for (;; i = (i + 1) & mask) {
        if (t->state[i] == 0) break;
        if (t->state[i] == 1 && t->keys[i] == h) { *fresh = 0; return &t->vals[i]; }
        if (t->state[i] == 2 && tomb == UINT64_MAX) tomb = i;
    }
    if (tomb != UINT64_MAX) i = tomb; else t->used++;
    t->state[i] = 1; t->keys[i] = h; t->vals[i] = 0; t->live++;
    *fresh = 1;
    return &t->vals[i];
}

static int64_t find(const oxo_table *t, uint64_t h) {
    const uint64_t mask = t->cap - 1;
    for (uint64_t i = mix(h) & mask;; i = (i + 1) & mask) {
        if (t->state[i] == 0) return -1;
        if (t->state[i] == 1 && t->keys[i] == h) return (int64_t)i;
    }
}

/* src/lib.rs:100-104  entry(h).or_insert(0) += 1, returns the new count */
uint64_t oxo_table_count_hash(oxo_table *t, uint64_t h) {
    int f;
    uint64_t *v = find_or_insert(t, h, &f);
    return ++*v;
}

void oxo_table_add_hash(oxo_table *t, uint64_t h, uint64_t c) {
    int f;
    *find_or_insert(t, h, &f) += c;
}

/* src/lib.rs:185-188  missing -> 0 */
uint64_t oxo_table_get_hash(const oxo_table *t, uint64_t h) {
    int64_t i = find(t, h);
    return i < 0 ? 0 : t->vals[i];
}

/* src/lib.rs:675-681  counts.insert(hash, count) */
void oxo_table_set_hash(oxo_table *t, uint64_t h, uint64_t v) {
    int f;
    *find_or_insert(t, h, &f) = v;
}

int oxo_table_contains(const oxo_table *t, uint64_t h) { return find(t, h) >= 0; }

/* src/lib.rs:213-224 */
int oxo_table_drop_hash(oxo_table *t, uint64_t h) {
    int64_t i = find(t, h);
    if (i < 0) return 0;
    t->state[i] = 2; t->live--;
    return 1;
}

/* src/lib.rs:227-246: remove count < min_count */
uint64_t oxo_table_mincut(oxo_table *t, uint64_t min_count) {
    uint64_t n = 0;
    for (uint64_t i = 0; i < t->cap; ++i)
        if (t->state[i] == 1 && t->vals[i] < min_count) { t->state[i] = 2; t->live--; n++; }
    return n;
}

/* src/lib.rs:249-267: remove count > max_count */
uint64_t oxo_table_maxcut(oxo_table *t, uint64_t max_count) {
    uint64_t n = 0;
    for (uint64_t i = 0; i < t->cap; ++i)
        if (t->state[i] == 1 && t->vals[i] > max_count) { t->state[i] = 2; t->live--; n++; }
    return n;
}

/* src/lib.rs:493-514: 0 for an empty table */
uint64_t oxo_table_min(const oxo_table *t) {
    uint64_t m = UINT64_MAX;
    if (!t->live) return 0;
    for (uint64_t i = 0; i < t->cap; ++i)
        if (t->state[i] == 1 && t->vals[i] < m) m = t->vals[i];
    return m;
}

uint64_t oxo_table_max(const oxo_table *t) {
    uint64_t m = 0;
    for (uint64_t i = 0; i < t->cap; ++i)
        if (t->state[i] == 1 && t->vals[i] > m) m = t->vals[i];
    return m;
}

/* src/lib.rs:537-539 */
uint64_t oxo_table_sum(const oxo_table *t) {
    uint64_t s = 0;
    for (uint64_t i = 0; i < t->cap; ++i)
        if (t->state[i] == 1) s += t->vals[i];
    return s;
}

/* ------------------------------------------------------------------ consume */

static int consume_one(oxo_table *t, scratch_t *s, const uint8_t *seq, size_t len, uint32_t ksize,
                       int skip_bad, uint64_t *n_out) {
    uint64_t n = 0; /* src/lib.rs:550 */
    if (ksize != 0 && len >= ksize) {
        if (scratch_reserve(s, len)) return OXO_ERR_NOMEM;
        prepare(s, seq, len);
        int64_t last_bad = -1;
        for (uint32_t j = 0; j + 1 < ksize; ++j)
            if (!is_acgt(s->up[j])) last_bad = j;
        for (size_t i = 0; i + ksize <= len; ++i) { /* src/lib.rs:586 */
            if (!is_acgt(s->up[i + ksize - 1])) last_bad = (int64_t)(i + ksize - 1);
            if (last_bad >= (int64_t)i) {
                if (skip_bad) continue; /* force=true yields Ok(0): src/lib.rs:589 */
                *n_out = n;             /* Err(_): src/lib.rs:593-596 */
                return OXO_ERR_BAD_KMER;
            }
            uint64_t h = window_hash(s->up + i, s->rc + (len - ksize - i), ksize);
            if (h == 0) continue;       /* Ok(0) => continue: src/lib.rs:589 */
            oxo_table_count_hash(t, h); /* src/lib.rs:590-592 */
            n++;                        /* src/lib.rs:599 */
        }
    }
    *n_out = n;
    return OXO_OK;
}

int oxo_consume(oxo_table *t, const uint8_t *seq, size_t len, uint32_t ksize, int skip_bad,
                uint64_t *n_out) {
    scratch_t s = {0, 0, 0};
    int rc = consume_one(t, &s, seq, len, ksize, skip_bad, n_out);
    free(s.up); free(s.rc);
    return rc;
}

void oxo_table_merge(oxo_table *dst, const oxo_table *src, uint64_t *counts_added,
                     uint64_t *new_keys) {
    /* src/lib.rs:798-806 */
    uint64_t added = 0, fresh_keys = 0;
    for (uint64_t i = 0; i < src->cap; ++i) {
        if (src->state[i] != 1) continue;
        int f;
        uint64_t *v = find_or_insert(dst, src->keys[i], &f);
        if (*v == 0) fresh_keys++; /* reference tests `*current_count == 0` */
        *v += src->vals[i];
        added += src->vals[i];
    }
    if (counts_added) *counts_added = added;
    if (new_keys) *new_keys = fresh_keys;
}

typedef struct {
    oxo_table *table;
    const uint8_t *bases;
    const uint64_t *offsets;
    uint64_t lo, hi, total;
    uint32_t ksize;
} shard_job;

static void *shard_main(void *arg) {
    shard_job *j = (shard_job *)arg;
    scratch_t s = {0, 0, 0};
    for (uint64_t r = j->lo; r < j->hi; ++r) {
        uint64_t n = 0;
        consume_one(j->table, &s, j->bases + j->offsets[r], j->offsets[r + 1] - j->offsets[r],
                    j->ksize, 1, &n);
        j->total += n;
    }
    free(s.up); free(s.rc);
    return NULL;
}

int oxo_consume_batch(oxo_table *t, const uint8_t *bases, const uint64_t *offsets,
                      uint64_t n_reads, uint32_t ksize, int skip_bad, int nthreads,
                      uint64_t *total_out, int64_t *err_read, uint64_t *err_pos) {
    uint64_t total = 0;
    if (err_read) *err_read = -1;
    if (err_pos) *err_pos = 0;
    if (nthreads > 1 && skip_bad) {
        shard_job *jobs = (shard_job *)calloc((size_t)nthreads, sizeof *jobs);
        pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof *th);
        for (int i = 0; i < nthreads; ++i) {
            jobs[i].table = i == 0 ? t : oxo_table_new();
            jobs[i].bases = bases; jobs[i].offsets = offsets; jobs[i].ksize = ksize;
            jobs[i].lo = n_reads * (uint64_t)i / (uint64_t)nthreads;
            jobs[i].hi = n_reads * (uint64_t)(i + 1) / (uint64_t)nthreads;
            if (i > 0) pthread_create(&th[i], NULL, shard_main, &jobs[i]);
        }
        shard_main(&jobs[0]);
        for (int i = 1; i < nthreads; ++i) pthread_join(th[i], NULL);
        for (int i = 0; i < nthreads; ++i) {
            total += jobs[i].total;
            if (i > 0) { oxo_table_merge(t, jobs[i].table, NULL, NULL); oxo_table_free(jobs[i].table); }
        }
        free(jobs); free(th);
        *total_out = total;
        return OXO_OK;
    }
    scratch_t s = {0, 0, 0};
    int rc = OXO_OK;
    for (uint64_t r = 0; r < n_reads; ++r) {
        uint64_t n = 0;
        rc = consume_one(t, &s, bases + offsets[r], offsets[r + 1] - offsets[r], ksize, skip_bad, &n);
        total += n;
        if (rc != OXO_OK) {
            if (err_read) *err_read = (int64_t)r;
            if (err_pos) *err_pos = n;
            break;
        }
    }
    free(s.up); free(s.rc);
    *total_out = total;
    return rc;
}

/* ------------------------------------------------------------------- export */

typedef struct { uint64_t k, v; } pair_t;
static int cmp_pair(const void *a, const void *b) {
    uint64_t x = ((const pair_t *)a)->k, y = ((const pair_t *)b)->k;
    return x < y ? -1 : x > y;
}

uint64_t oxo_table_export_sorted(const oxo_table *t, uint64_t *keys, uint64_t *vals, uint64_t cap) {
    if (cap == 0) return t->live;
    pair_t *p = (pair_t *)malloc((t->live + 1) * sizeof *p);
    uint64_t n = 0;
    for (uint64_t i = 0; i < t->cap; ++i)
        if (t->state[i] == 1) { p[n].k = t->keys[i]; p[n].v = t->vals[i]; n++; }
    qsort(p, n, sizeof *p, cmp_pair);
    uint64_t w = n < cap ? n : cap;
    for (uint64_t i = 0; i < w; ++i) { keys[i] = p[i].k; vals[i] = p[i].v; }
    free(p);
    return w;
}

/* src/lib.rs:465-488, zero=false branch: tally of counts, sorted by frequency */
uint64_t oxo_table_histo_sparse(const oxo_table *t, uint64_t *freq, uint64_t *n, uint64_t cap) {
    oxo_table *f = oxo_table_new();
    for (uint64_t i = 0; i < t->cap; ++i)
        if (t->state[i] == 1) oxo_table_count_hash(f, t->vals[i]);
    uint64_t m = oxo_table_export_sorted(f, freq, n, cap);
    oxo_table_free(f);
    return m;
}

/* src/lib.rs:610-624: key sets only; counts (even 0) are irrelevant */
void oxo_setop_sizes(const oxo_table *a, const oxo_table *b, uint64_t *inter, uint64_t *uni) {
    uint64_t both = 0;
    for (uint64_t i = 0; i < a->cap; ++i)
        if (a->state[i] == 1 && find(b, a->keys[i]) >= 0) both++;
    if (inter) *inter = both;
    if (uni) *uni = a->live + b->live - both;
}

/* src/lib.rs:708-722 */
double oxo_jaccard(const oxo_table *a, const oxo_table *b) {
    uint64_t inter, uni;
    oxo_setop_sizes(a, b, &inter, &uni);
    if (uni == 0) return 1.0;
    return (double)inter / (double)uni;
}

void oxo_table_digest(const oxo_table *t, uint64_t *n, uint64_t *sum, uint64_t *xor_keys,
                      uint64_t *sum_hc) {
    uint64_t s = 0, x = 0, hc = 0;
    for (uint64_t i = 0; i < t->cap; ++i)
        if (t->state[i] == 1) { s += t->vals[i]; x ^= t->keys[i]; hc += t->keys[i] * t->vals[i]; }
    *n = t->live; *sum = s; *xor_keys = x; *sum_hc = hc;
}

/*
 * oxli_oracle.h -- CPU restatement of the oxli k-mer counting hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or as
 * the timed CPU baseline.  The product path (oxli_b200/) never links, imports
 * or falls back to it.
 *
 * Parity status: PINNED.  The reference (Rust, /root/reference/src/lib.rs)
 * cannot be compiled in this image (no cargo/rustc; its arithmetic lives in
 * the un-vendored crates sourmash 0.23.0 -> murmurhash3 0.0.5, Cargo.lock
 * 1200-1203 / 628-631).  This restatement follows the reference's call sites
 * and the public MurmurHash3_x64_128 definition, and is pinned by
 *   - the 18 (k-mer, hash) known-answer vectors in the reference's own tests
 *     (tests/golden/ref_kats.json lists them with file:line),
 *   - the SMHasher verification constant 0x6384BA69 for MurmurHash3_x64_128,
 *   - every consume()/histo()/jaccard() behavioural assertion of the reference
 *     test-suite replayed in tests/test_oracle_*.py,
 *   - the published k-mer totals for doc/example.fa (README.md:94-99,
 *     doc/api.md:16-25).
 */
#ifndef OXLI_ORACLE_H
#define OXLI_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* status codes */
#define OXO_OK 0
#define OXO_ERR_WRONG_KSIZE 1 /* hash_kmer: len != ksize   (src/lib.rs:66-67) */
#define OXO_ERR_BAD_KMER 2    /* non-ACGT window in error mode (src/lib.rs:593-596) */
#define OXO_ERR_NOMEM 3

typedef struct oxo_table oxo_table;

/* MurmurHash3_x64_128 (public definition; crate murmurhash3 0.0.5). */
void oxo_murmur3_x64_128(const uint8_t *data, size_t len, uint32_t seed, uint64_t out[2]);
/* SMHasher VerificationTest for x64_128; returns 0x6384BA69 when correct. */
uint32_t oxo_murmur3_smhasher_verification(void);

/* Canonical k-mer hash: murmur3_x64_128(min(fw, revcomp(fw)) as upper-case
 * ASCII, seed 42).h1  -- sourmash SeqToHashes / _hash_murmur as driven from
 * src/lib.rs:65-81.  Returns OXO_ERR_WRONG_KSIZE or OXO_ERR_BAD_KMER. */
int oxo_hash_kmer(const uint8_t *kmer, size_t len, uint32_t ksize, uint64_t *out);

/* hash every window of one sequence; out[i] = hash, 0 for a bad window.
 * (sourmash SeqToHashes with force=true, as used at src/lib.rs:873-881).
 * n_out = max(L-k+1, 0) entries. */
void oxo_hash_windows(const uint8_t *seq, size_t len, uint32_t ksize, uint64_t *out);

/* table (stands in for HashMap<u64,u64>, src/lib.rs:33) */
oxo_table *oxo_table_new(void);
void oxo_table_free(oxo_table *t);
void oxo_table_clear(oxo_table *t);
uint64_t oxo_table_len(const oxo_table *t);                 /* src/lib.rs:665-667 */
uint64_t oxo_table_count_hash(oxo_table *t, uint64_t h);    /* src/lib.rs:100-104 */
void oxo_table_add_hash(oxo_table *t, uint64_t h, uint64_t c); /* += c */
uint64_t oxo_table_get_hash(const oxo_table *t, uint64_t h);/* src/lib.rs:185-188 */
void oxo_table_set_hash(oxo_table *t, uint64_t h, uint64_t v); /* src/lib.rs:675-681 */
int oxo_table_contains(const oxo_table *t, uint64_t h);
int oxo_table_drop_hash(oxo_table *t, uint64_t h);          /* src/lib.rs:213-224 */
uint64_t oxo_table_mincut(oxo_table *t, uint64_t min_count);/* src/lib.rs:227-246 */
uint64_t oxo_table_maxcut(oxo_table *t, uint64_t max_count);/* src/lib.rs:249-267 */
uint64_t oxo_table_min(const oxo_table *t);                 /* src/lib.rs:493-501 */
uint64_t oxo_table_max(const oxo_table *t);                 /* src/lib.rs:506-514 */
uint64_t oxo_table_sum(const oxo_table *t);                 /* src/lib.rs:537-539 */

/* consume: src/lib.rs:545-607 (store_kmers=false branch).
 *   skip_bad != 0: bad windows skipped, not counted.
 *   skip_bad == 0: windows before the first bad window are counted and stay
 *                  counted, then OXO_ERR_BAD_KMER with *n_out = number counted
 *                  (the {n} of "bad k-mer encountered at position {n}").
 * A valid window whose hash is 0 is skipped and not counted (src/lib.rs:589).
 * The caller does the `consumed` bookkeeping (src/lib.rs:604). */
int oxo_consume(oxo_table *t, const uint8_t *seq, size_t len, uint32_t ksize,
                int skip_bad, uint64_t *n_out);

/* batch of reads, CSR layout; same per-read semantics, reads in order; stops
 * at the first erroring read in error mode (*err_read = its index, *err_pos =
 * its n; -1 when none).  nthreads > 1 shards reads over threads with private
 * tables that are merged at the end (skip mode only). */
int oxo_consume_batch(oxo_table *t, const uint8_t *bases, const uint64_t *offsets,
                      uint64_t n_reads, uint32_t ksize, int skip_bad, int nthreads,
                      uint64_t *total_out, int64_t *err_read, uint64_t *err_pos);

/* export all (hash,count) pairs sorted by hash; returns number written
 * (<= cap).  Pass cap=0 to query the size. */
uint64_t oxo_table_export_sorted(const oxo_table *t, uint64_t *keys, uint64_t *vals, uint64_t cap);

/* histo(zero=False): sorted (freq, n) pairs (src/lib.rs:465-488). Returns
 * number of pairs (<= cap entries written). */
uint64_t oxo_table_histo_sparse(const oxo_table *t, uint64_t *freq, uint64_t *n, uint64_t cap);

/* |A n B| and |A u B| on key sets (src/lib.rs:610-624), jaccard (708-722). */
void oxo_setop_sizes(const oxo_table *a, const oxo_table *b, uint64_t *inter, uint64_t *uni);
double oxo_jaccard(const oxo_table *a, const oxo_table *b);

/* merge: src/lib.rs:778-837; returns counts added and new keys. */
void oxo_table_merge(oxo_table *dst, const oxo_table *src, uint64_t *counts_added, uint64_t *new_keys);

/* order-independent digests used by at-scale parity checks */
void oxo_table_digest(const oxo_table *t, uint64_t *n, uint64_t *sum, uint64_t *xor_keys,
                      uint64_t *sum_hc);

#ifdef __cplusplus
}
#endif
#endif

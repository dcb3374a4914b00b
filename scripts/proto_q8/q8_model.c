/* q8_model.c -- sequential CPU model of a compact count-table layout for the next round.
 * NOT product code and not used by it: a self-checking sketch of the slot encoding only.
 *
 * Why: ncu shows no L2 residency for the 128 MiB, 16-byte-slot table of the C2 workload, and
 * profiles/r1_microbench_l2_residency_sweep.txt puts random sector-load + RED updates at 67 G keys/s
 * on a 64 MiB table against 45 at 128 MiB.  An 8-byte slot halves the table.
 *
 * Layout (cap = 2^c slots, 4 slots = one 32-byte sector = one bucket, no probing past the bucket):
 *     key' = key * PHI (odd multiplier: a bijection on u64, undone with PHI^-1)
 *     bucket(key) = key' >> (66 - c)            top c-2 bits
 *     r(key)      = key' & (2^(66-c) - 1)       the rest: with the bucket index it IS the key
 *     slot word   = [ count : C = c-2 bits | r : 66-c bits ],  0 = empty (a stored count is >= 1)
 * Every increment is a blind add (RED on the device):
 *     - slot found, top count bit clear -> add 1 to the count field;
 *     - slot found, top count bit set   -> the key is "promoted": add to the exact spill table instead;
 *     - no slot, bucket has an empty one -> claim it (CAS on the device) with count 1;
 *     - no slot, bucket full            -> the key lives in the spill table only.
 *   count(key) = count field (if the key has a slot) + spill[key].
 * Why blind adds stay exact: after the top bit of a count field is set, only threads that loaded the
 * slot BEFORE that moment still add to it.  Their number is bounded by the keys in flight in one launch
 * (CTAs x threads x keys in flight per thread = 444 x 256 x 4 = 455 k today), so the field cannot
 * travel the remaining 2^(C-1) to a carry as long as 2^(C-1) > that bound: C >= 20, c >= 22, i.e. a
 * table of at least 32 MiB -- which is below the L2 knee anyway.  The count sits in the TOP bits so that
 * even a carry would fall off the word instead of into the key.
 *
 * Build and run:  gcc -O2 -o /tmp/q8_model scripts/proto_q8/q8_model.c && /tmp/q8_model
 */
#include <inttypes.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PHI 0x9E3779B97F4A7C15ULL
static uint64_t phi_inverse(void) {  /* Newton iteration for the inverse of an odd number mod 2^64 */
    uint64_t x = PHI;
    for (int i = 0; i < 6; ++i) x *= 2 - PHI * x;
    return x;
}

/* ---- exact spill table: the current 16-byte-slot layout, sequential ------------------------- */
typedef struct { uint64_t *keys, *vals; uint64_t cap, size; } spill_t;
static void spill_init(spill_t *s, uint64_t cap) {
    s->cap = cap; s->size = 0;
    s->keys = malloc(cap * 8); s->vals = calloc(cap, 8);
    memset(s->keys, 0xff, cap * 8);
}
static uint64_t *spill_slot(spill_t *s, uint64_t key, int create) {
    uint64_t i = (key * PHI) >> 40 & (s->cap - 1);
    for (;;) {
        if (s->keys[i] == key) return &s->vals[i];
        if (s->keys[i] == ~0ULL) {
            if (!create) return NULL;
            s->keys[i] = key; s->size++;
            return &s->vals[i];
        }
        i = (i + 1) & (s->cap - 1);
    }
}

/* ---- the compact table ----------------------------------------------------------------------- */
typedef struct {
    uint64_t *slots; int c; int rbits, cbits; uint64_t rmask, one, top; uint64_t keys_in_slots, spilled_keys, promoted;
    spill_t spill;
} q8_t;
static void q8_init(q8_t *t, int c) {
    t->c = c; t->rbits = 66 - c; t->cbits = c - 2;
    t->rmask = (1ULL << t->rbits) - 1;
    t->one = 1ULL << t->rbits;               /* +1 in the count field */
    t->top = 1ULL << 63;                     /* top bit of the count field */
    t->slots = calloc(1ULL << c, 8);
    t->keys_in_slots = t->spilled_keys = t->promoted = 0;
    spill_init(&t->spill, 1ULL << c);  /* roomy: in this small model most keys end up promoted */
}
static uint64_t q8_bucket(const q8_t *t, uint64_t key) { return ((key * PHI) >> t->rbits) * 4; }
static uint64_t q8_rem(const q8_t *t, uint64_t key) { return (key * PHI) & t->rmask; }
static uint64_t q8_key_of(const q8_t *t, uint64_t slot_index, uint64_t word, uint64_t phi_inv) {
    return (((slot_index / 4) << t->rbits) | (word & t->rmask)) * phi_inv;
}
static void q8_add(q8_t *t, uint64_t key, uint64_t inc) {  /* inc = 1 on the hot path */
    uint64_t *b = t->slots + q8_bucket(t, key);
    const uint64_t r = q8_rem(t, key);
    for (int q = 0; q < 4; ++q) {
        if (b[q] != 0 && (b[q] & t->rmask) == r) {
            if (b[q] & t->top) { *spill_slot(&t->spill, key, 1) += inc; return; }   /* promoted */
            /* a bulk add (merge, load) must not jump over the top bit's half: split it */
            const uint64_t room = ((t->top - (b[q] & ~t->rmask)) >> t->rbits);      /* adds until the top bit sets */
            if (inc <= room) { b[q] += inc * t->one; if (b[q] & t->top) t->promoted++; return; }
            b[q] += room * t->one; t->promoted++;
            *spill_slot(&t->spill, key, 1) += inc - room;
            return;
        }
    }
    for (int q = 0; q < 4; ++q) {
        if (b[q] == 0) {
            b[q] = r;  /* claim; then add like anyone else (count 0 + r == 0 would read as empty only if r == 0 too) */
            t->keys_in_slots++;
            const uint64_t room = 1ULL << (t->cbits - 1);
            const uint64_t now = inc <= room ? inc : room;
            b[q] += now * t->one;
            if (b[q] & t->top) t->promoted++;
            if (inc > now) *spill_slot(&t->spill, key, 1) += inc - now;
            return;
        }
    }
    uint64_t *v = spill_slot(&t->spill, key, 1);   /* bucket full of other keys */
    if (*v == 0) t->spilled_keys++;
    *v += inc;
}
static uint64_t q8_get(q8_t *t, uint64_t key) {
    const uint64_t *b = t->slots + q8_bucket(t, key);
    const uint64_t r = q8_rem(t, key);
    uint64_t total = 0; int found = 0, full = 1;
    for (int q = 0; q < 4; ++q) {
        if (b[q] == 0) { full = 0; continue; }
        if ((b[q] & t->rmask) == r) { total = b[q] >> t->rbits; found = 1; if (!(b[q] & t->top)) return total; }
    }
    if (found || full) { uint64_t *v = spill_slot(&t->spill, key, 0); if (v) total += *v; }
    return total;
}

/* ---- check against a plain exact table ------------------------------------------------------- */
static uint64_t rng_state = 0x1234567;
static uint64_t rnd(void) { uint64_t x = (rng_state += 0x9E3779B97F4A7C15ULL); x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31); }

int main(void) {
    const int c = 16;                      /* small table so that promotion (2^13 counts) and full buckets happen */
    q8_t t; q8_init(&t, c);
    spill_t truth; spill_init(&truth, 1ULL << 18);
    const uint64_t phi_inv = phi_inverse();
    if (PHI * phi_inv != 1) { puts("bad inverse"); return 1; }
    const uint64_t n_keys = 40000;         /* load 0.61 of 65536 slots */
    uint64_t *keys = malloc(n_keys * 8);
    for (uint64_t i = 0; i < n_keys; ++i) keys[i] = rnd();
    keys[0] = 0; keys[1] = ~0ULL - 1;      /* edge keys (2^64-1 itself is the spill table's empty marker, as today) */
    uint64_t total = 0;
    for (uint64_t i = 0; i < 3000000; ++i) {
        const uint64_t x = rnd();
        /* a few very hot keys (poly-A like), the rest uniform */
        const uint64_t key = (x & 7) == 0 ? keys[x >> 3 & 3] : keys[(x >> 3) % n_keys];
        const uint64_t inc = (x >> 52) == 0 ? 1 + (x >> 36 & 0xffff) : 1;   /* sometimes a bulk add, as merge/load do */
        q8_add(&t, key, inc); *spill_slot(&truth, key, 1) += inc; total += inc;
    }
    uint64_t bad = 0, sum = 0;
    for (uint64_t i = 0; i < n_keys; ++i) {
        const uint64_t want = *spill_slot(&truth, keys[i], 1), got = q8_get(&t, keys[i]);
        if (want != got) { if (bad++ < 5) printf("key %016" PRIx64 ": want %" PRIu64 " got %" PRIu64 "\n", keys[i], want, got); }
    }
    for (int i = 0; i < 1000; ++i) if (q8_get(&t, rnd() | 1ULL << 63) != 0 && ++bad < 5) puts("absent key has a count");
    /* export: every slot word decodes back to its key */
    for (uint64_t s = 0; s < (1ULL << c); ++s) {
        if (!t.slots[s]) continue;
        const uint64_t key = q8_key_of(&t, s, t.slots[s], phi_inv);
        uint64_t *v = spill_slot(&truth, key, 0);
        if (!v) { if (bad++ < 5) printf("slot %" PRIu64 " decodes to an unknown key\n", s); continue; }
        sum += t.slots[s] >> t.rbits;
    }
    for (uint64_t i = 0; i < t.spill.cap; ++i) if (t.spill.keys[i] != ~0ULL) sum += t.spill.vals[i];
    if (sum != total) { printf("sum of counts %" PRIu64 " != %" PRIu64 "\n", sum, total); ++bad; }
    printf("c=%d: %" PRIu64 " keys in slots, %" PRIu64 " keys only in the spill table (%.1f %%), %" PRIu64 " promoted, %s\n", c,
           t.keys_in_slots, t.spilled_keys, 100.0 * t.spilled_keys / n_keys, t.promoted, bad ? "MISMATCH" : "all counts exact");
    return bad != 0;
}

"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on the same inputs.  Integer work, so every comparison is bit-exact."""
import hashlib

import numpy as np
import pytest

import oracle
from oracle import OracleTable
from synth import ragged_batch, synth_reads, uniform_offsets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from oxli_b200 import _capi

    assert _capi.lib.oxg_device_count() > 0, "GPU tests need a CUDA device"
    return _capi


def sha_pairs(k: np.ndarray, v: np.ndarray) -> str:
    inter = np.empty(2 * len(k), dtype="<u8")
    inter[0::2], inter[1::2] = k, v
    return hashlib.sha256(inter.tobytes()).hexdigest()


def assert_same_table(gpu, ora: OracleTable):
    gk, gv = gpu.export(1)
    ok, ov = ora.items_sorted()
    assert len(gk) == len(ok), (len(gk), len(ok))
    assert np.array_equal(gk, ok)
    assert np.array_equal(gv, ov)
    assert len(gpu) == len(ora)


# ---------------------------------------------------------------- K1: hashes

@pytest.mark.parametrize("k", [21, 31])
def test_hash_windows_example_fa(capi, example_seq, k):
    seq = example_seq.encode()
    got = capi.Table(k).hash_windows(seq)
    want = oracle.hash_windows(seq, k)
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 20, 21, 22, 24, 31, 32, 33, 40, 41, 51, 63, 64, 100, 255])
def test_hash_windows_every_k_with_junk(capi, k):
    rng = np.random.default_rng(k)
    bases, _ = ragged_batch(rng, 1, 9000, p_bad=0.004, p_lower=0.2, p_empty=0.0)
    if len(bases) < k + 10:
        bases = np.frombuffer(b"ACGTTGCAAC" * 100, dtype=np.uint8)
    got = capi.Table(k).hash_windows(bases)
    want = oracle.hash_windows(bases, k)
    assert np.array_equal(got, want), np.flatnonzero(got != want)[:10]


def test_hash_kats_through_gpu(capi, goldens):
    for e in goldens["kmer_hash"]:
        k = len(e["kmer"])
        assert int(capi.Table(k).hash_windows(e["kmer"].encode())[0]) == e["hash"], e
    # k=31 / k=21 / k=16 goldens derived in SURVEY.md 8c
    t = capi.Table(31)
    assert int(t.hash_windows(b"CGGAGGAAGCAAGAACAAAATATTTTTTCAT")[0]) == 13905159898738093091
    assert int(t.hash_windows(b"ATGAAAAAATATTTTGTTCTTGCTTCCTCCG")[0]) == 13905159898738093091
    assert int(capi.Table(21).hash_windows(b"AAATCTTATAAAATAACCACA")[0]) == 14908242140577922293
    assert int(capi.Table(16).hash_windows(b"ACGTACGTACGTACGT")[0]) == 4706917051267373191


def test_hash_palindromes_and_ties(capi):
    # fw == rc for the first 8+ bases forces the tie path of the canonical compare
    for k, seq in [(20, b"ACGTACGTACGTACGTACGT"), (31, b"AAAAAAAACCCCCCCGGGGGGGGTTTTTTTT"[:31]),
                   (31, b"ACGTACGTACGTACGAACGTACGTACGTACG"), (32, b"AAAAAAAATTTTTTTTAAAAAAAATTTTTTTT"),
                   (21, b"ATATATATATATATATATATA"), (31, b"A" * 31), (31, b"T" * 31)]:
        got = capi.Table(k).hash_windows(seq)
        assert np.array_equal(got, oracle.hash_windows(seq, k)), (k, seq)


# ------------------------------------------------------- K1+K2: consume parity

def test_example_fa_goldens(capi, example_seq, goldens):
    seq = np.frombuffer(example_seq.encode(), dtype=np.uint8)
    offs = np.array([0, len(seq)], dtype=np.uint64)
    for k in (21, 31):
        g = goldens["derived"][str(k)]
        t = capi.Table(k)
        st, n, er, _ = t.consume_batch(seq, offs)
        assert st == 0 and er == -1 and n == g["n_kmers"] == goldens["example_fa_kmers"][str(k)]
        kk, vv = t.export(1)
        assert len(kk) == g["distinct"] and sha_pairs(kk, vv) == g["sha256"]
        d = t.digest()
        assert (d["sum"], d["xor"], d["sum_hc"]) == (g["sum"], g["xor"], g["sum_hc"])
        assert [list(x) for x in t.histo()] == g["histo"]
        s = t.stats()
        assert (s["len"], s["sum"], s["min"], s["max"]) == (g["distinct"], g["sum"], g["min"], g["max"])


@pytest.mark.parametrize("k", [4, 15, 16, 19, 21, 24, 25, 29, 31, 32, 33, 41, 51, 63, 64])
def test_consume_ragged_batch_skip_mode(capi, k):
    rng = np.random.default_rng(100 + k)
    bases, offs = ragged_batch(rng, 4000, 260, p_bad=0.01)
    ora = OracleTable(k)
    want_total, _, _ = ora.consume_batch(bases, offs, True, nthreads=4)
    t = capi.Table(k)
    st, total, er, ep = t.consume_batch(bases, offs, True)
    assert (st, total, er) == (0, want_total, -1)
    assert_same_table(t, ora)


@pytest.mark.parametrize("k", [21, 31, 40])
def test_consume_incremental_equals_batch(capi, k):
    # many small calls (the reference's per-record usage) == one batch
    rng = np.random.default_rng(5)
    bases, offs = ragged_batch(rng, 60, 400, p_bad=0.005)
    one, many, ora = capi.Table(k), capi.Table(k), OracleTable(k)
    one.consume_batch(bases, offs)
    tot = 0
    for r in range(len(offs) - 1):
        seg = bases[int(offs[r]):int(offs[r + 1])]
        st, n, _, _ = many.consume_batch(seg, np.array([0, len(seg)], dtype=np.uint64))
        assert st == 0 and n == ora.consume(seg.tobytes())
        tot += n
    assert_same_table(one, ora)
    assert_same_table(many, ora)


@pytest.mark.parametrize("k", [5, 17, 21, 27, 31, 35, 40, 51, 63])
def test_consume_error_mode(capi, k):
    rng = np.random.default_rng(k)
    # clean reads, then one read with a bad byte in the middle, then more reads
    clean, offs = ragged_batch(rng, 300, 220, p_bad=0.0, p_empty=0.1)
    bad_read = np.frombuffer(b"ACGT" * 30, dtype=np.uint8).copy()
    bad_read[77] = ord("N")
    tail, toffs = ragged_batch(rng, 50, 220, p_bad=0.02)
    bases = np.concatenate([clean, bad_read, tail])
    offsets = np.concatenate([offs, [offs[-1] + len(bad_read)], offs[-1] + len(bad_read) + toffs[1:]]).astype(np.uint64)
    ora = OracleTable(k)
    want = ora.consume_batch(bases, offsets, skip_bad_kmers=False)
    t = capi.Table(k)
    st, total, er, ep = t.consume_batch(bases, offsets, skip_bad=False)
    assert st == capi.ERR_BAD_KMER
    assert (total, er, ep) == want
    assert er == 300 and ep == 77 - k + 1
    assert_same_table(t, ora)
    assert b"bad k-mer encountered at position %d" % ep in capi.lib.oxg_last_error()
    # no bad window anywhere -> error mode == skip mode
    t2, o2 = capi.Table(k), OracleTable(k)
    st, total, er, _ = t2.consume_batch(clean, offs, skip_bad=False)
    assert (st, total, er) == (0, o2.consume_batch(clean, offs, False)[0], -1)
    assert_same_table(t2, o2)


def test_error_mode_reference_cases(capi):
    # src/python/tests/test_basic.py:75-88 through the batch ABI
    for seq, pos in ((b"ATCGGX", 2), (b"XATCGG", 0)):
        t = capi.Table(4)
        a = np.frombuffer(seq, dtype=np.uint8)
        st, total, er, ep = t.consume_batch(a, np.array([0, len(a)], dtype=np.uint64), skip_bad=False)
        assert (st, er, ep, total) == (capi.ERR_BAD_KMER, 0, pos, pos)
        assert len(t) == pos
    # bad byte in a read shorter than k raises nothing (no window exists)
    t = capi.Table(5)
    a = np.frombuffer(b"ACGTACNN", dtype=np.uint8)
    st, total, er, _ = t.consume_batch(a, np.array([0, 6, 8], dtype=np.uint64), skip_bad=False)
    assert (st, total, er) == (0, 2, -1)


def test_skip_mode_reference_cases(capi):
    # test_basic.py:91-108, test_kmers_and_hashes.py:255-283, test_add.py:112-125
    t = capi.Table(3)
    a = np.frombuffer(b"XAAAAAXGGGG", dtype=np.uint8)
    st, n, _, _ = t.consume_batch(a, np.array([0, len(a)], dtype=np.uint64))
    assert n == 5 and len(t) == 2
    assert list(t.get_hashes([10679328328772601858, 12126843654075378313, 5])) == [3, 2, 0]
    t = capi.Table(5)
    a = np.frombuffer(b"ATGC" * 100000, dtype=np.uint8)
    st, n, _, _ = t.consume_batch(a, np.array([0, len(a)], dtype=np.uint64))
    ora = OracleTable(5)
    assert n == 399996 == ora.consume(a.tobytes()) and t.stats()["sum"] == 399996
    assert_same_table(t, ora)
    t = capi.Table(4)  # short read: nothing counted
    a = np.frombuffer(b"ACG", dtype=np.uint8)
    assert t.consume_batch(a, np.array([0, 3], dtype=np.uint64))[1] == 0 and len(t) == 0


def test_growth_from_tiny_table(capi):
    # 1.5 M distinct keys into a table created at minimum capacity: exercises the
    # load-limit deferral list, device rehash and replay
    k = 25
    bases = synth_reads(12000, 150, 2_000_000, seed=11)
    offs = uniform_offsets(12000, 150)
    ora = OracleTable(k)
    want, _, _ = ora.consume_batch(bases, offs, True, nthreads=8)
    t = capi.Table(k)
    assert t.capacity == 1024
    st, total, _, _ = t.consume_batch(bases, offs)
    assert total == want
    assert t.capacity >= 2 * len(ora)
    assert_same_table(t, ora)
    # and again on top (all keys present now): counts double
    t.consume_batch(bases, offs)
    ora.consume_batch(bases, offs, True, nthreads=8)
    assert_same_table(t, ora)


def test_high_coverage_input_does_not_inflate_the_table(capi):
    """Several launches over reads that keep re-covering one small genome: every key is created by
    the first launch, and the growth look-ahead must not enlarge a table that already fits them
    (a 4x larger table costs a third of the throughput on BASELINE.json configs[1])."""
    k, L, n, G = 31, 150, 1_400_000, 500_000   # 210 Mbases = 4 launches of 64 Mi windows
    d_bases = capi.device_alloc(n * L + 64)
    d_offs = capi.device_alloc((n + 1) * 8)
    try:
        capi.synth_reads_device(d_bases, n, L, G, seed=11)
        capi.h2d(d_offs, uniform_offsets(n, L))
        for hint in (G, 0):
            t = capi.Table(k, capacity_hint=hint)
            st, total, _, _ = t.consume_batch_device(d_bases, d_offs, n, n * L, True)
            assert st == 0 and total == n * (L - k + 1)
            assert G - k <= len(t) <= G
            assert t.capacity * 16 <= 128 << 20, t.capacity
            if hint:
                assert t.capacity == capi.Table(k, capacity_hint=hint).capacity
    finally:
        capi.device_free(d_bases)
        capi.device_free(d_offs)


def test_multi_chunk_streaming(capi):
    # > 2 host chunks (64 MiB each) from pageable memory; reads straddle chunk edges
    k, L, n, G = 31, 151, 900_000, 300_000
    d_bases = capi.device_alloc(n * L + 64)
    try:
        capi.synth_reads_device(d_bases, n, L, G, seed=3, sub_ppm=2000, n_ppm=500)
        bases = np.empty(n * L, dtype=np.uint8)
        capi.d2h(bases, d_bases)
    finally:
        capi.device_free(d_bases)
    assert np.array_equal(bases[: 1000 * L], synth_reads(1000, L, G, seed=3, sub_ppm=2000, n_ppm=500))
    offs = uniform_offsets(n, L)
    assert bases.nbytes > 2 * (64 << 20)
    ora = OracleTable(k)
    want, _, _ = ora.consume_batch(bases, offs, True, nthreads=8)
    t = capi.Table(k)
    st, total, _, _ = t.consume_batch(bases, offs)
    assert total == want
    assert_same_table(t, ora)
    # same bytes from pinned memory (no staging copy) on top: every count doubles
    pinned = capi.pinned_empty(bases.nbytes)
    pinned[:] = bases
    st, total2, _, _ = t.consume_batch(pinned, offs)
    capi.pinned_free(pinned)
    assert total2 == want
    gk, gv = t.export(1)
    ok, ov = ora.items_sorted()
    assert np.array_equal(gk, ok) and np.array_equal(gv, 2 * ov)


def test_error_mode_across_host_chunks(capi):
    """Error mode on a host batch of three staging chunks whose first bad k-mer lies in the second:
    the pre-scan stops there, and the counted prefix equals skip mode on exactly that prefix
    (src/lib.rs:586-600: everything before the failing window is counted, then the error)."""
    k, L, n, G = 31, 151, 900_000, 300_000
    d_bases = capi.device_alloc(n * L + 64)
    try:
        capi.synth_reads_device(d_bases, n, L, G, seed=5)
        bases = np.empty(n * L, dtype=np.uint8)
        capi.d2h(bases, d_bases)
    finally:
        capi.device_free(d_bases)
    offs = uniform_offsets(n, L)
    r_bad, at = 500_000, 60                      # byte 75.5 M: second 64 MiB chunk
    bases[r_bad * L + at] = ord("N")
    bases[(r_bad + 300_000) * L + 7] = ord("N")  # a later one, in the third chunk, must not matter
    for src in ("pageable", "pinned"):
        buf = bases
        if src == "pinned":
            buf = capi.pinned_empty(bases.nbytes)
            buf[:] = bases
        t = capi.Table(k)
        st, total, er, ep = t.consume_batch(buf, offs, skip_bad=False)
        assert (st, er, ep) == (capi.ERR_BAD_KMER, r_bad, at - k + 1)
        ref = capi.Table(k)
        cut = r_bad * L + at                     # the clean prefix of the failing read is a read of its own
        pre_offs = np.concatenate([offs[: r_bad + 1], np.array([cut], dtype=np.uint64)])
        st2, want, er2, _ = ref.consume_batch(buf[:cut], pre_offs, skip_bad=True)
        assert (st2, er2) == (0, -1) and total == want == r_bad * (L - k + 1) + (at - k + 1)
        a, b = t.export(1), ref.export(1)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        if src == "pinned":
            capi.pinned_free(buf)


# ------------------------------------------------------------- device-resident

def test_device_resident_consume_and_generator(capi):
    k, L, n, G = 31, 150, 200_000, 500_000
    want_bases = synth_reads(n, L, G, seed=0xC20001, sub_ppm=10000, n_ppm=1000)
    d_bases = capi.device_alloc(n * L + 64)
    d_offs = capi.device_alloc((n + 1) * 8)
    try:
        capi.synth_reads_device(d_bases, n, L, G, 0xC20001, sub_ppm=10000, n_ppm=1000)
        got = np.empty(n * L, dtype=np.uint8)
        capi.d2h(got, d_bases)
        assert np.array_equal(got, want_bases)
        offs = uniform_offsets(n, L)
        capi.h2d(d_offs, offs)
        for skip in (True, False):
            ora = OracleTable(k)
            want = ora.consume_batch(want_bases, offs, skip, nthreads=8 if skip else 1)
            t = capi.Table(k, capacity_hint=2 * G)
            st, total, er, ep = t.consume_batch_device(d_bases, d_offs, n, n * L, skip)
            assert (total, er, ep) == want
            assert st == (0 if er < 0 else capi.ERR_BAD_KMER)
            assert_same_table(t, ora)
    finally:
        capi.device_free(d_bases)
        capi.device_free(d_offs)


def test_c3_shaped_k21_errors_and_ns(capi):
    # BASELINE.json configs[2] in miniature: k=21, 1 % substitutions, 0.1 % N, both bad-k-mer modes
    k, L, n, G = 21, 150, 300_000, 2_000_000
    bases = synth_reads(n, L, G, seed=0xC30001, sub_ppm=10000, n_ppm=1000)
    offs = uniform_offsets(n, L)
    for skip in (True, False):
        ora = OracleTable(k)
        want = ora.consume_batch(bases, offs, skip, nthreads=8 if skip else 1)
        t = capi.Table(k)
        st, total, er, ep = t.consume_batch(bases, offs, skip)
        assert (total, er, ep) == want and st == (0 if er < 0 else capi.ERR_BAD_KMER)
        assert_same_table(t, ora)
        if skip:
            assert t.histo() == ora.histo(zero=False)
            s = t.stats()
            assert (s["len"], s["sum"], s["min"], s["max"]) == (len(ora), ora.sum_counts, ora.min, ora.max)


def test_long_reads_k21(capi):
    # configs[4]: 10-kbp reads (intra-read tiling: one read spans ~40 warp tiles)
    k, L, n, G = 21, 10_000, 1500, 400_000
    bases = synth_reads(n, L, G, seed=0xC50001, sub_ppm=3000, n_ppm=200)
    offs = uniform_offsets(n, L)
    ora = OracleTable(k)
    want, _, _ = ora.consume_batch(bases, offs, True, nthreads=8)
    t = capi.Table(k)
    st, total, _, _ = t.consume_batch(bases, offs)
    assert total == want
    assert_same_table(t, ora)
    # jaccard against a table built from the first half of the reads (key sets nest)
    half, oh = capi.Table(k), OracleTable(k)
    half.consume_batch(bases[: n // 2 * L], offs[: n // 2 + 1])
    oh.consume_batch(bases[: n // 2 * L], offs[: n // 2 + 1], True, nthreads=8)
    assert t.setop_sizes(half) == ora.setop_sizes(oh) and t.jaccard(half) == ora.jaccard(oh)


# ------------------------------------------------------------------ table ops

def test_hash_level_ops(capi):
    rng = np.random.default_rng(9)
    t, ora = capi.Table(21), OracleTable(21)
    special = np.array([0, 1, 2**64 - 1, 2**64 - 2, 2**63, 42], dtype=np.uint64)
    keys = np.concatenate([special, rng.integers(0, 2**63, size=5000, dtype=np.uint64)])
    stream = rng.choice(keys, size=60000)
    for h in stream:
        ora.count_hash(int(h))
    t.count_hashes(stream[:30000])
    t.count_hashes(stream[30000:])
    assert_same_table(t, ora)
    probe = np.concatenate([keys, rng.integers(0, 2**63, size=100, dtype=np.uint64)])
    assert list(t.get_hashes(probe)) == [ora.get_hash(int(h)) for h in probe]
    # count_hash returns the new count (src/lib.rs:100-104)
    for h in (0, 2**64 - 1, 12345):
        assert int(t.count_hashes([h], want_counts=True)[0]) == ora.count_hash(h)
    # __setitem__ incl. value 0 and the out-of-band key
    for h, v in ((7, 0), (2**64 - 1, 0), (int(keys[10]), 99), (2**64 - 1, 5)):
        t.set_hash(h, v)
        ora.set_hash(h, v)
    assert_same_table(t, ora)
    s = t.stats()
    assert (s["len"], s["sum"], s["min"], s["max"]) == (len(ora), ora.sum_counts, ora.min, ora.max)
    assert t.histo() == ora.histo(zero=False)
    # drop
    victims = [int(x) for x in keys[:200]] + [2**64 - 1, 999999999]
    want_removed = sum(1 for h in set(victims) if ora.contains(h))
    assert t.erase_hashes(victims) == want_removed
    for h in victims:
        ora.drop_hash(h)
    assert_same_table(t, ora)
    # a list long enough for the mark-and-rebuild path: duplicates, missing keys, the out-of-band key
    t.set_hash(2**64 - 1, 3); ora.set_hash(2**64 - 1, 3)
    bulk = [int(x) for x in keys[200:2200]] + [int(x) for x in keys[300:400]] + [2**64 - 1, 5, 123456789]
    want_removed = sum(1 for h in set(bulk) if ora.contains(h))
    assert t.erase_hashes(bulk) == want_removed
    for h in bulk:
        ora.drop_hash(h)
    assert_same_table(t, ora)
    assert list(t.get_hashes(probe)) == [ora.get_hash(int(h)) for h in probe]
    # mincut / maxcut
    assert t.cut(0, 3) == ora.mincut(3)
    assert_same_table(t, ora)
    assert t.cut(1, 15) == ora.maxcut(15)
    assert_same_table(t, ora)
    # empty-table scalars (test_histo.py:12-68)
    e = capi.Table(4)
    assert e.stats() == {"len": 0, "sum": 0, "min": 0, "max": 0} and e.histo() == [] and len(e) == 0


def test_huge_counts_in_histo(capi):
    t, ora = capi.Table(21), OracleTable(21)
    for h, v in ((1, 70000), (2, 70000), (3, 2**40), (4, 65535), (5, 65536), (6, 1023), (7, 1024), (8, 0)):
        t.set_hash(h, v)
        ora.set_hash(h, v)
    assert t.histo() == ora.histo(zero=False)
    s = t.stats()
    assert (s["min"], s["max"], s["sum"]) == (0, 2**40, ora.sum_counts)


def test_export_order_is_stable(capi):
    rng = np.random.default_rng(2)
    t = capi.Table(21)
    t.count_hashes(rng.integers(0, 2**64 - 1, size=20000, dtype=np.uint64))
    k1, v1 = t.export(0)
    k2, v2 = t.export(0)
    assert np.array_equal(k1, k2) and np.array_equal(v1, v2)  # dump() == list(iter), test_dump.py:35-43
    ks, vs = t.export(1)
    assert np.array_equal(ks, np.sort(k1))
    kc, vc = t.export(2)
    order = np.lexsort((k1, v1))
    assert np.array_equal(kc, k1[order]) and np.array_equal(vc, v1[order])


def test_set_operations_and_similarity(capi):
    rng = np.random.default_rng(4)
    pool = rng.integers(0, 2**64 - 1, size=30000, dtype=np.uint64)
    ka = np.concatenate([pool[:20000], np.array([2**64 - 1], dtype=np.uint64)])
    kb = np.concatenate([pool[10000:], np.array([2**64 - 1, 0], dtype=np.uint64)])
    a, b, oa, ob = capi.Table(21), capi.Table(21), OracleTable(21), OracleTable(21)
    sa, sb = rng.choice(ka, 50000), rng.choice(kb, 50000)
    a.count_hashes(sa); b.count_hashes(sb)
    for h in sa:
        oa.count_hash(int(h))
    for h in sb:
        ob.count_hash(int(h))
    b.set_hash(77, 0); ob.set_hash(77, 0)  # zero-valued member
    assert a.setop_sizes(b) == oa.setop_sizes(ob)
    assert a.jaccard(b) == oa.jaccard(ob)  # one f64 divide of two exact integers
    assert set(map(int, a.setop(b, 0))) == oa.union(ob)
    assert set(map(int, a.setop(b, 1))) == oa.intersection(ob)
    assert set(map(int, a.setop(b, 2))) == oa.difference(ob)
    assert set(map(int, a.setop(b, 3))) == oa.symmetric_difference(ob)
    e1, e2 = capi.Table(21), capi.Table(21)
    assert e1.jaccard(e2) == 1.0 and e1.jaccard(a) == 0.0 and a.jaccard(a) == 1.0
    # cosine (reference tests use rel_tol 1e-5 against scipy, test_metrics.py)
    ak, av = oa.items_sorted(); bk, bv = ob.items_sorted()
    da, db = dict(zip(ak.tolist(), av.tolist())), dict(zip(bk.tolist(), bv.tolist()))
    dot = sum(v * db[k] for k, v in da.items() if k in db)
    want = dot / (np.sqrt(sum(v * v for v in da.values())) * np.sqrt(sum(v * v for v in db.values())))
    assert abs(a.cosine(b) - want) <= 1e-9 * want
    assert e1.cosine(a) == 0.0


def test_merge(capi):
    rng = np.random.default_rng(8)
    a, b, oa, ob = capi.Table(21), capi.Table(21), OracleTable(21), OracleTable(21)
    sa = rng.integers(0, 5000, size=20000, dtype=np.uint64)
    sb = rng.integers(2500, 9000, size=20000, dtype=np.uint64)
    a.count_hashes(sa); b.count_hashes(sb)
    for h in sa:
        oa.count_hash(int(h))
    for h in sb:
        ob.count_hash(int(h))
    a.set_hash(3000, 0); oa.set_hash(3000, 0)  # existing zero-valued key counts as "new" (src/lib.rs:801)
    b.set_hash(2**64 - 1, 4); ob.set_hash(2**64 - 1, 4)
    assert a.merge(b) == oa.add(ob)
    assert_same_table(a, oa)
    assert_same_table(b, ob)


def test_export_sorted_on_the_device(capi):
    # dump(sortkeys) / dump(sortcounts) (src/lib.rs:330-381): LSD radix sort on the device; many
    # tiles, counts above one byte, keys 0 and 2^64-1 (the latter lives outside the slot array)
    bases = synth_reads(120_000, 150, 60_000, seed=5, sub_ppm=3000)  # ~300x coverage + error singletons
    offs = uniform_offsets(120_000, 150)
    t = capi.Table(21)
    t.consume_batch(bases, offs)
    t.count_hashes(np.array([0, 0, 2**64 - 1, 2**64 - 1, 2**64 - 1, 7], dtype=np.uint64))
    k0, v0 = t.export(0)
    assert len(k0) == len(t) > 100_000 and v0.max() > 300
    k0b, v0b = t.export(0)
    assert np.array_equal(k0, k0b) and np.array_equal(v0, v0b)  # slot order is stable between calls
    k1, v1 = t.export(1)
    order = np.argsort(k0, kind="stable")
    assert np.array_equal(k1, k0[order]) and np.array_equal(v1, v0[order])
    assert k1[0] == 0 and v1[0] == 2 and k1[-1] == 2**64 - 1 and v1[-1] == 3
    k2, v2 = t.export(2)
    order = np.lexsort((k0, v0))  # by count, then by hash
    assert np.array_equal(k2, k0[order]) and np.array_equal(v2, v0[order])

"""Pins the CPU oracle (oracle/) against the reference's own golden vectors and
behavioural assertions.  Each test names the reference test it replays
(paths relative to the reference root, src/python/tests/)."""
import numpy as np
import pytest

import oracle
from oracle import OracleTable

from synth import ragged_batch


def test_kmer_hash_known_answers(goldens):
    # test_basic.py:132-143, test_kmers_and_hashes.py:12-46,101-107,261, test_dump.py:13-17, test_add.py:33
    for e in goldens["kmer_hash"]:
        assert oracle.hash_kmer(e["kmer"]) == e["hash"], e


def test_smhasher_verification(goldens):
    assert oracle.smhasher_verification() == int(goldens["smhasher_x64_128_verification"], 16)


def test_murmur_tail_paths_match_pure_python():
    # independent pure-Python MurmurHash3_x64_128 for every length 0..48 (body + both tail words)
    def rotl(x, r):
        return ((x << r) | (x >> (64 - r))) & (2**64 - 1)

    def fmix(k):
        k ^= k >> 33; k = k * 0xFF51AFD7ED558CCD % 2**64
        k ^= k >> 33; k = k * 0xC4CEB9FE1A85EC53 % 2**64
        return k ^ (k >> 33)

    def mm(data, seed):
        c1, c2, M = 0x87C37B91114253D5, 0x4CF5AD432745937F, 2**64
        h1 = h2 = seed
        nb = len(data) // 16
        for b in range(nb):
            k1 = int.from_bytes(data[16 * b:16 * b + 8], "little")
            k2 = int.from_bytes(data[16 * b + 8:16 * b + 16], "little")
            k1 = k1 * c1 % M; k1 = rotl(k1, 31); k1 = k1 * c2 % M; h1 ^= k1
            h1 = rotl(h1, 27); h1 = (h1 + h2) % M; h1 = (h1 * 5 + 0x52DCE729) % M
            k2 = k2 * c2 % M; k2 = rotl(k2, 33); k2 = k2 * c1 % M; h2 ^= k2
            h2 = rotl(h2, 31); h2 = (h2 + h1) % M; h2 = (h2 * 5 + 0x38495AB5) % M
        tail = data[16 * nb:]
        if len(tail) > 8:
            k2 = int.from_bytes(tail[8:], "little")
            k2 = k2 * c2 % M; k2 = rotl(k2, 33); k2 = k2 * c1 % M; h2 ^= k2
        if tail:
            k1 = int.from_bytes(tail[:8], "little")
            k1 = k1 * c1 % M; k1 = rotl(k1, 31); k1 = k1 * c2 % M; h1 ^= k1
        h1 ^= len(data); h2 ^= len(data)
        h1 = (h1 + h2) % M; h2 = (h2 + h1) % M
        h1 = fmix(h1); h2 = fmix(h2)
        h1 = (h1 + h2) % M; h2 = (h2 + h1) % M
        return h1, h2

    rng = np.random.default_rng(1)
    for n in range(49):
        data = bytes(rng.integers(0, 256, size=n, dtype=np.uint8))
        assert oracle.murmur3_x64_128(data, 42) == mm(data, 42), n


def test_hash_rc_and_lowercase():
    # test_basic.py:35-40, test_kmers_and_hashes.py:53-64
    assert oracle.hash_kmer("AAA") == oracle.hash_kmer("TTT")
    assert oracle.hash_kmer("acgttg") == oracle.hash_kmer("ACGTTG") == oracle.hash_kmer("CAACGT")


def test_hash_kmer_errors():
    t = OracleTable(4)
    with pytest.raises(RuntimeError):
        t.hash_kmer("ATC")  # src/lib.rs:66-67
    with pytest.raises(RuntimeError):
        t.hash_kmer("ATCN")  # src/lib.rs:79
    with pytest.raises(ValueError):
        t.count("ATC")  # test_basic.py:43-52
    with pytest.raises(ValueError):
        t.get("ATCGG")


def test_count_get_roundtrip():
    # test_basic.py:15-32, 112-127
    t = OracleTable(4)
    assert t.get("ATCG") == 0
    assert t.count("ATCG") == 1
    assert t.get("ATCG") == 1
    kmer = "TAAACCCTAACCCTAACCCTAACCCTAACCC"
    t = OracleTable(31)
    h = t.hash_kmer(kmer)
    assert t.count(kmer) == 1 and t.count(kmer) == 2 and t.count_hash(h) == 3 and t.get(kmer) == 3


def test_consume_basic():
    # test_basic.py:55-72
    t = OracleTable(4)
    assert t.consume("ATCG") == 1
    t = OracleTable(4)
    assert t.consume("ATCGG") == 2
    assert t.get("ATCG") == 1 and t.get("TCGG") == 1 and t.get("CCGA") == 1


def test_consume_error_mode():
    # test_basic.py:75-88, doc/api.md:52-55,77-78
    t = OracleTable(4)
    with pytest.raises(ValueError, match="bad k-mer encountered at position 2"):
        t.consume("ATCGGX", skip_bad_kmers=False)
    assert t.get("ATCG") == 1 and t.get("TCGG") == 1  # counted before the error stay counted
    assert t.consumed == 0  # early return precedes the consumed update (src/lib.rs:595 vs 604)
    t = OracleTable(4)
    with pytest.raises(ValueError, match="bad k-mer encountered at position 0"):
        t.consume("XATCGG", skip_bad_kmers=False)
    assert len(t) == 0


def test_consume_skip_mode():
    # test_basic.py:91-108, test_kmers_and_hashes.py:255-283
    t = OracleTable(4)
    assert t.consume("XATCGG") == 2
    assert t.get("ATCG") == 1 and t.get("TCGG") == 1 and t.get("CCGA") == 1
    t = OracleTable(3)
    assert t.consume("XAAAAAXGGGG") == 5
    assert len(t) == 2
    assert t.get_hash(10679328328772601858) == 3 and t.get_hash(12126843654075378313) == 2


def test_consumed_and_short_reads():
    # test_attr.py:60-114, test_dunders.py:19-25
    t = OracleTable(4)
    assert t.consume("ACG") == 0 and len(t) == 0 and t.consumed == 3
    t.consume("ACGTN")
    assert t.consumed == 8
    t.count("AAAA")
    assert t.consumed == 12


def test_large_repeat():
    # test_add.py:112-125: "ATGC"*100000 -> 399,996 5-mers
    t = OracleTable(5)
    assert t.consume("ATGC" * 100000) == 399996
    assert t.sum_counts == 399996


def test_histo_min_max():
    # test_histo.py
    t = OracleTable(4)
    assert t.min == 0 and t.max == 0 and t.histo(zero=False) == [] and t.histo(zero=True) == [(0, 0)]
    t.count("AAAA"); t.count("TTTT"); t.consume("CCCCCC")
    assert t.min == 2 and t.max == 3
    t = OracleTable(4)
    for k in ("AAAA", "AAAA", "TTTT", "CCCC"):
        t.count(k)
    assert t.histo(zero=False) == [(1, 1), (3, 1)]
    assert t.histo(zero=True) == [(0, 0), (1, 1), (2, 0), (3, 1)]


def test_setops_and_jaccard():
    # test_setops.py, test_metrics.py:173-272
    a, b = OracleTable(4), OracleTable(4)
    assert a.jaccard(b) == 1.0
    for k in ("AAAA", "AATT", "GGGG"):
        a.count(k)
    for k in ("AATT", "GGGG", "CCAA"):  # GGGG == CCCC canonical
        b.count(k)
    assert a.jaccard(b) == 2 / 4
    assert a.intersection(b) == {a.hash_kmer("AATT"), a.hash_kmer("GGGG")}
    assert len(a.union(b)) == 4 and a.setop_sizes(b) == (2, 4)
    assert a.difference(b) == {a.hash_kmer("AAAA")}
    assert a.symmetric_difference(b) == {a.hash_kmer("AAAA"), a.hash_kmer("CCAA")}
    c = OracleTable(4)
    c["ACGT"] = 0  # zero-valued key is still a member (test_metrics.py:150-151)
    assert a.jaccard(c) == 0.0 and len(c) == 1


def test_example_fa_goldens(example_seq, goldens):
    # README.md:94-99, doc/api.md:16-25 (k-mer totals) + derived digests
    for k in ("21", "31"):
        g = goldens["derived"][k]
        t = OracleTable(int(k))
        assert t.consume(example_seq) == g["n_kmers"] == goldens["example_fa_kmers"][k]
        d = t.digest()
        assert (d["n"], d["sum"], d["xor"], d["sum_hc"]) == (g["distinct"], g["sum"], g["xor"], g["sum_hc"])
        assert [list(x) for x in t.histo(zero=False)] == g["histo"]
        assert t.sha256_sorted() == g["sha256"]
    # survey-recorded constants, independent of the json written by the oracle
    t = OracleTable(21); t.consume(example_seq)
    assert t.sha256_sorted() == "88a0e3266782949111190bbac408f7b35f21e63cb28de4f67c1555279aa8e26a"
    assert t.hash_kmer("AAATCTTATAAAATAACCACA") == 14908242140577922293


def test_batch_equals_per_read_and_threads():
    rng = np.random.default_rng(7)
    bases, offs = ragged_batch(rng, 3000, 200, p_bad=0.01)
    for k in (4, 21, 31, 33):
        a, b, c = OracleTable(k), OracleTable(k), OracleTable(k)
        tot = 0
        for r in range(len(offs) - 1):
            tot += a.consume(bases[int(offs[r]):int(offs[r + 1])].tobytes())
        tb, er, _ = b.consume_batch(bases, offs, True, nthreads=1)
        tc, _, _ = c.consume_batch(bases, offs, True, nthreads=4)
        assert tot == tb == tc and er == -1
        assert a.sha256_sorted() == b.sha256_sorted() == c.sha256_sorted()
        assert a.consumed == b.consumed == int(offs[-1])


def test_batch_error_mode_stops_at_first_bad_read():
    k = 5
    reads = [b"ACGTACGT", b"ACG", b"NN", b"ACGTANGTACGT", b"ACGTACGT"]
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(r) for r in reads])
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8)
    t = OracleTable(k)
    total, er, ep = t.consume_batch(bases, offs, skip_bad_kmers=False)
    # read 2 ("NN") is shorter than k: no window, no error.  read 3 fails at its window 1.
    assert (total, er, ep) == (4 + 1, 3, 1)
    ref = OracleTable(k)
    ref.consume("ACGTACGT"); ref.consume("ACGTA")
    assert t.sha256_sorted() == ref.sha256_sorted()


def test_merge_counts():
    # test_add.py:6-34
    a, b = OracleTable(5), OracleTable(5)
    a.consume("ATGCATGC"); b.consume("CATGGCATG")
    before = len(a)
    added, new = a.add(b)
    assert added == b.sum_counts and len(a) == before + new
    assert a.get_hash(3442404512935954368) >= 1  # CATGG, test_add.py:33

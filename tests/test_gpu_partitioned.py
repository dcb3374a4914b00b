"""The partitioned counting pipeline (pass A: hash + scatter into per-partition fragments,
pass B: aggregate duplicates in shared memory + merge) against the CPU oracle and against the
fused kernel.  The pipeline is chosen by launch size in production; here it is forced, so
that inputs small enough for the oracle go through it.  Bit-exact."""
import numpy as np
import pytest

from oracle import OracleTable
from synth import ragged_batch, synth_reads, uniform_offsets

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from oxli_b200 import _capi

    assert _capi.lib.oxg_device_count() > 0, "GPU tests need a CUDA device"
    return _capi


@pytest.fixture()
def part(capi):
    """force the partitioned pipeline for one test; the knobs go back to automatic afterwards"""
    def choose(n_parts=0, groups=0):
        capi.set_pipeline("part", n_parts, groups)
    choose()
    yield choose
    capi.set_pipeline("auto")


def assert_same_table(gpu, ora: OracleTable):
    gk, gv = gpu.export(1)
    ok, ov = ora.items_sorted()
    assert len(gk) == len(ok), (len(gk), len(ok))
    assert np.array_equal(gk, ok)
    assert np.array_equal(gv, ov)


@pytest.mark.parametrize("k", [15, 21, 31, 32, 33, 51, 63])
def test_ragged_batch_skip_mode(capi, part, k):
    rng = np.random.default_rng(300 + k)
    bases, offs = ragged_batch(rng, 4000, 260, p_bad=0.01)
    ora = OracleTable(k)
    want_total, _, _ = ora.consume_batch(bases, offs, True, nthreads=4)
    t = capi.Table(k)
    st, total, er, _ = t.consume_batch(bases, offs, True)
    assert (st, total, er) == (0, want_total, -1)
    a, b = t.last_consume_pass_ms()
    assert a > 0 and b > 0, "the partitioned pipeline did not run"
    assert_same_table(t, ora)


@pytest.mark.parametrize("n_parts,groups", [(2, 1), (64, 1), (1024, 1), (2048, 1), (4096, 1), (256, 4), (8192, 3)])
def test_every_partition_geometry(capi, part, n_parts, groups):
    part(n_parts, groups)
    bases = synth_reads(30_000, 150, 200_000, seed=7, sub_ppm=10_000, n_ppm=1_000)
    offs = uniform_offsets(30_000, 150)
    for k in (21, 31):
        ora = OracleTable(k)
        want, _, _ = ora.consume_batch(bases, offs, True, nthreads=4)
        t = capi.Table(k)
        st, total, _, _ = t.consume_batch(bases, offs, True)
        assert (st, total) == (0, want)
        assert_same_table(t, ora)
        assert t.histo() == ora.histo(zero=False)


def test_high_coverage_duplicates_meet_in_shared_memory(capi, part):
    # 60x coverage of a tiny genome: almost every occurrence is a duplicate inside the launch
    bases = synth_reads(40_000, 150, 100_000, seed=11)
    offs = uniform_offsets(40_000, 150)
    ora = OracleTable(31)
    want, _, _ = ora.consume_batch(bases, offs, True, nthreads=4)
    t = capi.Table(31, capacity_hint=100_000)
    st, total, _, _ = t.consume_batch(bases, offs, True)
    assert (st, total) == (0, want)
    assert_same_table(t, ora)
    # twice the same batch doubles every count
    t.consume_batch(bases, offs, True)
    k1, v1 = t.export(1)
    ok, ov = ora.items_sorted()
    assert np.array_equal(k1, ok) and np.array_equal(v1, 2 * ov)


@pytest.mark.parametrize("seq", [b"A" * 150, b"AT" * 75, b"ACG" * 50, b"T" * 150])
def test_low_complexity_floods_one_partition(capi, part, seq):
    # every window of every read is one of a handful of k-mers: fragments overflow into the spill list
    n = 20_000
    bases = np.frombuffer(seq * n, dtype=np.uint8)
    offs = uniform_offsets(n, len(seq))
    for k in (21, 31):
        ora = OracleTable(k)
        want, _, _ = ora.consume_batch(bases, offs, True, nthreads=4)
        t = capi.Table(k)
        st, total, _, _ = t.consume_batch(bases, offs, True)
        assert (st, total) == (0, want)
        assert_same_table(t, ora)


def test_growth_from_tiny_table_and_singletons(capi, part):
    # nothing hinted, nearly every k-mer distinct: the table runs into its load limit inside
    # pass B, defers (key, count) pairs, grows and replays
    bases = synth_reads(20_000, 150, 50_000_000, seed=3)
    offs = uniform_offsets(20_000, 150)
    ora = OracleTable(21)
    want, _, _ = ora.consume_batch(bases, offs, True, nthreads=4)
    t = capi.Table(21)
    st, total, _, _ = t.consume_batch(bases, offs, True)
    assert (st, total) == (0, want)
    assert_same_table(t, ora)


def test_direct_mode_when_nothing_repeats(capi, part):
    # a hint far above the launch size tells pass B that nearly every key is new: it then updates
    # the table directly instead of going through its shared-memory table
    bases = synth_reads(30_000, 150, 80_000_000, seed=21, sub_ppm=5_000, n_ppm=500)
    offs = uniform_offsets(30_000, 150)
    for k in (21, 31):
        ora = OracleTable(k)
        want, _, _ = ora.consume_batch(bases, offs, True, nthreads=4)
        t = capi.Table(k, capacity_hint=50_000_000)
        st, total, _, _ = t.consume_batch(bases, offs, True)
        assert (st, total) == (0, want)
        st, total, _, _ = t.consume_batch(bases[: 150 * 10_000], offs[:10_001], True)  # a second batch on top
        want2, _, _ = ora.consume_batch(bases[: 150 * 10_000], offs[:10_001], True, nthreads=4)
        assert (st, total) == (0, want2)
        assert_same_table(t, ora)


def test_error_mode_goes_through_the_same_pipeline(capi, part):
    rng = np.random.default_rng(9)
    clean, offs = ragged_batch(rng, 2000, 220, p_bad=0.0, p_empty=0.05)
    bad = np.frombuffer(b"ACGT" * 30, dtype=np.uint8).copy()
    bad[61] = ord("N")
    bases = np.concatenate([clean, bad, clean[:5000]])
    offsets = np.concatenate([offs, [offs[-1] + len(bad)], [offs[-1] + len(bad) + 5000]]).astype(np.uint64)
    ora = OracleTable(31)
    want = ora.consume_batch(bases, offsets, skip_bad_kmers=False)
    t = capi.Table(31)
    st, total, er, ep = t.consume_batch(bases, offsets, skip_bad=False)
    assert st == capi.ERR_BAD_KMER and (total, er, ep) == want
    assert_same_table(t, ora)


def test_partitioned_equals_fused_at_size(capi):
    # 4 M reads: too large for the oracle in a test, large enough for several launches and for
    # the automatic choice; the two pipelines must build the same table
    n, L, k = 4_000_000, 150, 31
    d_bases = capi.device_alloc(n * L + 64)
    d_offs = capi.device_alloc((n + 1) * 8)
    capi.synth_reads_device(d_bases, n, L, 2_000_000, 0xC20001, sub_ppm=2000, n_ppm=500)
    capi.h2d(d_offs, np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    tables = {}
    try:
        for name in ("fused", "part", "auto"):
            capi.set_pipeline(name)
            t = capi.Table(k)
            st, total, _, _ = t.consume_batch_device(d_bases, d_offs, n, n * L, True)
            assert st == 0
            a, b = t.last_consume_pass_ms()
            assert (a > 0) == (name != "fused")
            tables[name] = (total, t.digest(), t.histo())
    finally:
        capi.set_pipeline("auto")
        capi.device_free(d_bases); capi.device_free(d_offs)
    assert tables["fused"] == tables["part"] == tables["auto"]
    assert tables["part"][1]["sum"] == tables["part"][0]

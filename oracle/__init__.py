"""ctypes front-end for the CPU oracle (oracle/oxli_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from oxli_b200/.
Parity status: PINNED (see oxli_oracle.h).

`OracleTable` mirrors the subset of the reference `oxli.KmerCountTable` that
lies on the hot path (src/lib.rs:29-62, 65-81, 100-104, 145-194, 464-539,
545-607, 610-638, 708-722, 778-837) so that parity tests read like the
reference's own tests.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboxli_oracle.so")

u64 = C.c_uint64
u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oxli_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def _load():
    lib = C.CDLL(build())
    sig = {
        "oxo_murmur3_x64_128": (None, [C.c_char_p, C.c_size_t, C.c_uint32, u64p]),
        "oxo_murmur3_smhasher_verification": (C.c_uint32, []),
        "oxo_hash_kmer": (C.c_int, [C.c_char_p, C.c_size_t, C.c_uint32, u64p]),
        "oxo_hash_windows": (None, [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]),
        "oxo_table_new": (C.c_void_p, []),
        "oxo_table_free": (None, [C.c_void_p]),
        "oxo_table_clear": (None, [C.c_void_p]),
        "oxo_table_len": (u64, [C.c_void_p]),
        "oxo_table_count_hash": (u64, [C.c_void_p, u64]),
        "oxo_table_add_hash": (None, [C.c_void_p, u64, u64]),
        "oxo_table_get_hash": (u64, [C.c_void_p, u64]),
        "oxo_table_set_hash": (None, [C.c_void_p, u64, u64]),
        "oxo_table_contains": (C.c_int, [C.c_void_p, u64]),
        "oxo_table_drop_hash": (C.c_int, [C.c_void_p, u64]),
        "oxo_table_mincut": (u64, [C.c_void_p, u64]),
        "oxo_table_maxcut": (u64, [C.c_void_p, u64]),
        "oxo_table_min": (u64, [C.c_void_p]),
        "oxo_table_max": (u64, [C.c_void_p]),
        "oxo_table_sum": (u64, [C.c_void_p]),
        "oxo_consume": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32, C.c_int, u64p]),
        "oxo_consume_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, u64, C.c_uint32, C.c_int,
                                        C.c_int, u64p, C.POINTER(C.c_int64), u64p]),
        "oxo_table_export_sorted": (u64, [C.c_void_p, C.c_void_p, C.c_void_p, u64]),
        "oxo_table_histo_sparse": (u64, [C.c_void_p, C.c_void_p, C.c_void_p, u64]),
        "oxo_setop_sizes": (None, [C.c_void_p, C.c_void_p, u64p, u64p]),
        "oxo_jaccard": (C.c_double, [C.c_void_p, C.c_void_p]),
        "oxo_table_merge": (None, [C.c_void_p, C.c_void_p, u64p, u64p]),
        "oxo_table_digest": (None, [C.c_void_p, u64p, u64p, u64p, u64p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


OK, ERR_WRONG_KSIZE, ERR_BAD_KMER = 0, 1, 2


def murmur3_x64_128(data: bytes, seed: int = 42) -> tuple[int, int]:
    out = (u64 * 2)()
    lib().oxo_murmur3_x64_128(data, len(data), seed, out)
    return int(out[0]), int(out[1])


def smhasher_verification() -> int:
    return int(lib().oxo_murmur3_smhasher_verification())


def hash_kmer(kmer: str | bytes, ksize: int | None = None) -> int:
    """src/lib.rs:65-81.  RuntimeError on wrong length / non-ACGT, as pyo3's
    anyhow conversion produces."""
    b = kmer.encode() if isinstance(kmer, str) else bytes(kmer)
    k = len(b) if ksize is None else ksize
    out = u64()
    rc = lib().oxo_hash_kmer(b, len(b), k, C.byref(out))
    if rc == ERR_WRONG_KSIZE:
        raise RuntimeError("wrong ksize")
    if rc == ERR_BAD_KMER:
        raise RuntimeError("invalid DNA character in k-mer")
    return int(out.value)


def hash_windows(seq: bytes | np.ndarray, ksize: int) -> np.ndarray:
    """Per-window hashes, 0 for a bad window (sourmash SeqToHashes force=true)."""
    a = np.frombuffer(seq, dtype=np.uint8) if not isinstance(seq, np.ndarray) else np.ascontiguousarray(seq, dtype=np.uint8)
    n = max(len(a) - ksize + 1, 0)
    out = np.zeros(n, dtype=np.uint64)
    if n:
        lib().oxo_hash_windows(a.ctypes.data, len(a), ksize, out.ctypes.data)
    return out


class OracleTable:
    """CPU stand-in for the reference `oxli.KmerCountTable` (hot-path subset)."""

    def __init__(self, ksize: int):
        if not 0 <= ksize <= 255:
            raise OverflowError("ksize out of range for u8")
        self.ksize = ksize
        self.consumed = 0
        self._t = lib().oxo_table_new()

    def __del__(self):
        if getattr(self, "_t", None) and _lib is not None:
            _lib.oxo_table_free(self._t)
            self._t = None

    # -- single k-mer paths (src/lib.rs:65-81, 145-194) --
    def hash_kmer(self, kmer: str) -> int:
        return hash_kmer(kmer, self.ksize)

    def count_hash(self, h: int) -> int:
        return int(lib().oxo_table_count_hash(self._t, h))

    def count(self, kmer: str) -> int:
        if len(kmer.encode()) % 256 != self.ksize:  # `kmer.len() as u8` (src/lib.rs:146)
            raise ValueError("kmer size does not match count table ksize")
        c = self.count_hash(self.hash_kmer(kmer))
        self.consumed += len(kmer)
        return c

    def get(self, kmer: str) -> int:
        if len(kmer.encode()) % 256 != self.ksize:
            raise ValueError("kmer size does not match count table ksize")
        return self.get_hash(self.hash_kmer(kmer))

    def get_hash(self, h: int) -> int:
        return int(lib().oxo_table_get_hash(self._t, h))

    def get_hash_array(self, hs) -> list[int]:
        return [self.get_hash(h) for h in hs]

    def __getitem__(self, kmer):
        return self.get(kmer)

    def __setitem__(self, kmer, v):
        lib().oxo_table_set_hash(self._t, self.hash_kmer(kmer), v)

    def set_hash(self, h, v):
        lib().oxo_table_set_hash(self._t, h, v)

    def contains(self, h) -> bool:
        return bool(lib().oxo_table_contains(self._t, h))

    def drop_hash(self, h):
        lib().oxo_table_drop_hash(self._t, h)

    def drop(self, kmer):
        self.drop_hash(self.hash_kmer(kmer))

    def mincut(self, m):
        return int(lib().oxo_table_mincut(self._t, m))

    def maxcut(self, m):
        return int(lib().oxo_table_maxcut(self._t, m))

    def __len__(self):
        return int(lib().oxo_table_len(self._t))

    # -- consume (src/lib.rs:545-607) --
    def consume(self, seq: str | bytes, skip_bad_kmers: bool = True) -> int:
        b = seq.encode() if isinstance(seq, str) else bytes(seq)
        n = u64()
        buf = C.create_string_buffer(b, len(b)) if b else None
        rc = lib().oxo_consume(self._t, C.cast(buf, C.c_void_p) if buf else None, len(b), self.ksize,
                               1 if skip_bad_kmers else 0, C.byref(n))
        if rc == ERR_BAD_KMER:
            raise ValueError(f"bad k-mer encountered at position {n.value}")
        self.consumed += len(b)
        return int(n.value)

    def consume_batch(self, bases: np.ndarray, offsets: np.ndarray, skip_bad_kmers: bool = True,
                      nthreads: int = 1):
        """CSR batch; returns (total, err_read, err_pos)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        total, er, ep = u64(), C.c_int64(), u64()
        lib().oxo_consume_batch(self._t, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                                self.ksize, 1 if skip_bad_kmers else 0, nthreads, C.byref(total),
                                C.byref(er), C.byref(ep))
        if er.value < 0:
            self.consumed += int(offsets[-1] - offsets[0])
        else:
            self.consumed += int(offsets[er.value] - offsets[0])
        return int(total.value), int(er.value), int(ep.value)

    # -- scans (src/lib.rs:464-539) --
    @property
    def min(self):
        return int(lib().oxo_table_min(self._t))

    @property
    def max(self):
        return int(lib().oxo_table_max(self._t))

    @property
    def sum_counts(self):
        return int(lib().oxo_table_sum(self._t))

    def items_sorted(self) -> tuple[np.ndarray, np.ndarray]:
        n = len(self)
        k = np.empty(max(n, 1), dtype=np.uint64)
        v = np.empty(max(n, 1), dtype=np.uint64)
        if n:
            lib().oxo_table_export_sorted(self._t, k.ctypes.data, v.ctypes.data, n)
        return k[:n], v[:n]

    @property
    def hashes(self) -> list[int]:
        return [int(x) for x in self.items_sorted()[0]]

    def histo(self, zero: bool = True) -> list[tuple[int, int]]:
        n = len(self)
        f = np.empty(max(n, 1), dtype=np.uint64)
        c = np.empty(max(n, 1), dtype=np.uint64)
        m = int(lib().oxo_table_histo_sparse(self._t, f.ctypes.data, c.ctypes.data, max(n, 1))) if n else 0
        sparse = [(int(f[i]), int(c[i])) for i in range(m)]
        if not zero:
            return sparse
        d = dict(sparse)
        return [(i, d.get(i, 0)) for i in range(self.max + 1)]

    # -- set comparisons (src/lib.rs:610-638, 708-722) --
    def setop_sizes(self, other: "OracleTable") -> tuple[int, int]:
        i, u = u64(), u64()
        lib().oxo_setop_sizes(self._t, other._t, C.byref(i), C.byref(u))
        return int(i.value), int(u.value)

    def jaccard(self, other: "OracleTable") -> float:
        return float(lib().oxo_jaccard(self._t, other._t))

    def hash_set(self) -> set[int]:
        return set(self.hashes)

    def union(self, o):
        return self.hash_set() | o.hash_set()

    def intersection(self, o):
        return self.hash_set() & o.hash_set()

    def difference(self, o):
        return self.hash_set() - o.hash_set()

    def symmetric_difference(self, o):
        return self.hash_set() ^ o.hash_set()

    def add(self, other: "OracleTable") -> tuple[int, int]:
        if self.ksize != other.ksize:
            raise ValueError("KmerCountTables must have the same ksize")
        a, n = u64(), u64()
        lib().oxo_table_merge(self._t, other._t, C.byref(a), C.byref(n))
        self.consumed += other.consumed
        return int(a.value), int(n.value)

    def digest(self) -> dict:
        n, s, x, hc = u64(), u64(), u64(), u64()
        lib().oxo_table_digest(self._t, C.byref(n), C.byref(s), C.byref(x), C.byref(hc))
        return {"n": int(n.value), "sum": int(s.value), "xor": int(x.value), "sum_hc": int(hc.value)}

    def sha256_sorted(self) -> str:
        import hashlib

        k, v = self.items_sorted()
        inter = np.empty(2 * len(k), dtype="<u8")
        inter[0::2], inter[1::2] = k, v
        return hashlib.sha256(inter.tobytes()).hexdigest()

"""TEST INFRASTRUCTURE. numpy mirror of the library's synthetic read generator (tableops.cuh:
synth_reads_kernel) plus ragged test-batch builders.  Test helper."""
from __future__ import annotations

import numpy as np

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def genome_base(seed: int, pos: np.ndarray) -> np.ndarray:
    pos = np.asarray(pos, dtype=np.uint64)
    with np.errstate(over="ignore"):
        w = splitmix64(np.uint64(seed) + (pos >> np.uint64(5)))
    return ((w >> (np.uint64(2) * (pos & np.uint64(31)))) & np.uint64(3)).astype(np.uint32)


def synth_reads(n_reads: int, read_len: int, genome_len: int, seed: int, first_read: int = 0,
                sub_ppm: int = 0, n_ppm: int = 0) -> np.ndarray:
    """uint8 array of n_reads*read_len ASCII bases, identical to oxg_synth_reads_device."""
    step = 100_000
    if n_reads > step:  # bound the temporaries
        return np.concatenate([synth_reads(min(step, n_reads - a), read_len, genome_len, seed, first_read + a,
                                           sub_ppm, n_ppm) for a in range(0, n_reads, step)])
    r = np.arange(n_reads, dtype=np.uint64)
    with np.errstate(over="ignore"):
        rk = splitmix64((np.uint64(seed) ^ np.uint64(0x5EEDF00D)) + (np.uint64(first_read) + r) * np.uint64(0x2545F4914F6CDD1D))
        start = rk % np.uint64(genome_len - read_len + 1)
        rev = (splitmix64(rk + np.uint64(1)) & np.uint64(1)).astype(bool)
        j = np.arange(read_len, dtype=np.uint64)[None, :]
        fw = genome_base(seed, start[:, None] + j)
        rc = np.uint32(3) - genome_base(seed, start[:, None] + np.uint64(read_len - 1) - j)
        b = np.where(rev[:, None], rc, fw).astype(np.uint32)
        e = splitmix64(rk[:, None] + np.uint64(2) + j)
        sub = (e % np.uint64(1000000)).astype(np.uint32) < np.uint32(sub_ppm)
        b = np.where(sub, (b + np.uint32(1) + ((e >> np.uint64(32)) % np.uint64(3)).astype(np.uint32)) & np.uint32(3), b)
        isn = ((e >> np.uint64(20)) % np.uint64(1000000)).astype(np.uint32) < np.uint32(n_ppm)
    out = np.frombuffer(b"ACGT", dtype=np.uint8)[b]
    out = np.where(isn, np.uint8(ord("N")), out)
    return np.ascontiguousarray(out.reshape(-1))


def uniform_offsets(n_reads: int, read_len: int) -> np.ndarray:
    return np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(read_len)


def ragged_batch(rng: np.random.Generator, n_reads: int, max_len: int, p_bad: float = 0.01,
                 p_lower: float = 0.1, p_empty: float = 0.05, alphabet: bytes = b"ACGT",
                 bad_chars: bytes = b"NXRYn-*\x00\xff.") -> tuple[np.ndarray, np.ndarray]:
    """Reads of random length (some empty / shorter than k) with stray non-ACGT bytes
    and lower-case runs.  Returns (bases uint8, offsets uint64)."""
    lens = rng.integers(0, max_len + 1, size=n_reads)
    lens[rng.random(n_reads) < p_empty] = 0
    offsets = np.zeros(n_reads + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lens)
    total = int(offsets[-1])
    bases = np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=total)].copy()
    lower = rng.random(total) < p_lower
    bases[lower] |= 0x20
    bad = rng.random(total) < p_bad
    bases[bad] = np.frombuffer(bad_chars, dtype=np.uint8)[rng.integers(0, len(bad_chars), size=int(bad.sum()))]
    return bases, offsets

"""Host side of the synthetic input generator: the counter hash and the cfg-5 read-length
profile, in numpy, so that benches and tools can lay out batches without touching the GPU.  The bases
themselves come from the device generator (``bitnuc_b200.device.synth_ascii`` / ``synth_words``):
word j of stream s = splitmix64((seed ^ s * 0x9E3779B97F4A7C15) + j), base 32 j + i = "ACGT"[(W >> 2 i) & 3].
"""
from __future__ import annotations

import numpy as np

DEFAULT_SEED = 0x5EEDB17C0DE5


def splitmix64(x) -> np.ndarray:
    """splitmix64 finaliser of a uint64 array (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def cfg5_read_lengths(total_bases: int, seed: int = DEFAULT_SEED) -> np.ndarray:
    """Read lengths 50 + (splitmix64(seed + 5 + r) mod 9951) (uniform 50 bp .. 10 kbp) until their sum
    reaches ``total_bases`` (BASELINE.json configs[4])."""
    n_guess = int(total_bases / 5025 * 1.02) + 16
    while True:
        r = np.arange(n_guess, dtype=np.uint64)
        lens = (np.uint64(50) + splitmix64(r + np.uint64(seed + 5)) % np.uint64(9951)).astype(np.uint64)
        cum = np.cumsum(lens)
        if int(cum[-1]) >= total_bases:
            return lens[: int(np.searchsorted(cum, total_bases)) + 1]
        n_guess *= 2


def cfg5_injected_n(n_reads: int, lens: np.ndarray):
    """Reads that get an 'N' (h1(r) mod 100003 == 0) and the position inside each (h2(r) mod len)."""
    r = np.arange(n_reads, dtype=np.uint64)
    victims = np.flatnonzero(splitmix64(r + np.uint64(0xABCDEF)) % np.uint64(100003) == 0)
    pos = (splitmix64(victims.astype(np.uint64) + np.uint64(0x123457)) % lens[victims]).astype(np.int64)
    return victims, pos

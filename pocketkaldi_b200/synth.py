"""Deterministic synthetic inputs for parity tests and benchmarks.

PCM is a counter-based stream keyed by (seed, utterance id, sample index), so
any shard of a corpus can be regenerated on any host or GPU rank bit-for-bit
(SURVEY.md section 8d). Only integer arithmetic is used: four 16-bit uniforms
from a splitmix64 hash are summed (Irwin-Hall, n = 4) and scaled to sigma ~ 3000,
which keeps |sample| <= 10392 (no clipping) and is identical in numpy and in the
CUDA generator (synth_pcm_kernel in pocketkaldi_b200/csrc/cmvn.cu).
"""

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
PCM_SCALE = 5196  # (3000 / sigma(sum of 4 u16)) * 65536


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def synth_pcm(seed, utt_ids, n_samples):
    """int16 [len(utt_ids)][n_samples]; sample (u, i) depends only on (seed, u, i)."""
    with np.errstate(over="ignore"):
        utt = np.asarray(utt_ids, dtype=np.uint64).reshape(-1, 1)
        key = _mix64(np.uint64(seed) * _GOLD + utt)
        idx = np.arange(n_samples, dtype=np.uint64).reshape(1, -1)
        h = _mix64(key + idx * _GOLD)
    m = np.uint64(0xFFFF)
    s = ((h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m)
         + (h >> np.uint64(48))).astype(np.int64) - 2 * 65535
    return ((s * PCM_SCALE) >> 16).astype(np.int16)


def synth_global_cmvn(dim=40, count=36162480.0, mean=17.0):
    """A plausible global CMVN stats vector (dim sums + count) when none is supplied."""
    g = np.empty(dim + 1, np.float32)
    g[:dim] = np.float32(mean * count) * (1.0 + 0.01 * np.arange(dim, dtype=np.float32))
    g[dim] = count
    return g

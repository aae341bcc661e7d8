"""Seeded synthetic inputs of the shapes BASELINE.json names.

There is no dataset in the reference (its ``Data/`` tree is git-ignored and private),
so every test and benchmark uses these generators.  All take an explicit
``np.random.Generator`` or seed so the CPU oracle and the CUDA path see identical bits.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


def uniform_sequences(rng, S: int, M: int = 256, tmin: int = 90, tmax: int = 110) -> List[np.ndarray]:
    """Uniform codewords — the worst case for the emission matrix (every symbol appears)."""
    return [rng.integers(0, M, size=int(rng.integers(tmin, tmax + 1))).astype(np.int64)
            for _ in range(S)]


def clustered_sequences(rng, S: int, N: int = 4, M: int = 256, tmin: int = 90, tmax: int = 110,
                        shift: int = 0, spread: int = 40) -> List[np.ndarray]:
    """Left-to-right structured codewords: the utterance is cut into N sorted segments and
    segment s emits ``(M//N)*s + shift + U{0..spread-1} (mod M)`` so training is
    non-trivial and different ``shift`` values give distinguishable words."""
    out = []
    for _ in range(S):
        T = int(rng.integers(tmin, tmax + 1))
        seg = np.sort(rng.integers(0, N, size=T))
        out.append((((M // N) * seg + shift + rng.integers(0, spread, size=T)) % M).astype(np.int64))
    return out


def word_corpus(seed: int, W: int, S: int, N: int = 4, M: int = 256, tmin: int = 90, tmax: int = 110,
                kind: str = "clustered") -> List[List[np.ndarray]]:
    """W words x S sequences.  Word w uses shift = w * (M // (4*W) + 1) * 3 so that the
    words overlap partially (recognition is neither trivial nor hopeless)."""
    rng = np.random.default_rng(seed)
    corpus = []
    for w in range(W):
        if kind == "uniform":
            corpus.append(uniform_sequences(rng, S, M, tmin, tmax))
        else:
            corpus.append(clustered_sequences(rng, S, N, M, tmin, tmax, shift=(w * 7) % M))
    return corpus


def pack_corpus(corpus: List[List[np.ndarray]], M: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Flatten a corpus into the C-ABI layout: (obs, offsets[int64, R+1], word_of_seq[int32, R]).
    obs dtype is uint8 when M <= 256 else uint16."""
    dt = np.uint8 if M <= 256 else np.uint16
    seqs = [s for word in corpus for s in word]
    word_of_seq = np.concatenate([np.full(len(word), w, dtype=np.int32) for w, word in enumerate(corpus)]) \
        if corpus else np.zeros(0, np.int32)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    offsets = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    obs = np.concatenate(seqs).astype(dt) if seqs else np.zeros(0, dt)
    return obs, offsets, word_of_seq


def fixed_length_codewords(seed: int, W: int, S: int, T: int, N: int = 4, M: int = 256,
                           dtype=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Vectorised generator for the large benchmark configs (config 3/4/5): W words x S
    sequences of fixed length T, left-to-right structure (N segments with random cut points,
    segment s emits (M//N)*s + U{0..M/(2N)+7} + 7*w mod M), already packed.  Sequence r of
    word w is row w*S + r."""
    rng = np.random.default_rng(seed)
    dt = dtype or (np.uint8 if M <= 256 else np.uint16)
    wide = np.uint16 if M <= 256 else np.uint32
    R = W * S
    obs = np.empty((R, T), dtype=dt)
    spread = max(M // (2 * N), 1) + 8
    tgrid = np.arange(T, dtype=np.int32)[None, :]
    chunk = max(1, (1 << 24) // max(T, 1))
    for lo in range(0, R, chunk):
        hi = min(R, lo + chunk)
        cuts = np.sort(rng.integers(0, T + 1, size=(hi - lo, N - 1), dtype=np.int32), axis=1) if N > 1 \
            else np.zeros((hi - lo, 0), np.int32)
        seg = np.zeros((hi - lo, T), dtype=wide)
        for c in range(N - 1):
            seg += (tgrid >= cuts[:, c:c + 1]).astype(wide)
        sym = seg * wide(M // N) + rng.integers(0, spread, size=(hi - lo, T), dtype=wide)
        shift = ((np.arange(lo, hi) // S) * 7).astype(wide)[:, None]
        obs[lo:hi] = ((sym + shift) % wide(M)).astype(dt)
    offsets = np.arange(R + 1, dtype=np.int64) * T
    word_of_seq = np.repeat(np.arange(W, dtype=np.int32), S)
    return obs.reshape(-1), offsets, word_of_seq


def mfcc_mixture(seed: int, F: int, K: int = 256) -> np.ndarray:
    """F synthetic 13-dim 'MFCC' frames: mixture of K Gaussians with MFCC-like
    per-coefficient scales (coef 0 = energy, sigma ~300; coefs 1..12 from 40 down to 5)."""
    rng = np.random.default_rng(seed)
    scales = np.concatenate([[300.0], np.linspace(40.0, 5.0, 12)])
    means = rng.normal(size=(K, 13)) * scales
    means[:, 0] -= 400.0
    comp = rng.integers(0, K, size=F)
    X = means[comp] + rng.normal(size=(F, 13)) * (scales * 0.35)
    return np.ascontiguousarray(X, dtype=np.float64)


def random_codebook(seed: int, K: int = 256) -> np.ndarray:
    rng = np.random.default_rng(seed)
    scales = np.concatenate([[300.0], np.linspace(40.0, 5.0, 12)])
    C = rng.normal(size=(K, 13)) * scales
    C[:, 0] -= 400.0
    return np.ascontiguousarray(C)

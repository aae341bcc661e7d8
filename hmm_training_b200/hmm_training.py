"""Drop-in for the reference's HMM/hmm_training.py: same function names, arguments, prints,
file side effects and return values; VQ encoding and Baum-Welch run in libhmmb200.so.

    get_observations      HMM/hmm_training.py:82-120
    hmm_training          HMM/hmm_training.py:265-541
    training_with_save    HMM/hmm_training.py:215-247
plus ``hmm_training_batched`` (all words in one launch — the reference's caller
``train_hmm`` loops words serially, HMM/main.py:147-154) which is what exposes the GPU's
parallelism, and log-space helpers with the reference's semantics.
"""
from __future__ import annotations

import json
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib, engine
from .codevector_classes import CentroidDataMFCC, RawDataMFCC, frames_matrix
from .hmm_classes import DataStorageHMM, HMMTrained


def safe_log(x):
    """HMM/hmm_training.py:46-54."""
    if isinstance(x, np.ndarray):
        result = np.full_like(x, float("-inf"), dtype=float)
        mask = x > 0
        result[mask] = np.log(x[mask])
        return result
    return math.log(x) if x > 0 else float("-inf")


def safe_exp(x):
    """HMM/hmm_training.py:56-64."""
    if isinstance(x, np.ndarray):
        result = np.zeros_like(x, dtype=float)
        mask = x != float("-inf")
        result[mask] = np.exp(x[mask])
        return result
    return math.exp(x) if x != float("-inf") else 0.0


def log_sum_exp(log_probs):
    """HMM/hmm_training.py:66-79."""
    if isinstance(log_probs, np.ndarray):
        finite_mask = log_probs != float("-inf")
        if not np.any(finite_mask):
            return float("-inf")
        finite_probs = log_probs[finite_mask]
        max_val = np.max(finite_probs)
        return max_val + math.log(np.sum(np.exp(finite_probs - max_val)))
    return log_probs if log_probs != float("-inf") else float("-inf")


def get_observations(recordings: List[List[RawDataMFCC]], centroids: List[CentroidDataMFCC]) -> list:
    """VQ-encode every frame of every recording (HMM/hmm_training.py:82-120): one
    ``np.ndarray`` of int64 centroid indices per recording.  All recordings go to the GPU
    in a single hmmb_vq_encode call."""
    lens = [len(r) for r in recordings]
    if sum(lens) == 0:
        return [np.array([]) for _ in recordings]  # np.array([]) is what the reference builds for an empty recording
    # a recording is a list of RawDataMFCC (the reference's layout) or a packed [T, 13] matrix
    # (codevector_classes.load_mfcc_matrix: the fast loader that skips the per-frame objects)
    if len(centroids) == 0:
        # the reference's inner loop does not run and every frame keeps closest_centroid_id = 0 (:103)
        return [np.zeros(n, dtype=np.int64) if n else np.array([]) for n in lens]
    X = np.concatenate([frames_matrix(r) for r in recordings if len(r)], axis=0)
    C = frames_matrix(centroids)
    idx = engine.vq_encode(X, C).astype(np.int64)
    out, pos = [], 0
    for n in lens:
        out.append(idx[pos:pos + n].copy() if n else np.array([]))
        pos += n
    return out


def _initial_params(N: int, M: int, word_name: Optional[str], load_initial_params: bool, show_progress: bool):
    """Warm start from ../Data/Eighty-five-percent_20/<word>.json if it matches (N, M), else
    the reference's defaults (HMM/hmm_training.py:270-320), with the same messages."""
    pi = A = B = None
    if load_initial_params and word_name:
        try:
            saved = DataStorageHMM.load_hmm(word_name, "../Data/Eighty-five-percent_20")
            if saved.states == N and saved.symbols == M:
                pi, A, B = saved.Pi.copy(), saved.A.copy(), saved.B.copy()
                if show_progress:
                    print(f"Loaded initial parameters from saved model for word '{word_name}'")
            elif show_progress:
                print(f"Saved model dimensions ({saved.states} states, {saved.symbols} symbols) "
                      f"don't match expected ({N} states, {M} symbols). Using default initialization.")
        except (FileNotFoundError, json.JSONDecodeError, KeyError) as e:
            if show_progress:
                print(f"Could not load saved model for word '{word_name}': {str(e)}. Using default initialization.")
        except Exception as e:
            if show_progress:
                print(f"Unexpected error loading saved model for word '{word_name}': {str(e)}. "
                      f"Using default initialization.")
    if pi is None:
        if N != 4:
            # the reference's literals are 4-state (:301, :307-312); with N != 4 it fails with IndexError
            raise IndexError(f"default initial parameters exist for N == 4 only (got N = {N}); "
                             f"pass init=(pi, A, B) or use the warm-start file")
        pi, A, B = engine.default_init(N, M)
        if show_progress:
            print("Using default initial state probabilities")
            print("Using default transition matrix")
            print("Using default emission matrix")
    return np.asarray(pi, float), np.asarray(A, float), np.asarray(B, float)


def hmm_training_batched(observations_by_word: Sequence[Sequence[np.ndarray]], N: int = 4, M: int = 256,
                         epsilon: float = 1e-6, max_iterations: int = 100, init=None, show_progress: bool = False,
                         word_names: Optional[Sequence[str]] = None, allreduce=None, rank: int = 0, world: int = 1):
    """Train W word models at once.  ``init`` = (pi [W,N], A [W,N,N], B [W,N,M]) or None for
    the reference defaults.  Returns (A [W,N,N], B [W,N,M], pi [W,N], ll_hist [W,max_it], iters [W]);
    every word follows exactly the iteration sequence the reference's per-word loop would."""
    W = len(observations_by_word)
    seqs = [np.asarray(o) for word in observations_by_word for o in word]
    if any(len(o) == 0 for o in seqs):
        raise IndexError("index -1 is out of bounds for axis 1 with size 0")  # hmm_training.py:376
    obs, offsets = _lib.pack_sequences(seqs, M)
    wos = np.concatenate([np.full(len(word), w, dtype=np.int32) for w, word in enumerate(observations_by_word)]) \
        if W else np.zeros(0, np.int32)
    if init is None:
        p, a, b = engine.default_init(N, M)
        if N != 4:
            raise IndexError("default initial parameters exist for N == 4 only")
        init = (np.tile(p, (W, 1)), np.tile(a, (W, 1, 1)), np.tile(b, (W, 1, 1)))
    pi, A, B, hist, iters = engine.bw_fit(obs, offsets, wos, W, N, M, init[0], init[1], init[2], epsilon,
                                          max_iterations, allreduce=allreduce, rank=rank, world=world)
    if show_progress:
        for w in range(W):
            name = word_names[w] if word_names else str(w)
            for it in range(int(iters[w])):
                print(f"[{name}] Iteration {it + 1} Log-likelihood: {hist[w, it]:.6f}")
    return A, B, pi, hist, iters


def hmm_training(observations: List[np.ndarray], N: int = 4, M: int = 256, epsilon: float = 1e-6,
                 max_iterations: int = 100, show_progress=True, word_name: str = None,
                 load_initial_params: bool = True) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Baum-Welch for one word (HMM/hmm_training.py:265-541); returns (A, B, pi) in that order."""
    pi0, A0, B0 = _initial_params(N, M, word_name, load_initial_params, show_progress)
    if max_iterations <= 0:
        # the reference's loop body never runs and its closing print reads a variable the body assigns (:516)
        raise UnboundLocalError("cannot access local variable 'current_log_likelihood_sum' where it is not associated "
                                "with a value")
    A, B, pi, hist, iters = hmm_training_batched([observations], N, M, epsilon, max_iterations,
                                                 init=(pi0[None], A0[None], B0[None]))
    it = int(iters[0])
    prev = float("-inf")
    diff = epsilon + 10
    cur = float("nan")
    for k in range(it):
        cur = float(hist[0, k])
        diff = abs(cur - prev) if prev != float("-inf") else float("inf")
        if show_progress:
            print(f"Iteration {k + 1}")
            print(f"Log-likelihood: {cur:.6f}, Diff: {diff:.8f}")
        prev = cur
    print(f"Log-likelihood: {cur:.6f}, Diff: {diff:.8f}")
    if it >= max_iterations:
        print(f"Reached maximum iterations ({max_iterations})")
    else:
        print(f"Converged after {it} iterations")
    return A[0], B[0], pi[0]


def training_with_save(word_recordings: List[RawDataMFCC], centroids: List[CentroidDataMFCC], word_name: str,
                       max_iterations=100, show_progress=True, load_initial_params=False) -> HMMTrained:
    """HMM/hmm_training.py:215-247: encode, train with N = 4 (hard-coded there, :226), save to
    ../Data/ResultsHMM/<word>.json relative to the CWD."""
    print("Converting recordings to observations...")
    observations = get_observations(word_recordings, centroids)
    print(f"Generated {len(observations)} observation sequences")
    print(f"Sequence lengths: {[len(obs) for obs in observations]}")
    print("Starting Baum-Welch training...")
    A, B, pi = hmm_training(observations, N=4, M=len(centroids), max_iterations=max_iterations,
                            show_progress=show_progress, word_name=word_name,
                            load_initial_params=load_initial_params)
    hmm_model = HMMTrained(states=4, symbols=len(centroids), A=A, B=B, Pi=pi, word=word_name)
    DataStorageHMM.save_hmm(hmm_model, print_messages=False)
    return hmm_model


def train_hmm_batched(recordings_by_word: Dict[str, List[List[RawDataMFCC]]], centroids: List[CentroidDataMFCC],
                      max_iterations: int = 100, show_progress: bool = False, save: bool = True,
                      base_dir: str = "../Data/ResultsHMM", load_initial_params: bool = False) -> List[HMMTrained]:
    """Batched equivalent of the loop in the reference's HMM/main.py:train_hmm (:133-164):
    one VQ-encode call and one Baum-Welch call for the whole vocabulary."""
    words = list(recordings_by_word.keys())
    all_recs = [rec for w in words for rec in recordings_by_word[w]]
    obs = get_observations(all_recs, centroids)
    by_word, pos = [], 0
    for w in words:
        n = len(recordings_by_word[w])
        by_word.append(obs[pos:pos + n])
        pos += n
    M = len(centroids)
    init = None
    if load_initial_params:
        # per-word warm start exactly as hmm_training does it (:270-320), defaults where no file matches
        ps, As, Bs = zip(*[_initial_params(4, M, w, True, show_progress) for w in words])
        init = (np.stack(ps), np.stack(As), np.stack(Bs))
    A, B, pi, hist, iters = hmm_training_batched(by_word, 4, M, 1e-6, max_iterations, init=init,
                                                 show_progress=show_progress, word_names=words)
    models = []
    for i, w in enumerate(words):
        m = HMMTrained(states=4, symbols=M, A=A[i], B=B[i], Pi=pi[i], word=w)
        if save:
            DataStorageHMM.save_hmm(m, base_dir, print_messages=False)
        models.append(m)
    return models

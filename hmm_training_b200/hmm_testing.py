"""Drop-in for the reference's HMM/hmm_testing.py recognition path:

    calculate_log_likelihood   HMM/hmm_testing.py:49-104
    test_hmm                   HMM/hmm_testing.py:107-163
plus ``score_all`` (every utterance against every model in one launch).  The
confusion-matrix plotting (:166-218) is reporting, out of scope.
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _lib, engine
from .codevector_classes import DataStorage, RawDataMFCC
from .hmm_classes import HMMTrained
from .hmm_training import get_observations


def _stack_models(all_hmm: Sequence[HMMTrained]):
    N, M = int(all_hmm[0].states), int(all_hmm[0].symbols)
    for h in all_hmm:
        if int(h.states) != N or int(h.symbols) != M:
            raise ValueError("all models must share (states, symbols) to be scored in one batch")
    pi = np.stack([np.asarray(h.Pi, float) for h in all_hmm])
    A = np.stack([np.asarray(h.A, float) for h in all_hmm])
    B = np.stack([np.asarray(h.B, float) for h in all_hmm])
    return N, M, pi, A, B


def score_all(observations: Sequence[np.ndarray], all_hmm: Sequence[HMMTrained]):
    """[U, W] log-likelihood matrix and the index of the winning model per utterance
    (-1 = "unknown": every score is -inf, hmm_testing.py:161)."""
    N, M, pi, A, B = _stack_models(all_hmm)
    seqs = [np.asarray(o) for o in observations]
    if any(len(o) == 0 for o in seqs):
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")  # hmm_testing.py:75
    obs, offsets = _lib.pack_sequences(seqs, M)
    return engine.score(obs, offsets, N, M, pi, A, B)


def calculate_log_likelihood(recording_observations: np.ndarray, hmm: HMMTrained) -> float:
    """log P(O | lambda) of one recording by the forward algorithm (hmm_testing.py:49-104)."""
    ll, _ = score_all([recording_observations], [hmm])
    return float(ll[0, 0])


def test_hmm(all_hmm: List[HMMTrained], test_recordings_dict: Dict[str, List[List[RawDataMFCC]]],
             base_dir="../Data", show_progress=False) -> Tuple[List[str], List[str]]:
    """hmm_testing.py:107-163: loads <base_dir>/CodeVector/codevector.json itself, encodes every
    test recording, scores it against every model, predicts the first model with the
    strictly largest log-likelihood ("unknown" if all are -inf)."""
    print("Starting HMM testing...")
    storage = DataStorage()
    centroids = storage.load_centroids(os.path.join(base_dir, "CodeVector", "codevector.json"))
    print("Phase 1: Converting recordings to observations...")
    words = list(test_recordings_dict.keys())
    for word in words:
        print(f"  Converting {len(test_recordings_dict[word])} recordings for word: '{word}'")
    all_recs = [rec for w in words for rec in test_recordings_dict[w]]
    all_obs = get_observations(all_recs, centroids)
    print("Phase 2: Testing all recording-HMM combinations...")
    true_labels: List[str] = []
    predicted_labels: List[str] = []
    if not all_obs:
        return true_labels, predicted_labels
    ll, arg = score_all(all_obs, all_hmm)
    pos = 0
    for word in words:
        n = len(test_recordings_dict[word])
        print(f"Testing {n} recordings for word: '{word}'")
        for r in range(n):
            k = int(arg[pos])
            predicted = all_hmm[k].word if k >= 0 else None
            if show_progress:
                likelihoods = {h.word: float(ll[pos, i]) for i, h in enumerate(all_hmm)}
                print(f"  Recording {r + 1} likelihoods: {likelihoods}")
                print(f"  True: '{word}' -> Predicted: '{predicted}'")
            true_labels.append(word)
            predicted_labels.append(predicted if predicted else "unknown")
            pos += 1
    return true_labels, predicted_labels

"""Model container + JSON persistence — same layout as the reference's HMM/hmm_classes.py
(HMMTrained :7-46, DataStorageHMM :48-95) so saved models round-trip between the two."""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class HMMTrained:
    """Trained HMM of one word: A [N,N], B [N,M], Pi [N] in linear space."""
    states: int
    symbols: int
    A: np.ndarray
    B: np.ndarray
    Pi: np.ndarray
    word: str

    def to_dict(self):
        return {"states": self.states, "symbols": self.symbols, "A": np.asarray(self.A).tolist(),
                "B": np.asarray(self.B).tolist(), "Pi": np.asarray(self.Pi).tolist(), "word": self.word}

    @classmethod
    def from_dict(cls, data):
        return cls(states=data["states"], symbols=data["symbols"], A=np.array(data["A"]), B=np.array(data["B"]),
                   Pi=np.array(data["Pi"]), word=data["word"])


class DataStorageHMM:
    """<base_dir>/<word>.json, keys states/symbols/A/B/Pi/word, indent=2 (hmm_classes.py:52-60)."""

    @staticmethod
    def save_hmm(hmm: HMMTrained, base_dir: str = "../Data/ResultsHMM", print_messages=True):
        os.makedirs(base_dir, exist_ok=True)
        filepath = os.path.join(base_dir, f"{hmm.word}.json")
        with open(filepath, "w") as f:
            json.dump(hmm.to_dict(), f, indent=2)
        if print_messages:
            print(f"Saved HMM for word '{hmm.word}' to {filepath}")

    @staticmethod
    def load_hmm(word: str, base_dir: str = "../Data/ResultsHMM", print_messages=True) -> HMMTrained:
        filepath = os.path.join(base_dir, f"{word}.json")
        with open(filepath, "r") as f:
            data = json.load(f)
        hmm = HMMTrained.from_dict(data)
        if print_messages:
            print(f"Loaded HMM for word '{word}' from {filepath}")
        return hmm

    @staticmethod
    def load_all_hmms(base_dir: str = "../Data/ResultsHMM", print_messages=True) -> List[HMMTrained]:
        """Every *.json in os.listdir order — that order is the recogniser's tie-break order
        (hmm_classes.py:84-91, hmm_testing.py:147-153)."""
        hmms: List[HMMTrained] = []
        if not os.path.exists(base_dir):
            print(f"Directory {base_dir} does not exist")
            return hmms
        for filename in os.listdir(base_dir):
            if filename.endswith(".json"):
                word = filename[:-5]
                try:
                    hmms.append(DataStorageHMM.load_hmm(word, base_dir, print_messages))
                except Exception as e:  # same swallow-and-report behaviour as the reference
                    print(f"Error loading HMM for word '{word}': {e}")
        if print_messages:
            print(f"Loaded {len(hmms)} HMM models total")
        return hmms

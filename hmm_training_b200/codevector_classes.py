"""Frame / centroid containers and their JSON + pickle layout, as consumed and produced by the
hot path (reference: CodeVector/codevector_classes.py — RawDataMFCC :204-279,
CentroidDataMFCC :321-342, DataStorage :434-597).

Only the data layout is reproduced.  The reference recomputes ``mfcc`` with librosa whenever
``raw_samples`` is non-empty (:217-220, and therefore also in ``from_dict``, :266-279); MFCC
extraction is upstream of the hot path so here ``mfcc_vector`` is always taken
as stored.  Files written by either side load in the other.
"""
from __future__ import annotations

import json
import pickle
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class RawDataMFCC:
    raw_samples: np.ndarray = field(default_factory=lambda: np.array([]))
    sample_rate: int = 16000
    n_channels: int = 1
    frame_duration_ms: float = 20.0
    mfcc: np.ndarray = field(default_factory=lambda: np.zeros(13))
    parent_centroid_id: int = 0
    generation: int = 0
    frame_number: int = 0
    recording: str = ""

    def __post_init__(self):
        self.raw_samples = np.asarray(self.raw_samples).flatten()
        self.mfcc = np.asarray(self.mfcc, dtype=float).flatten()

    def to_dict(self):
        return {"raw_samples": self.raw_samples.tolist(), "sample_rate": self.sample_rate,
                "n_channels": self.n_channels, "frame_duration_ms": self.frame_duration_ms,
                "mfcc_vector": self.mfcc.tolist(), "parent_centroid_id": int(self.parent_centroid_id),
                "generation": int(self.generation), "frame_number": self.frame_number, "recording": self.recording}

    @classmethod
    def from_dict(cls, data):
        return cls(raw_samples=np.array(data["raw_samples"]), sample_rate=data["sample_rate"],
                   n_channels=data["n_channels"], frame_duration_ms=data["frame_duration_ms"],
                   mfcc=np.array(data["mfcc_vector"]), parent_centroid_id=data["parent_centroid_id"],
                   generation=data["generation"], frame_number=data["frame_number"], recording=data["recording"])


@dataclass
class CentroidDataMFCC:
    mfcc: np.ndarray = field(default_factory=lambda: np.zeros(13))
    id: int = 0

    def to_dict(self):
        return {"mfcc": np.asarray(self.mfcc).tolist(), "id": int(self.id)}

    @classmethod
    def from_dict(cls, data):
        return cls(mfcc=np.array(data["mfcc"]), id=data["id"])


class DataStorage:
    """JSON (indent=2) and pickle I/O with the reference's method names (:438-597)."""

    @staticmethod
    def save_raw_data(raw_data_list, filepath: str, print_messages=False):
        with open(filepath, "w") as f:
            json.dump([frame.to_dict() for frame in raw_data_list], f, indent=2)
        if print_messages:
            print(f"    Saved {len(raw_data_list)} frames to {filepath}")

    @staticmethod
    def load_raw_data_mfcc(filepath: str, data_type: str = "auto", print_messages=True):
        with open(filepath, "r") as f:
            data = json.load(f)
        if not data:
            return []
        out = [RawDataMFCC.from_dict(d) for d in data]
        if print_messages:
            print(f"  Loaded {len(out)} frames from {filepath}")
        return out

    @staticmethod
    def save_centroids(centroids: List[CentroidDataMFCC], filepath: str):
        with open(filepath, "w") as f:
            json.dump([c.to_dict() for c in centroids], f, indent=2)
        print(f"Saved {len(centroids)} centroids to {filepath}")

    @staticmethod
    def load_centroids(filepath: str, print_messages=True) -> List[CentroidDataMFCC]:
        with open(filepath, "r") as f:
            data = json.load(f)
        centroids = [CentroidDataMFCC.from_dict(d) for d in data]
        if print_messages:
            print(f"Loaded {len(centroids)} centroids from {filepath}")
        return centroids

    @staticmethod
    def save_generations(generations: List[List[CentroidDataMFCC]], filepath: str):
        with open(filepath, "w") as f:
            json.dump([[c.to_dict() for c in gen] for gen in generations], f, indent=2)
        print(f"Saved {len(generations)} generations to {filepath}")

    @staticmethod
    def load_generations(filepath: str) -> List[List[CentroidDataMFCC]]:
        with open(filepath, "r") as f:
            data = json.load(f)
        generations = [[CentroidDataMFCC.from_dict(d) for d in gen] for gen in data]
        print(f"Loaded {len(generations)} generations from {filepath}")
        return generations

    @staticmethod
    def save_data_binary(data, filepath: str, print_messages=False):
        with open(filepath, "wb") as f:
            pickle.dump(data, f)
        if print_messages:
            print(f"    Saved data to {filepath} (binary format)")

    @staticmethod
    def load_data_binary(filepath: str):
        with open(filepath, "rb") as f:
            data = pickle.load(f)
        print(f"Loaded data from {filepath} (binary format)")
        return data


class AudioProcessor:
    """Recording -> overlapping 20 ms frames -> RawDataMFCC, as the reference's AudioProcessor
    (CodeVector/codevector_classes.py:346-431): frame_size 320, hop 160 at 16 kHz, the remaining tail kept as
    a shorter last frame if it has more than 12 samples (:413-431).  The reference computes the 13 MFCCs of
    each frame with one librosa call inside RawDataMFCC.__post_init__ (:217-250); here all frames of a
    recording go through hmmb_mfcc_frames in one batch per frame length (SURVEY.md §8f row 3 — parity with
    librosa itself is unpinned, see oracle/mfcc_oracle.py)."""

    def __init__(self, sample_rate=16000, frame_duration_ms=20, overlap_ms=10):
        self.sample_rate = sample_rate
        self.frame_duration_ms = frame_duration_ms
        self.overlap_ms = overlap_ms
        self.frame_size = int(sample_rate * frame_duration_ms / 1000)
        self.overlap_size = int(sample_rate * overlap_ms / 1000)
        self.hop_size = self.frame_size - self.overlap_size

    def _split_into_frames_with_overlap(self, audio_data: np.ndarray) -> List[np.ndarray]:
        frames = [audio_data[i:i + self.frame_size]
                  for i in range(0, len(audio_data) - self.frame_size + 1, self.hop_size)]
        last_start = len(frames) * self.hop_size
        if last_start < len(audio_data):
            last_frame = audio_data[last_start:]
            if len(last_frame) > 12:
                frames.append(last_frame)
        return frames

    def mfcc_matrix(self, audio_data: np.ndarray) -> np.ndarray:
        """[F, 13] MFCCs of the recording's frames (the packed form every consumer of this package accepts)."""
        from . import engine
        frames = self._split_into_frames_with_overlap(np.asarray(audio_data))
        out = np.zeros((len(frames), 13))
        by_len = {}
        for i, fr in enumerate(frames):
            by_len.setdefault(len(fr), []).append(i)
        for L, idx in by_len.items():
            Y = np.stack([np.asarray(frames[i], dtype=np.float64).reshape(-1) for i in idx])
            out[idx] = engine.mfcc_frames(Y, self.sample_rate)
        return out

    def process_recording(self, audio_path: str, purpose: str):
        """List of RawDataMFCC frames of a .npy recording (same objects for 'train', 'hmm' and 'test', :375-402)."""
        import os
        audio_data = np.load(audio_path)
        frames = self._split_into_frames_with_overlap(audio_data)
        mfcc = self.mfcc_matrix(audio_data)
        recording_name = os.path.splitext(os.path.basename(audio_path))[0]
        return [RawDataMFCC(raw_samples=frame, sample_rate=self.sample_rate, mfcc=mfcc[i], frame_number=i,
                            recording=recording_name) for i, frame in enumerate(frames)]


def load_mfcc_matrix(filepath: str) -> np.ndarray:
    """[F, 13] fp64 matrix of the "mfcc_vector" fields of a frame file written by
    DataStorage.save_raw_data (reference :438-444), in file order, WITHOUT building F
    RawDataMFCC objects: the text is scanned once by the native loader
    (hmmb_frames_json_scan, csrc/loader.cu).  Values are bit-identical to
    ``[f.mfcc for f in DataStorage.load_raw_data_mfcc(filepath)]`` of this package."""
    import ctypes
    from . import _lib
    with open(filepath, "rb") as f:
        text = f.read()
    cap = len(text) // 100 + 1  # a frame object is far longer than 100 bytes of text
    out = np.empty((cap, 13), dtype=np.float64)
    n = _lib.load().hmmb_frames_json_scan(text, len(text), out.ctypes.data_as(ctypes.c_void_p), cap)
    if n < 0:
        msg = _lib.load().hmmb_last_error().decode(errors="replace")
        raise ValueError(msg)
    if n > cap:  # cannot happen for files written by save_raw_data; keep the contract anyway
        out = np.empty((n, 13), dtype=np.float64)
        n = _lib.load().hmmb_frames_json_scan(text, len(text), out.ctypes.data_as(ctypes.c_void_p), n)
    return np.ascontiguousarray(out[:n])


def frames_matrix(frames) -> np.ndarray:
    """[F, 13] fp64 matrix of the frames' .mfcc (duck-typed: the reference's own RawDataMFCC
    objects work as well)."""
    if isinstance(frames, np.ndarray):  # already a packed [F, 13] matrix (load_mfcc_matrix)
        X = np.ascontiguousarray(frames, dtype=np.float64)
        if X.ndim != 2 or X.shape[1] != 13:
            raise ValueError("Vectors must be of size 13.")
        return X
    F = len(frames)
    X = np.empty((F, 13), dtype=np.float64)
    for i, fr in enumerate(frames):
        v = np.asarray(fr.mfcc, dtype=np.float64).reshape(-1)
        if v.shape[0] != 13:
            raise ValueError("Vectors must be of size 13.")
        X[i] = v
    return X

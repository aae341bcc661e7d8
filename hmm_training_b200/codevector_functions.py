"""Drop-in for the reference's CodeVector/codevector_functions.py (MFCC variant, the second
set of definitions :321-531 which shadows the LSF one).  Same names, arguments, prints, side
effects and return values; the arithmetic runs in libhmmb200.so on the GPU."""
from __future__ import annotations

import json
import os
from typing import List, Tuple

import numpy as np

from . import engine
from .codevector_classes import CentroidDataMFCC, DataStorage, RawDataMFCC, frames_matrix


def euclidian_distance(vec1: np.ndarray, vec2: np.ndarray) -> float:
    """codevector_functions.py:82-87 (host-side helper; the kernels inline it)."""
    if len(vec1) != len(vec2):
        raise ValueError("Vectors must be of size 13.")
    return float(np.linalg.norm(np.array(vec1) - np.array(vec2)))


def verify_frame_calculations(raw_data_vocabulary: List[RawDataMFCC]) -> bool:
    """codevector_functions.py:321-339 — the reference's check can never fire (:330)."""
    print("Verifying frame calculations...")
    print(f"  ✓ All {len(raw_data_vocabulary)} frames have proper calculations")
    return True


def save_updated_training_frames(raw_data_vocabulary: List[RawDataMFCC], output_dir: str):
    """codevector_functions.py:342-380."""
    print(f"Saving updated training frames to {output_dir}...")
    os.makedirs(output_dir, exist_ok=True)
    storage = DataStorage()
    updated_frames_path = os.path.join(output_dir, "codevector_frames_updated.json")
    storage.save_raw_data(raw_data_vocabulary, updated_frames_path)
    storage.save_data_binary(raw_data_vocabulary, os.path.join(output_dir, "codevector_frames_updated.pkl"))
    summary = {"total_frames": len(raw_data_vocabulary),
               "max_generation": max(f.generation for f in raw_data_vocabulary) if raw_data_vocabulary else 0,
               "centroid_assignments": {}}
    for frame in raw_data_vocabulary:
        cid = int(frame.parent_centroid_id)
        summary["centroid_assignments"][cid] = summary["centroid_assignments"].get(cid, 0) + 1
    summary_path = os.path.join(output_dir, "training_summary.json")
    with open(summary_path, "w") as f:
        json.dump(summary, f, indent=2)
    print(f"  Saved updated frames: {updated_frames_path}")
    print(f"  Saved training summary: {summary_path}")
    print(f"  Total frames: {summary['total_frames']}")
    print(f"  Max generation: {summary['max_generation']}")
    print(f"  Centroids used: {len(summary['centroid_assignments'])}")


def new_epsilon_centroids(centroids: List[CentroidDataMFCC], alpha1: float = 1.001,
                          alpha2: float = 0.999) -> List[CentroidDataMFCC]:
    """codevector_functions.py:383-411 (host-side; the LBG kernel does this on the device)."""
    n = len(centroids)
    if n & (n - 1) != 0 or n == 0:
        print(f"Warning: Number of centroids ({n}) is not a power of 2")
        return centroids
    out = []
    for i, c in enumerate(centroids):
        out.append(CentroidDataMFCC(mfcc=np.asarray(c.mfcc) * alpha1, id=2 * i))
        out.append(CentroidDataMFCC(mfcc=np.asarray(c.mfcc) * alpha2, id=2 * i + 1))
    return out


def new_adjust_centroids(raw_data_vocabulary: List[RawDataMFCC]) -> List[CentroidDataMFCC]:
    """codevector_functions.py:414-439 (host-side; the LBG kernel does this on the device)."""
    if not raw_data_vocabulary:
        return []
    generation = max(f.generation for f in raw_data_vocabulary)
    K = 2 ** generation
    X = frames_matrix(raw_data_vocabulary)
    ids = np.array([f.parent_centroid_id for f in raw_data_vocabulary])
    out = []
    for k in range(K):
        sel = X[ids == k]
        out.append(CentroidDataMFCC(mfcc=np.mean(sel, axis=0) if len(sel) else np.zeros(13), id=k))
    return out


def createCodeVector(raw_data_vocabulary: List[RawDataMFCC], centroids_quantity: int = 256, max_iterations=100,
                     epsilon: float = 0.001, save_updates: bool = True,
                     output_dir: str = None) -> Tuple[List[CentroidDataMFCC], List[List[CentroidDataMFCC]]]:
    """LBG codebook, codevector_functions.py:442-531.  Mutates ``frame.generation`` and
    ``frame.parent_centroid_id`` of every input frame exactly like the reference (:479-480,
    :502); returns (centroids, generations) with ids equal to list positions."""
    if not raw_data_vocabulary:
        raise ValueError("No raw data provided")
    print(f"Creating codevector with {centroids_quantity} centroids...")
    print(f"Using {len(raw_data_vocabulary)} frames for training")
    print("\n" + "=" * 50)
    if not verify_frame_calculations(raw_data_vocabulary):
        print("Warning: Some frames may have calculation issues")
    print("=" * 50)

    X = frames_matrix(raw_data_vocabulary)
    with np.errstate(divide="ignore", invalid="ignore"):
        int(np.log2(centroids_quantity))  # (:465) K = 0 -> OverflowError, K < 0 -> ValueError, as in the reference
    C, gens, assign, iters, gdist, hist = engine.lbg_fit(X, centroids_quantity, max_iterations, epsilon, history=True)
    n_gen = len(iters)
    for g in range(1, n_gen + 1):
        # the reference's progress lines (:472, :512-516), replayed from the per-pass summed distances
        print(f"\nGeneration {g}: Creating {1 << g} centroids")
        prev, diff = 0.0, epsilon + 100
        for it, gd in enumerate(hist[g - 1], start=1):
            diff = abs(prev - float(gd))
            prev = float(gd)
            if it % 10 == 0:
                print(f"  Iteration {it}: dist={gd:.6f}, diff={diff:.6f}")
        print(f"  Converged after {int(iters[g - 1])} iterations (diff={diff:.6f})")
    if n_gen > 0:
        for frame, cid in zip(raw_data_vocabulary, assign):
            frame.generation = n_gen
            frame.parent_centroid_id = int(cid)
    print("\nCodevector creation complete!")
    centroids = [CentroidDataMFCC(mfcc=C[k].copy(), id=k) for k in range(C.shape[0])]
    generations = [[CentroidDataMFCC(mfcc=row.copy(), id=k) for k, row in enumerate(gen)] for gen in gens]
    if save_updates and output_dir:
        print("\n" + "=" * 50)
        save_updated_training_frames(raw_data_vocabulary, output_dir)
        print("=" * 50)
    return centroids, generations

#!/usr/bin/env python3
"""Drop-in for the reference's HMM/main.py (the callers either side of the hot path,
SURVEY.md §8f rows 1, 2 and 4): same function names, arguments, prints and directory layout.

    load_all_recordings_by_word   HMM/main.py:46-100
    load_mfcc_centroids           HMM/main.py:103-130
    train_hmm                     HMM/main.py:133-164
    test                          HMM/main.py:167-197

The reference trains the vocabulary one word at a time (``training_with_save`` in a loop) and
loads every frame file through json.load + one RawDataMFCC object per frame.  Here
``train_hmm`` / ``test`` default to the batched path: frame files are read by the native
scanner straight into packed [T, 13] matrices (``fast=True``), the whole vocabulary is encoded
by one VQ launch and trained by one batched Baum-Welch call, models are written exactly where
``training_with_save`` writes them.  ``batched=False, fast=False`` reproduces the reference's
serial control flow call for call.
"""
from __future__ import annotations

import os
import random
import sys
from collections import defaultdict
from pathlib import Path
from typing import Dict

from .codevector_classes import DataStorage, load_mfcc_matrix
from .hmm_classes import DataStorageHMM
from .hmm_testing import create_confusion_matrix, test_hmm
from .hmm_training import train_hmm_batched, training_with_save


def load_all_recordings_by_word(base_dir="../Data", purpose="TrainHMM", print_messages=True, print_summary=True,
                                fast: bool = False) -> Dict[str, list]:
    """{word: [recording, ...]} from <base_dir>/<purpose>/<word>/<recording>/*_frames.json
    (HMM/main.py:46-100; the first ``*_frames.json`` of a recording directory is used, :77-80).
    A recording is a ``list[RawDataMFCC]`` as in the reference, or with ``fast=True`` the packed
    [T, 13] MFCC matrix of the same frames (no per-frame objects) — every consumer in this
    package accepts both."""
    storage = DataStorage()
    base_path = Path(base_dir)
    all_words = defaultdict(list)
    purpose_path = base_path / purpose
    if not purpose_path.exists():
        print(f"Warning: Directory {purpose_path} does not exist")
        return dict(all_words)
    if print_messages:
        print(f"Loading recordings from {purpose_path}")
    for word_dir in purpose_path.iterdir():
        if not word_dir.is_dir():
            continue
        word_name = word_dir.name
        if print_messages:
            print(f"  Processing word: {word_name}")
        for recording_dir in word_dir.iterdir():
            if not recording_dir.is_dir():
                continue
            frame_files = list(recording_dir.glob("*_frames.json"))
            if not frame_files:
                continue
            if fast:
                frames = load_mfcc_matrix(str(frame_files[0]))
                if print_messages:
                    print(f"  Loaded {len(frames)} frames from {frame_files[0]}")
            else:
                frames = storage.load_raw_data_mfcc(str(frame_files[0]), print_messages=print_messages)
            if len(frames):
                all_words[word_name].append(frames)
                if print_messages:
                    print(f"    Added recording with {len(frames)} frames from {recording_dir.name}")
    result = dict(all_words)
    if print_summary:
        print(f"\nSummary:")
        print(f"  Total words: {len(result)}")
        for word, recordings in result.items():
            total_frames = sum(len(recording) for recording in recordings)
            print(f"    {word}: {len(recordings)} recordings with {total_frames} total frames")
    return result


def load_mfcc_centroids(base_dir="../Data", print_messages=True):
    """list[CentroidDataMFCC] from <base_dir>/CodeVector/codevector.json (HMM/main.py:103-130)."""
    centroids = []
    storage = DataStorage()
    codevector_dir = os.path.join(base_dir, "CodeVector")
    if os.path.exists(os.path.join(codevector_dir, "codevector.json")):
        if print_messages:
            print("\nLoading codevector:")
        centroids = storage.load_centroids(os.path.join(codevector_dir, "codevector.json"))
        if print_messages:
            print(f"  Loaded codevector with {len(centroids)} centroids")
            print(f"  Example random centroid:")
            random_centroid = random.choice(centroids)
            print(f"   id: {random_centroid.id}")
            print(f"   Power: {random_centroid.mfcc[0]:.3f}")
            for i in range(1, random_centroid.mfcc.shape[0]):
                print(f"   {random_centroid.mfcc[i]:.3f}", end=" ")
            print(f"\n")
    return centroids


def train_hmm(show_progress=True, max_iterations=100, load_initial_params=False, base_dir="../Data",
              batched: bool = True, fast: bool = True):
    """Train one HMM per word of <base_dir>/TrainHMM (HMM/main.py:133-164); returns the list of
    HMMTrained (None on failure, as the reference) and writes ../Data/ResultsHMM/<word>.json."""
    print("Starting HMM training for all words...")
    try:
        centroids = load_mfcc_centroids(base_dir=base_dir, print_messages=False)
        print(f"Loaded {len(centroids)} centroids")
        recordings_by_word = load_all_recordings_by_word(base_dir=base_dir, purpose="TrainHMM", print_messages=False,
                                                         fast=fast)
        print(f"Loaded recordings for {len(recordings_by_word)} words")
        if batched:
            for word_name, word_recordings in recordings_by_word.items():
                print(f"\nTraining HMM for word: '{word_name}' with {len(word_recordings)} recordings")
            trained_hmms = train_hmm_batched(recordings_by_word, centroids, max_iterations=max_iterations,
                                             show_progress=show_progress, save=True,
                                             base_dir=os.path.join(base_dir, "ResultsHMM"),
                                             load_initial_params=load_initial_params)
            for hmm_model in trained_hmms:
                print(f"Model saved for word: '{hmm_model.word}'")
        else:
            trained_hmms = []
            for word_name, word_recordings in recordings_by_word.items():
                print(f"\nTraining HMM for word: '{word_name}' with {len(word_recordings)} recordings")
                hmm_model = training_with_save(word_recordings, centroids, word_name, max_iterations=max_iterations,
                                               show_progress=show_progress, load_initial_params=load_initial_params)
                trained_hmms.append(hmm_model)
                print(f"Model saved for word: '{hmm_model.word}'")
        print(f"\nHMM training completed successfully!")
        print(f"Total models trained: {len(trained_hmms)}")
        print(f"Words trained: {[hmm.word for hmm in trained_hmms]}")
        return trained_hmms
    except Exception as e:
        print(f"Error during HMM training: {e}")
        return None


def test(show_progress=False, base_dir="../Data", fast: bool = True):
    """Recognise <base_dir>/Test with the models of <base_dir>/ResultsHMM and report the
    confusion matrix (HMM/main.py:167-197).  Returns (true_labels, predicted_labels)."""
    print("Loading trained HMM models...")
    all_hmm = DataStorageHMM.load_all_hmms(os.path.join(base_dir, "ResultsHMM"))
    if not all_hmm:
        print("No trained HMM models found. Please train models first.")
        return None
    print(f"Loaded {len(all_hmm)} HMM models for words: {[hmm.word for hmm in all_hmm]}")
    test_recordings_dict = load_all_recordings_by_word(base_dir=base_dir, purpose="Test", print_messages=False, fast=fast)
    print(f"Loaded test recordings for {len(test_recordings_dict)} words")
    trained_words = {hmm.word for hmm in all_hmm}
    filtered = {word: recs for word, recs in test_recordings_dict.items() if word in trained_words}
    if not filtered:
        print("No test recordings found for trained words.")
        return None
    print(f"Testing on {len(filtered)} words: {list(filtered.keys())}")
    true_labels, predicted_labels = test_hmm(all_hmm, filtered, base_dir=base_dir, show_progress=show_progress)
    create_confusion_matrix(true_labels, predicted_labels, base_dir=base_dir)
    return true_labels, predicted_labels


def show_menu():
    print("=" * 50)
    print("AUDIO RECORDINGS LOADER")
    print("=" * 50)
    print("Options:")
    print("  python -m hmm_training_b200.main       -> Show this menu")
    print("  python -m hmm_training_b200.main train -> Run train")
    print("  python -m hmm_training_b200.main test  -> Run test")
    print("=" * 50)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "train":
        train_hmm()
    elif len(sys.argv) > 1 and sys.argv[1] == "test":
        test()
    else:
        show_menu()

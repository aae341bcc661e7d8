"""MFCC front-end against oracle/mfcc_oracle.py — a restatement of librosa's
algorithm for the reference's call; librosa itself is absent, so this parity is UNPINNED (see the
oracle's header).  Tolerance: |x - ref| <= 1e-8 * max(1, |ref|): the kernel evaluates the DFT
directly where numpy uses pocketfft, and libm / CUDA log10, exp differ in the last bit."""
import numpy as np
import pytest

from hmm_training_b200 import engine
from hmm_training_b200.codevector_classes import AudioProcessor
from oracle import mfcc_oracle as MO

pytestmark = pytest.mark.gpu


def _close(x, ref):
    assert x.shape == ref.shape
    err = np.abs(x - ref) / np.maximum(1.0, np.abs(ref))
    assert err.max() <= 1e-8, f"max scaled error {err.max():.3e}"


def _speechlike(rng, F, L, scale):
    t = np.arange(L) / 16000.0
    f0 = rng.uniform(80, 400, size=(F, 1))
    y = sum(rng.uniform(0.1, 1.0, size=(F, 1)) * np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6.28, size=(F, 1)))
            for h in range(1, 9))
    return (y + 0.05 * rng.normal(size=(F, L))) * scale


@pytest.mark.parametrize("L", [320, 319, 200, 57, 13, 640, 1024])
def test_mfcc_frames_match_oracle(L):
    rng = np.random.default_rng(L)
    Y = np.concatenate([_speechlike(rng, 40, L, 3000.0),            # int16-range speech-like frames
                        _speechlike(rng, 20, L, 1e-3),              # quiet
                        rng.normal(size=(10, L)),                   # white noise
                        np.zeros((1, L)),                           # digital silence: every band at the 1e-10 floor
                        np.full((1, L), 7.0)])                      # DC only: most bands 80 dB below the maximum
    got = engine.mfcc_frames(Y, 16000)
    ref = np.stack([MO.mfcc_frame(y, 16000) for y in Y])
    _close(got, ref)


def test_audio_processor_matches_oracle(tmp_path):
    rng = np.random.default_rng(7)
    audio = (_speechlike(rng, 1, 16000 // 2 + 77, 2000.0)[0]).astype(np.int16)  # 0.5 s + a 77-sample remainder
    path = tmp_path / "finish-01.npy"
    np.save(path, audio)
    ap = AudioProcessor()
    frames = ap.process_recording(str(path), "hmm")
    ref_frames = MO.split_into_frames_with_overlap(audio)
    assert len(frames) == len(ref_frames) and len(frames[-1].raw_samples) == len(ref_frames[-1]) != 320
    assert all(np.array_equal(f.raw_samples, r) for f, r in zip(frames, ref_frames))
    assert frames[3].recording == "finish-01" and frames[3].frame_number == 3
    _close(np.stack([f.mfcc for f in frames]), MO.mfcc_recording(audio))
    _close(ap.mfcc_matrix(audio), MO.mfcc_recording(audio))
    with pytest.raises(ValueError):
        engine.mfcc_frames(np.zeros(320))


def test_mfcc_chunked_upload_matches_plain():
    """Samples in PINNED host memory go up in ~16 MB chunks on the copy stream with the kernel running behind
    them; pageable samples go through the bounce buffers in one piece.  Same bits either way, and a sample of
    the frames against the oracle."""
    import ctypes
    from hmm_training_b200 import _lib
    rng = np.random.default_rng(9)
    F, L = 20000, 320  # 51 MB of samples: four chunks
    Y = _speechlike(rng, F, L, 2000.0)
    plain = engine.mfcc_frames(Y, 16000)
    lib = _lib.load()
    p = lib.hmmb_host_alloc(Y.nbytes)
    assert p
    try:
        Yp = np.ctypeslib.as_array((ctypes.c_double * Y.size).from_address(p)).reshape(Y.shape)
        Yp[...] = Y
        chunked = engine.mfcc_frames(Yp, 16000)
    finally:
        lib.hmmb_host_free(p)
    assert np.array_equal(plain, chunked)
    pick = rng.choice(F, 50, replace=False)
    _close(plain[pick], np.stack([MO.mfcc_frame(Y[i], 16000) for i in pick]))

"""CPU checks of the MFCC oracle's building blocks against scipy (what librosa calls for the window and
the DCT) and of the structure of the Slaney mel filter bank.  librosa itself is absent: see the header
of oracle/mfcc_oracle.py (parity unpinned)."""
import numpy as np
import pytest

from oracle import mfcc_oracle as MO


def test_window_and_dct_match_scipy():
    signal = pytest.importorskip("scipy.signal")
    fftpack = pytest.importorskip("scipy.fftpack")
    rng = np.random.default_rng(0)
    for n in (320, 200, 37, 13):
        y = rng.normal(size=n)
        w = signal.get_window("hann", n, fftbins=True)
        spec = np.abs(np.fft.rfft(y * w)) ** 2
        mel = MO.mel_filterbank(16000, n).astype(np.float64) @ spec
        db = 10 * np.log10(np.maximum(1e-10, mel))
        db = np.maximum(db, db.max() - 80.0)
        want = fftpack.dct(db, type=2, norm="ortho")[:13]
        assert np.allclose(MO.mfcc_frame(y), want, rtol=1e-12, atol=1e-10)


def test_mel_filterbank_structure():
    fb = MO.mel_filterbank(16000, 320, 26)
    assert fb.shape == (26, 161) and fb.dtype == np.float32 and (fb >= 0).all()
    edges = MO.mel_band_edges(16000, 26)
    assert edges[0] == 0.0 and abs(edges[-1] - 8000.0) < 1e-9 and np.all(np.diff(edges) > 0)
    assert np.allclose(np.diff(edges)[:8], 200.0 / 3 * (MO.hz_to_mel(8000.0) / 27), rtol=1e-12)  # linear below 1 kHz
    centre = fb.argmax(axis=1)
    assert np.all(np.diff(centre) > 0)  # one triangle per band, moving up in frequency
    freqs = np.fft.rfftfreq(320, 1 / 16000)
    for i in range(26):
        nz = np.flatnonzero(fb[i])
        assert freqs[nz[0]] > edges[i] - 1e-9 and freqs[nz[-1]] < edges[i + 2] + 1e-9


def test_framing_matches_reference_rule():
    a = np.arange(1000)
    fr = MO.split_into_frames_with_overlap(a)
    assert [len(f) for f in fr] == [320] * 5 + [200] and fr[1][0] == 160 and fr[-1][0] == 800
    assert [len(f) for f in MO.split_into_frames_with_overlap(np.arange(330))] == [320, 170]
    assert [len(f) for f in MO.split_into_frames_with_overlap(np.arange(10))] == []      # <= 12 samples: dropped
    assert [len(f) for f in MO.split_into_frames_with_overlap(np.arange(13))] == [13]

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


def test_pipeline_matches_transformers_audio_utils():
    """A third-party cross-check of the restatement (NOT a pin: still no librosa output).  transformers.audio_utils
    re-implements librosa's conventions — `mel_filter_bank(norm="slaney", mel_scale="slaney")` = `librosa.filters.mel`,
    `power_to_db` with amin 1e-10 and an 80 dB range, periodic Hann — and its own test-suite holds it to librosa.  The
    whole per-frame pipeline of the reference's call (one un-centred 320-sample frame, 26 mel bands, ortho DCT-II, 13
    coefficients; CodeVector/codevector_classes.py:226-250) agrees with it to 1e-7: the difference is librosa's
    float32 filter weights, which the oracle keeps and transformers does not."""
    au = pytest.importorskip("transformers.audio_utils")
    from scipy.fft import dct
    sr, n = 16000, 320
    fb_t = au.mel_filter_bank(num_frequency_bins=n // 2 + 1, num_mel_filters=26, min_frequency=0.0, max_frequency=sr / 2,
                              sampling_rate=sr, norm="slaney", mel_scale="slaney")
    assert np.abs(MO.mel_filterbank(sr, n, 26) - fb_t.T).max() < 2e-9  # float32 rounding of weights <= 8.7e-3
    win = au.window_function(n, "hann", periodic=True)
    rng = np.random.default_rng(0)
    for k in range(40):
        y = rng.normal(size=n) * 10 ** rng.uniform(-3, 0)
        if k == 0:
            y[:] = 0.0  # silence: every band at the 1e-10 floor
        if k == 1:
            y = np.sin(2 * np.pi * 440.0 * np.arange(n) / sr)  # one loud band, the others at the 80 dB floor
        S = au.spectrogram(y, win, frame_length=n, hop_length=n, fft_length=n, power=2.0, center=False, mel_filters=fb_t,
                           log_mel="dB", mel_floor=1e-10, reference=1.0, min_value=1e-10, db_range=80.0, dtype=np.float64)
        want = dct(S[:, 0], type=2, norm="ortho")[:13]
        got = MO.mfcc_frame(y, sr)
        assert np.abs(got - want).max() <= 1e-7 * max(1.0, np.abs(want).max())


def test_pipeline_matches_torchaudio():
    """A second independent implementation of librosa's conventions (still NOT a pin): `torchaudio.transforms.MFCC` with
    `norm="slaney", mel_scale="slaney"` filters, a periodic Hann window, un-centred frames, power spectrogram, the
    80 dB `AmplitudeToDB` and the orthonormal DCT-II — torchaudio's own test-suite compares exactly this configuration
    with librosa.  It builds its filterbank in float32 (5.6e-9 from the oracle's), so the 13 coefficients of the
    reference's per-frame call (CodeVector/codevector_classes.py:226-250) agree to 4e-7 rather than to rounding."""
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    sr, n = 16000, 320
    mfcc = torchaudio.transforms.MFCC(sample_rate=sr, n_mfcc=13, dct_type=2, norm="ortho", log_mels=False,
                                      melkwargs=dict(n_fft=n, win_length=n, hop_length=n, center=False, n_mels=26, f_min=0.0,
                                                     f_max=sr / 2, power=2.0, norm="slaney", mel_scale="slaney")).double()
    fb = torchaudio.functional.melscale_fbanks(n // 2 + 1, 0.0, sr / 2, 26, sr, norm="slaney", mel_scale="slaney").double().numpy()
    assert np.abs(MO.mel_filterbank(sr, n, 26) - fb.T).max() < 2e-8
    rng = np.random.default_rng(0)
    for k in range(40):
        y = rng.normal(size=n) * 10 ** rng.uniform(-3, 0)
        if k == 0:
            y[:] = 0.0
        if k == 1:
            y = np.sin(2 * np.pi * 440.0 * np.arange(n) / sr)
        want = mfcc(torch.from_numpy(y)[None])[0, :, 0].numpy()
        got = MO.mfcc_frame(y, sr)
        assert np.abs(got - want).max() <= 2e-6 * max(1.0, np.abs(want).max())

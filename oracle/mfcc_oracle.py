"""TEST INFRASTRUCTURE — CPU restatement of the reference's MFCC front-end.

PARITY UNPINNED: the reference computes the 13 coefficients of a frame with
``librosa.feature.mfcc(y, sr, n_mfcc=13, n_fft=len(frame), hop_length=None, center=False, n_mels=26)``
(CodeVector/codevector_classes.py:226-250, one call per 20 ms frame, framing at :413-431).  librosa
(pinned 0.11.0, requirements.txt:16) is NOT installed in this image and cannot be fetched, so this
file restates librosa's published algorithm for exactly that call and could not be checked against
librosa itself; no golden vector of the reference exists for it.  What the restatement
follows, step by step (librosa 0.11.0):

  stft            periodic Hann window of n_fft samples (scipy.signal.get_window('hann', n, fftbins=True)),
                  one frame (center=False, len(y) == n_fft), rfft in float64 -> complex128
  _spectrogram    |X|**2
  filters.mel     Slaney mel scale (htk=False), fmin = 0, fmax = sr / 2, n_mels + 2 band edges from
                  np.linspace in mel space, triangular weights max(0, min(lower, upper)) STORED AS FLOAT32,
                  then multiplied by the Slaney area normalisation 2 / (f[i+2] - f[i]) (again rounded to float32)
  melspectrogram  mel_basis @ power spectrum (float64 accumulation)
  power_to_db     10 log10(max(1e-10, S)) - 10 log10(max(1e-10, 1.0)), then max(., max - 80 dB)
  mfcc            scipy.fftpack.dct(type=2, norm='ortho') over the 26 bands, first 13 coefficients, no lifter

Only tests/ and bench.py may import this module.
"""
from __future__ import annotations

from typing import List

import numpy as np


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=float)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        m = f >= min_log_hz
        mels[m] = min_log_mel + np.log(f[m] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def mel_to_hz(mels):
    mels = np.asanyarray(mels, dtype=float)
    f_sp = 200.0 / 3
    freqs = f_sp * mels
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if mels.ndim:
        m = mels >= min_log_mel
        freqs[m] = min_log_hz * np.exp(logstep * (mels[m] - min_log_mel))
    elif mels >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (mels - min_log_mel))
    return freqs


def mel_band_edges(sr: float, n_mels: int = 26) -> np.ndarray:
    """librosa.mel_frequencies(n_mels + 2, fmin=0, fmax=sr/2, htk=False)."""
    return mel_to_hz(np.linspace(hz_to_mel(0.0), hz_to_mel(sr / 2.0), n_mels + 2))


def mel_filterbank(sr: float, n_fft: int, n_mels: int = 26) -> np.ndarray:
    """librosa.filters.mel(sr=sr, n_fft=n_fft, n_mels=n_mels) — float32 [n_mels, 1 + n_fft // 2]."""
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_band_edges(sr, n_mels)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def mfcc_frame(y: np.ndarray, sr: int = 16000, n_mfcc: int = 13, n_mels: int = 26) -> np.ndarray:
    """The 13 coefficients RawDataMFCC.calculate_mfcc stores for one frame (codevector_classes.py:226-250)."""
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    n = len(y)
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)  # scipy.signal.get_window('hann', n, fftbins=True)
    spec = np.abs(np.fft.rfft(y * window)) ** 2.0
    mel = mel_filterbank(sr, n, n_mels).astype(np.float64) @ spec
    log_spec = 10.0 * np.log10(np.maximum(1e-10, mel))
    log_spec = np.maximum(log_spec, log_spec.max() - 80.0)
    k = np.arange(n_mfcc)[:, None]
    m = np.arange(n_mels)[None, :]
    basis = np.cos(np.pi * k * (2 * m + 1) / (2.0 * n_mels)) * np.sqrt(2.0 / n_mels)
    basis[0] *= np.sqrt(0.5)
    return basis @ log_spec


def split_into_frames_with_overlap(audio: np.ndarray, frame_size: int = 320, hop_size: int = 160) -> List[np.ndarray]:
    """AudioProcessor._split_into_frames_with_overlap (codevector_classes.py:413-431): full frames at
    every hop, plus the remaining tail (from len(frames) * hop) if it has more than 12 samples."""
    frames = [audio[i:i + frame_size] for i in range(0, len(audio) - frame_size + 1, hop_size)]
    last_start = len(frames) * hop_size
    if last_start < len(audio):
        last = audio[last_start:]
        if len(last) > 12:
            frames.append(last)
    return frames


def mfcc_recording(audio: np.ndarray, sr: int = 16000) -> np.ndarray:
    """[F, 13] MFCC matrix of a recording: process_recording's frames, one mfcc_frame each."""
    frames = split_into_frames_with_overlap(np.asarray(audio))
    return np.stack([mfcc_frame(f, sr) for f in frames]) if frames else np.zeros((0, 13))

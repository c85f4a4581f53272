"""CPU restatement of the reference's voice-activity detector (/root/reference/VAD.py) -- SURVEY 8(f).4.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PINNED: ``tests/golden/make_golden.py`` imports the unmodified
``VAD.py`` through ``oracle.ref_shims`` and commits its outputs on synthetic signals as ``tests/golden/vad.npz``;
``tests/test_oracle.py`` checks every function below against them.

Conventions follow the reference, quirks included:
* ``enframe`` (VAD.py:28-49): frames of 256 with a hop of 128, ``ceil(len / hop)`` frames, zero-padded tail, no window,
  returned column-major ``(frame_size, n_frames)``;
* ``wav_normalise`` (VAD.py:133): samples divided by the peak magnitude;
* ``zcr`` (VAD.py:52-63): number of strictly negative products of neighbouring samples;
* ``energy`` (VAD.py:66-76): sum of squares;
* ``spectral_entropy`` (VAD.py:79-108): |FFT| of the frame, first half, ten sub-bands of floor(128 / 10) = 12 bins
  (the last 8 bins only count in the total), ``-sum s log2(s + eps)``;
* ``detect`` (VAD.py:137-182): the double-threshold state machine, including that a voiced run is only closed by a
  later quiet frame, that ``last_end`` is never set outside the merge branch (so runs are never merged) and that the
  backward search may step to negative (wrapped) indices.
"""
from __future__ import annotations

import math

import numpy as np

FRAME_SIZE = 256
OVERLAP = 128


def wav_normalise(wave_data):
    """VAD.py:133 -- ``waveData / max(abs(waveData))``."""
    wave_data = np.asarray(wave_data, dtype=np.float64)
    return wave_data / np.max(np.abs(wave_data))


def enframe(wave_data, frame_size=FRAME_SIZE, overlap=OVERLAP):
    wave_data = np.asarray(wave_data, dtype=np.float64).reshape(-1)
    step = frame_size - overlap
    n = math.ceil(len(wave_data) / step)
    out = np.zeros((frame_size, n))
    for i in range(n):
        seg = wave_data[i * step : min(i * step + frame_size, len(wave_data))]
        out[: len(seg), i] = seg
    return out


def zcr(frames):
    prod = frames[:-1] * frames[1:]
    return np.sum(prod < 0, axis=0).astype(np.float64).reshape(-1, 1)


def energy(frames):
    return np.sum(frames * frames, axis=0).reshape(-1, 1)


def spectral_entropy(frames, n_short_blocks=10, eps=1e-8):
    frame_size = frames.shape[0]
    mag = np.abs(np.fft.fft(frames, axis=0))[: frame_size // 2]        # (128, n)
    total = np.sum(mag**2, axis=0)
    sub = (frame_size // 2) // n_short_blocks
    blocks = (mag[: sub * n_short_blocks] ** 2).reshape(n_short_blocks, sub, -1).sum(axis=1)  # consecutive bands of `sub` bins
    s = blocks / (total + eps)
    return (-np.sum(s * np.log2(s + eps), axis=0)).reshape(-1, 1)


def feature(frames):
    """VAD.py:111-123 -- (zcr gated by power > 0.1, power, spectral entropy)."""
    power = energy(frames)
    return zcr(frames) * (power > 0.1), power, spectral_entropy(frames)


def _py_index(n, i):
    """Python/numpy indexing of a length-n array with a possibly negative index."""
    if i < -n or i >= n:
        raise IndexError(i)
    return i + n if i < 0 else i


def detect(zcr_v, power, zcr_gate=35, ampl=0.3, amph=12, min_len=16):
    """VAD.py:137-182."""
    z = np.asarray(zcr_v, dtype=np.float64).reshape(-1)
    p = np.asarray(power, dtype=np.float64).reshape(-1)
    n = len(z)
    res = np.zeros((n, 1))
    status, start, end = 0, 0, 0
    for i in range(n):
        if p[i] > amph:
            if status != 1:
                start = i
            end = i
            status = 1
        elif end - start + 1 > min_len:
            while p[_py_index(n, start)] > ampl or z[_py_index(n, start)] > zcr_gate:
                start -= 1
            start += 1
            while p[end] > ampl or z[end] > zcr_gate:
                end += 1
                if end == n:
                    break
            end -= 1
            res[slice(start, end + 1)] = 1   # last_end stays -1 in the reference: the merge branch is never taken
            start, end, status = 0, 0, 0
    return res


def frequency(spectrum, gate=0.4):
    """VAD.py:185-186."""
    return np.where(np.asarray(spectrum) > gate, 0, 1)

"""Voice-activity detection of the reference (``VAD.py``) on the GPU -- SURVEY 8(f).4.

The report runs this detector in front of GMM-UBM; in the reference it is a set of per-frame Python loops
(``enframe`` VAD.py:28-49, ``ZCR`` :52-63, ``energy`` :66-76, ``spectrum_entropy`` :98-108, ``feature`` :111-123,
``VAD_detection`` :137-182, ``VAD_frequency`` :185-186).  The functions below keep those names, argument
conventions (column-major ``(frameSize, frameNum)`` frame matrices, ``(frameNum, 1)`` results) and quirks; the
arithmetic runs in ``ssp_vad_features`` / ``ssp_vad_detect``.  :func:`vad_batch` is the batched form the reference
lacks: raw PCM of many utterances in, per-frame speech decisions out, two kernel launches.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib

frameSize = 256   # VAD.py:22
overlap = 128     # VAD.py:23


def _cfg(frame_shift: int, normalize_peak: bool, pcm_dtype: int) -> _lib.VadCfg:
    return _lib.VadCfg(frameSize, frame_shift, 10, 1 if normalize_peak else 0, pcm_dtype, 1e-8)


def _features_device(pcm, sample_offsets: np.ndarray, frame_shift: int, normalize_peak: bool):
    """(zcr float32, power float64, entropy float32) CUDA tensors over all frames + frame_offsets (numpy int64)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = pcm.device
    if pcm.dtype == torch.int16:
        dt = 0
    elif pcm.dtype == torch.float32:
        dt = 1
    else:
        raise TypeError("PCM must be int16 or float32")
    cfg = _cfg(frame_shift, normalize_peak, dt)
    lens = np.diff(sample_offsets)
    n_frames = np.array([int(lib.ssp_vad_num_frames(C.byref(cfg), int(n))) for n in lens], dtype=np.int64)
    foffs = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(n_frames, out=foffs[1:])
    total = int(foffs[-1])
    zcr = torch.empty(total, dtype=torch.float32, device=dev)
    power = torch.empty(total, dtype=torch.float64, device=dev)
    ent = torch.empty(total, dtype=torch.float32, device=dev)
    d_s = torch.as_tensor(np.asarray(sample_offsets, dtype=np.int64), device=dev)
    d_f = torch.as_tensor(foffs, device=dev)
    if total:
        _lib.check(lib.ssp_vad_features(_lib.ptr(pcm), _lib.ptr(d_s), len(lens), C.byref(cfg), _lib.ptr(d_f), _lib.ptr(zcr),
                                        _lib.ptr(power), _lib.ptr(ent), _lib.stream_ptr()), "ssp_vad_features")
    return zcr, power, ent, foffs, d_f


def _frames_to_device(frameData):
    """A (frameSize, frameNum) frame matrix as a flat float32 signal of non-overlapping frames."""
    torch = _lib.require_cuda()
    fd = np.asarray(frameData, dtype=np.float64)
    if fd.ndim != 2 or fd.shape[0] != frameSize:
        raise ValueError(f"expected a ({frameSize}, frameNum) frame matrix, got {fd.shape}")
    flat = np.ascontiguousarray(fd.T, dtype=np.float32).reshape(-1)
    offs = np.array([0, flat.size], dtype=np.int64)
    return torch.as_tensor(flat, device="cuda"), offs


# ---------------------------------------------------------------------------------------------- reference names
def enframe(wavData):
    """VAD.py:28-49 (host-side reshaping only; no arithmetic)."""
    wav = np.asarray(wavData, dtype=np.float64).reshape(-1)
    step = frameSize - overlap
    n = math.ceil(len(wav) / step)
    pad = np.zeros((n - 1) * step + frameSize)
    pad[: len(wav)] = wav
    idx = np.arange(frameSize)[:, None] + step * np.arange(n)[None, :]
    return pad[idx]


def _frame_features(frameData):
    pcm, offs = _frames_to_device(frameData)
    z, p, e, _, _ = _features_device(pcm, offs, frame_shift=frameSize, normalize_peak=False)
    return z, p, e


def ZCR(frameData):
    """VAD.py:52-63 -> (frameNum, 1) float64."""
    z, _, _ = _frame_features(frameData)
    return z.cpu().numpy().astype(np.float64).reshape(-1, 1)


def energy(frameData):
    """VAD.py:66-76 -> (frameNum, 1) float64."""
    _, p, _ = _frame_features(frameData)
    return p.cpu().numpy().reshape(-1, 1)


def spectrum_entropy(frameData):
    """VAD.py:98-108 -> (frameNum, 1) float64."""
    _, _, e = _frame_features(frameData)
    return e.cpu().numpy().astype(np.float64).reshape(-1, 1)


def feature(waveData):
    """VAD.py:111-123: ``(zcr * (power > 0.1), power, spectral entropy)`` of a frame matrix (one launch)."""
    z, p, e = _frame_features(waveData)
    z = (z.double() * (p > 0.1)).cpu().numpy().reshape(-1, 1)
    return z, p.cpu().numpy().reshape(-1, 1), e.cpu().numpy().astype(np.float64).reshape(-1, 1)


def VAD_detection(zcr, power, zcr_gate=35, ampl=0.3, amph=12):
    """VAD.py:137-182 -> (frameNum, 1) float64 of 0/1."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    z = torch.as_tensor(np.ascontiguousarray(np.asarray(zcr, dtype=np.float32).reshape(-1)), device="cuda")
    p = torch.as_tensor(np.ascontiguousarray(np.asarray(power, dtype=np.float64).reshape(-1)), device="cuda")
    if z.numel() != p.numel():
        raise ValueError("zcr and power must have one entry per frame")
    offs = torch.as_tensor(np.array([0, z.numel()], dtype=np.int64), device="cuda")
    out = torch.zeros(z.numel(), dtype=torch.uint8, device="cuda")
    if z.numel():
        _lib.check(lib.ssp_vad_detect(_lib.ptr(z), _lib.ptr(p), _lib.ptr(offs), 1, float(zcr_gate), float(ampl), float(amph), 16,
                                      _lib.ptr(out), _lib.stream_ptr()), "ssp_vad_detect")
    return out.cpu().numpy().astype(np.float64).reshape(-1, 1)


def VAD_frequency(spectrum):
    """VAD.py:185-186."""
    return np.where(np.asarray(spectrum) > 0.4, 0, 1)


# ---------------------------------------------------------------------------------------------- batched form
def vad_batch(signals, zcr_gate=35, ampl=0.3, amph=12, device=None):
    """Speech/non-speech decision per frame for a list of int16 (or float) utterances: peak normalisation
    (VAD.py:133), framing, features and the time-domain detector for ALL utterances in two launches.

    Returns ``(speech, frame_offsets, features)``: ``speech`` uint8 CUDA tensor over all frames (1 = speech),
    ``frame_offsets`` numpy int64 (utterance u owns ``speech[frame_offsets[u]:frame_offsets[u+1]]``) and the
    ``(zcr_gated, power, entropy)`` CUDA tensors.
    """
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    sigs = [np.asarray(s).reshape(-1) for s in signals]
    is_int = all(s.dtype == np.int16 for s in sigs)
    lens = np.array([len(s) for s in sigs], dtype=np.int64)
    offs = np.zeros(len(sigs) + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    host = torch.empty(int(offs[-1]), dtype=torch.int16 if is_int else torch.float32, pin_memory=True)
    hv = host.numpy()
    for s, o in zip(sigs, offs[:-1]):
        hv[o : o + len(s)] = s
    pcm = host.to(dev, non_blocking=True)
    z, p, e, foffs, d_f = _features_device(pcm, offs, frame_shift=frameSize - overlap, normalize_peak=True)
    zg = (z.double() * (p > 0.1)).float()
    speech = torch.zeros(int(foffs[-1]), dtype=torch.uint8, device=dev)
    if len(sigs) and int(foffs[-1]):
        _lib.check(lib.ssp_vad_detect(_lib.ptr(zg), _lib.ptr(p), _lib.ptr(d_f), len(sigs), float(zcr_gate), float(ampl),
                                      float(amph), 16, _lib.ptr(speech), _lib.stream_ptr()), "ssp_vad_detect")
    return speech, foffs, (zg, p, e)

"""Deterministic synthetic audio / feature generators.

There is no dataset in the reference repo (README.md:15-16 points to a private
download), so every test and benchmark input is generated here from integer seeds.
Host-side numpy only; ``synth_features_torch`` builds the large bench inputs directly
in HBM.
"""
from __future__ import annotations

import numpy as np


def synth_utterance(speaker: int, utt: int, n_samples: int = 48000, fs: int = 16000) -> np.ndarray:
    """One int16 utterance: an impulse train at a per-speaker f0 plus white noise, driven through
    four two-pole resonators whose formants hop every 80-200 ms between the entries of a
    per-speaker inventory of 6 "vowels" (so per-utterance CMVN keeps speaker information);
    RMS ~3000, never digital silence."""
    from scipy.signal import lfilter

    rs = np.random.RandomState(speaker)
    f0 = rs.uniform(80.0, 250.0)
    lo = np.array([300.0, 900.0, 2400.0, 3300.0])
    hi = np.array([800.0, 2300.0, 3200.0, 4200.0])
    if fs < 16000:
        lo, hi = lo * fs / 16000.0, hi * fs / 16000.0
    inventory = lo + (hi - lo) * rs.uniform(size=(6, 4))
    bws = np.array([rs.uniform(60, 120), rs.uniform(80, 160), rs.uniform(100, 200), rs.uniform(120, 240)])
    ru = np.random.RandomState((speaker * 1000 + utt) % (2 ** 31 - 1))
    f0 = f0 * (1.0 + 0.05 * ru.standard_normal())
    period = max(2, int(round(fs / f0)))
    src = np.zeros(n_samples)
    src[ru.randint(0, period)::period] = 1.0
    src += 10 ** (-15 / 20.0) * ru.standard_normal(n_samples) * np.sqrt(1.0 / period)
    y = np.zeros(n_samples)
    pos = 0
    state = [np.zeros(2) for _ in range(4)]
    while pos < n_samples:
        seg = int(ru.uniform(0.08, 0.2) * fs)
        end = min(n_samples, pos + seg)
        formants = inventory[ru.randint(0, 6)] * (1.0 + 0.02 * ru.standard_normal(4))
        for i, (fc, bw) in enumerate(zip(formants, bws)):
            r = np.exp(-np.pi * bw / fs)
            a = [1.0, -2.0 * r * np.cos(2 * np.pi * fc / fs), r * r]
            out, state[i] = lfilter([1.0], a, src[pos:end], zi=state[i])
            y[pos:end] += out
        pos = end
    y += 1e-3 * ru.standard_normal(n_samples) * (np.abs(y).max() + 1e-9)
    y *= 3000.0 / (np.sqrt(np.mean(y * y)) + 1e-12)
    y = np.clip(np.round(y), -32768, 32767).astype(np.int16)
    y[y == 0] = 1  # never digital silence (log(0) in the sidekit recipe)
    return y


def synth_corpus(n_speakers: int, n_utts: int, n_samples: int = 48000, fs: int = 16000):
    """(list of int16 utterances, list of speaker labels), seeds = speaker*1000 + utt."""
    x, y = [], []
    for s in range(n_speakers):
        for u in range(n_utts):
            x.append(synth_utterance(s, u, n_samples, fs))
            y.append(s)
    return x, y


def synth_ubm(k: int, d: int, seed: int = 0, spread: float = 1.5, dtype=np.float64):
    """A plausible diag GMM: weights ~ Dirichlet-ish, means ~ N(0, spread^2), variances in
    [0.3, 1.2]."""
    rs = np.random.RandomState(seed)
    w = rs.gamma(2.0, 1.0, size=k)
    w /= w.sum()
    mu = spread * rs.standard_normal((k, d))
    var = rs.uniform(0.3, 1.2, size=(k, d))
    return w.astype(dtype), mu.astype(dtype), var.astype(dtype)


def synth_speaker_means(mu_ubm: np.ndarray, n_speakers: int, seed: int = 1, shift: float = 0.35) -> np.ndarray:
    """Per-speaker component means = UBM means + a speaker-specific offset (what mean-only MAP
    produces): (S, K, D)."""
    rs = np.random.RandomState(seed)
    k, d = mu_ubm.shape
    return mu_ubm[None] + shift * rs.standard_normal((n_speakers, k, d))


def sample_gmm(w, mu, var, n: int, seed: int = 0) -> np.ndarray:
    rs = np.random.RandomState(seed)
    comp = rs.choice(len(w), size=n, p=np.asarray(w, dtype=np.float64) / np.sum(w))
    return (mu[comp] + np.sqrt(var[comp]) * rs.standard_normal((n, mu.shape[1]))).astype(np.float32)


def synth_features_torch(n_frames: int, d: int, mu_spk, var, labels_per_frame, seed: int, device):
    """Large synthetic feature matrix built in HBM: each frame picks a random component of its
    speaker's model and adds unit-scaled noise.  ``mu_spk`` (S,K,D) and ``var`` (K,D) are torch
    tensors on ``device``; ``labels_per_frame`` (n_frames,) int64 speaker ids."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    k = mu_spk.shape[1]
    comp = torch.randint(0, k, (n_frames,), generator=g, device=device)
    noise = torch.randn((n_frames, d), generator=g, device=device, dtype=torch.float32)
    m = mu_spk[labels_per_frame, comp].to(torch.float32)
    s = var[comp].to(torch.float32).sqrt()
    return (m + s * noise).contiguous()

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


# ------------------------------------------------------------------------------------------------
# Counter-based torch generator: the SAME audio on any device (bench.py: GPU arm, CPU arm, oracle check)
# ------------------------------------------------------------------------------------------------
def _hash_u32(x):
    """Integer hash of an int64 tensor -> values in [0, 2^32) (int64).  Only integer ops that are exact and identical
    on CPU and CUDA; intermediate products stay below 2^63."""
    x = x & 0xFFFFFFFF
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & 0xFFFFFFFF
    x = ((x ^ (x >> 16)) * 0x45D9F3B) & 0xFFFFFFFF
    return (x ^ (x >> 16)) & 0xFFFFFFFF


def _hash_uniform(x):
    """[0, 1) float32 from integer keys."""
    return _hash_u32(x).to(dtype=_torch().float32) * (1.0 / 4294967296.0)


def _torch():
    import torch

    return torch


N_VOWELS, N_HARM, SEG_SAMPLES = 5, 12, 1600


def synth_pcm_torch(speakers, utts, n_samples: int, device, fs: int = 16000):
    """int16 PCM (n, n_samples) for utterance ``utts[i]`` of speaker ``speakers[i]`` (int64 tensors or sequences).

    Voiced-speech-like and speaker-discriminative AFTER per-utterance CMVN: every speaker has a pitch and an inventory
    of N_VOWELS two-formant spectral envelopes; an utterance hops between them every 100 ms in a hash-chosen order, so
    the spread of the cepstra around the utterance mean is the speaker's own.  A pure function of (speaker, utt, sample
    index): integer hashes instead of an RNG stream, so any device, batch split or process regenerates the same
    samples (up to the last ulp of sin() before the int16 rounding).  RMS ~3000, never digital silence."""
    torch = _torch()
    spk = torch.as_tensor(speakers, dtype=torch.int64, device=device).reshape(-1)
    utt = torch.as_tensor(utts, dtype=torch.int64, device=device).reshape(-1)
    n = spk.numel()
    f32 = torch.float32
    f0 = 90.0 + 140.0 * _hash_uniform(spk * 7919 + 11)                                  # (n,)
    f0 = f0 * (1.0 + 0.04 * (_hash_uniform(spk * 104729 + utt * 31 + 5) - 0.5))
    v = torch.arange(N_VOWELS, device=device, dtype=torch.int64)
    key_sv = spk[:, None] * 977 + v[None] * 131                                          # (n, V)
    form1 = 300.0 + 600.0 * _hash_uniform(key_sv + 1)
    form2 = 1000.0 + 1600.0 * _hash_uniform(key_sv + 2)
    h = torch.arange(1, N_HARM + 1, device=device, dtype=f32)
    fh = f0[:, None, None] * h[None, None]                                               # (n, 1, H)
    amp = (torch.exp(-((fh - form1[..., None]) / 180.0) ** 2) + 0.7 * torch.exp(-((fh - form2[..., None]) / 260.0) ** 2)
           + 0.03) / h[None, None]                                                       # (n, V, H)
    n_seg = (n_samples + SEG_SAMPLES - 1) // SEG_SAMPLES
    seg = torch.arange(n_seg, device=device, dtype=torch.int64)
    key_us = spk[:, None] * 1000003 + utt[:, None] * 7717 + seg[None] * 13               # (n, n_seg)
    vowel = _hash_u32(key_us + 3) % N_VOWELS
    gain = 0.6 + 0.8 * _hash_uniform(key_us + 4)
    amp_seg = torch.gather(amp, 1, vowel[..., None].expand(n, n_seg, N_HARM)) * gain[..., None]   # (n, n_seg, H)
    t = torch.arange(n_samples, device=device, dtype=f32) / float(fs)
    seg_of_t = torch.arange(n_samples, device=device, dtype=torch.int64) // SEG_SAMPLES
    sig = torch.zeros((n, n_samples), dtype=f32, device=device)
    phase0 = 6.2831853 * _hash_uniform(spk[:, None] * 613 + utt[:, None] * 17 + torch.arange(N_HARM, device=device)[None] + 9)
    for k in range(N_HARM):
        a_t = amp_seg[:, :, k][:, seg_of_t]                                              # (n, n_samples)
        sig += a_t * torch.sin(6.2831853 * (k + 1) * f0[:, None] * t[None] + phase0[:, k : k + 1])
    sample = torch.arange(n_samples, device=device, dtype=torch.int64)
    noise = _hash_uniform((spk[:, None] * 9176 + utt[:, None]) * 1048583 + sample[None]) - 0.5
    sig += 0.35 * sig.pow(2).mean(dim=1, keepdim=True).sqrt() * noise
    sig *= 3000.0 / (sig.pow(2).mean(dim=1, keepdim=True).sqrt() + 1e-12)
    pcm = sig.round().clamp(-32768, 32767).to(torch.int16)
    return torch.where(pcm == 0, torch.ones_like(pcm), pcm)

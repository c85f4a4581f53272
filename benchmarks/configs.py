#!/usr/bin/env python
"""Secondary measurements for the BASELINE.json configs that are not the bench.py headline
(SURVEY section 8(d)): config 2 (front-end throughput), config 3 (UBM EM), config 5 (scoring sweep
vs utterance length).  One JSON line per measurement; CUDA events on the launching stream, >= 3 warm-ups,
inputs larger than L2.

    python benchmarks/configs.py [--only 2,3,5] [--scale 1.0]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def timed(fn, steps, warmup):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def synth_pcm(n_utts, n_samples, dev, seed=0):
    import torch

    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    pcm = torch.empty(n_utts * n_samples, dtype=torch.int16, device=dev)
    tt = torch.arange(n_samples, device=dev, dtype=torch.float32) / 16000.0
    step = max(1, min(n_utts, (1 << 26) // n_samples))
    for lo in range(0, n_utts, step):
        hi = min(n_utts, lo + step)
        f0 = 80 + 170 * torch.rand((hi - lo, 1), generator=g, device=dev)
        sig = torch.zeros((hi - lo, n_samples), device=dev)
        for h in range(1, 6):
            sig += torch.sin(2 * np.pi * h * f0 * tt[None]) / h
        sig += 0.3 * torch.randn((hi - lo, n_samples), generator=g, device=dev)
        sig *= 3000.0 / sig.pow(2).mean(dim=1, keepdim=True).sqrt()
        pcm[lo * n_samples : hi * n_samples] = sig.round().clamp(-32768, 32767).to(torch.int16).flatten()
    return pcm


def config2(scale):
    """Batched MFCC+delta+delta-delta (39-d, CMVN) over 100k synthetic 3 s utterances."""
    import torch

    import speech_signal_processing_b200 as ssp

    dev = torch.device("cuda")
    n_utts, n_samp = int(100000 * scale), 48000
    pcm = synth_pcm(n_utts, n_samp, dev)
    offs = np.arange(n_utts + 1, dtype=np.int64) * n_samp
    fe = ssp.FrontEnd(ssp.sidekit_recipe(), delta_order=2, cmvn=True)
    out = torch.empty((n_utts * 298, 39), dtype=torch.float32, device=dev)
    ms = timed(lambda: fe.extract_device(pcm, offs, out=out), steps=5, warmup=3)
    frames = n_utts * 298
    pk, src = peaks()
    algo_bytes = 2 * n_samp * n_utts + 4 * 39 * frames  # 478 B/frame
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    return {"config": "2: MFCC+d+dd 39-d, 100k x 3 s", "n_utts": n_utts, "ms": ms, "frames_per_s": frames / (ms * 1e-3),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                         "peak_source": src, "bytes_per_frame": algo_bytes / frames}}


def config3(scale):
    """512-comp UBM EM on ~100 h of 39-d frames (36 M frames): time per EM iteration."""
    import torch

    import speech_signal_processing_b200 as ssp
    from speech_signal_processing_b200 import synth

    dev = torch.device("cuda")
    n, d, k = int(36_000_000 * scale), 39, 512
    w, mu, var = synth.synth_ubm(64, d, seed=0)
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    comp = torch.randint(0, 64, (n,), generator=g, device=dev)
    x = torch.as_tensor(mu, device=dev, dtype=torch.float32)[comp]
    x += torch.as_tensor(np.sqrt(var), device=dev, dtype=torch.float32)[comp] * torch.randn((n, d), generator=g, device=dev)
    del comp
    rs = np.random.RandomState(0)
    means_init = x[torch.as_tensor(rs.choice(n, k, replace=False), device=dev)].cpu().numpy().astype(np.float64)
    prec_init = np.tile(1.0 / x[: 1 << 20].var(dim=0).cpu().numpy().astype(np.float64), (k, 1))
    iters = 10
    gm = ssp.GaussianMixture(n_components=k, covariance_type="diag", weights_init=np.full(k, 1.0 / k), means_init=means_init,
                             precisions_init=prec_init, max_iter=iters, tol=0.0)
    import warnings

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ssp.GaussianMixture(n_components=k, weights_init=np.full(k, 1.0 / k), means_init=means_init, precisions_init=prec_init,
                            max_iter=2, tol=0.0).fit(x)  # warm-up
        torch.cuda.synchronize()
        e0.record()
        gm.fit(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flop = 8.0 * d * k * n  # logits + statistics
    mono = all(b >= a - 1e-4 for a, b in zip(gm.lower_bounds_, gm.lower_bounds_[1:]))
    return {"config": "3: 512-comp UBM EM, 36 M x 39-d frames", "n_frames": n, "ms_per_iteration": ms,
            "frames_per_s": n / (ms * 1e-3), "algorithmic_tflops": flop / (ms * 1e-3) / 1e12, "lower_bound_monotone": mono,
            "lower_bounds": [round(b, 4) for b in gm.lower_bounds_],
            "roofline": {"bound": "tensor", "achieved": flop / (ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                         "executed_tflops": 2.0 * 80 * k * (3 + 3 + 2) * n / (ms * 1e-3) / 1e12},
            "note": "tcgen05 3xTF32 E-step (logits evaluated in both passes: log-sum-exp, then statistics) + TMEM-accumulated "
                    "statistics GEMM (hi + lo); achieved = algorithmic 8*D*K FLOP per frame, executed = 2*80*K*(3+3+2)"}


def config5(scale):
    """2048-comp scoring sweep vs utterance length (1-30 s), ~3 M frames held constant."""
    import torch

    import speech_signal_processing_b200 as ssp
    from speech_signal_processing_b200 import synth

    dev = torch.device("cuda")
    k, d = 2048, 39
    w, mu, var = synth.synth_ubm(k, d, seed=0)
    out = []
    pk, src = peaks()
    for n_models in (1, 101):
        spk = synth.synth_speaker_means(mu, n_models, seed=1, shift=0.25) if n_models > 1 else mu[None]
        ms_set = ssp.ModelSet(np.tile(w, (n_models, 1)), spk, np.tile(var, (n_models, 1, 1)))
        for secs in (1, 2, 3, 5, 10, 20, 30):
            t = (secs * 16000 - 400) // 160 + 1
            n_utts = max(1, int(3_000_000 * scale) // t)
            total = n_utts * t
            x = torch.randn((total, d), device=dev)
            offs = np.arange(n_utts + 1, dtype=np.int64) * t
            ms = timed(lambda: ms_set.score(x, offs, precision="tf32"), steps=3, warmup=3)
            tf = 4.0 * d * k * total * n_models / (ms * 1e-3) / 1e12
            out.append({"config": "5: 2048-comp scoring sweep", "seconds": secs, "frames_per_utt": int(t), "n_utts": n_utts,
                        "n_models": n_models, "ms": ms, "frames_per_s": total / (ms * 1e-3), "tflops": tf,
                        "frac_tf32_roofline": tf / (pk["bf16_tflops_sustained"] / 2)})
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="2,3,5")
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    for c in a.only.split(","):
        res = {"2": config2, "3": config3, "5": config5}[c.strip()](a.scale)
        for r in res if isinstance(res, list) else [res]:
            print(json.dumps(r), flush=True)

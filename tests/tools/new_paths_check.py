"""Timing + error margins of the front-end paths added late in round 1 (direct-DFT frame sizes, librosa convention).
    python tests/tools/new_paths_check.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import speech_signal_processing_b200 as ssp  # noqa: E402
from oracle import frontend as ofe  # noqa: E402  (checker only)
from speech_signal_processing_b200 import synth  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def scaled_err(got, want):
    return float((np.abs(got.astype(np.float64) - want) / np.maximum(1.0, np.abs(want))).max())


out = {}
dev = torch.device("cuda:0")
g = torch.Generator(device=dev)
g.manual_seed(0)
# ---- utils.processing.MFCC at 16 kHz, 400/160 (FFT length 400 -> direct DFT) vs 512/160 (FFT)
n_utts, n_samp = 4000, 48000
pcm = (torch.randn(n_utts * n_samp, generator=g, device=dev) * 3000).round().clamp(-32768, 32767).to(torch.int16)
offs = np.arange(n_utts + 1, dtype=np.int64) * n_samp
for fsz in (400, 512):
    fe = ssp.FrontEnd(ssp.processing_recipe(16000, fsz, 160))
    ms = timed(lambda: fe.extract_device(pcm, offs))
    frames = int(fe.frame_counts(np.diff(offs)).sum())
    out[f"processing_{fsz}_160"] = {"ms": ms, "frames_per_s": frames / ms * 1e3}
sig = synth.synth_utterance(3, 1, 48000)
out["processing_400_err"] = scaled_err(ssp.MFCC(sig, 16000, 400, 160), ofe.processing_mfcc(sig, 16000, 400, 160))
# ---- librosa convention at 8 kHz (n_fft 2048, hop 512, 128 mel bands)
n_utts, n_samp = 4000, 24000
pcm8 = torch.randn(n_utts * n_samp, generator=g, device=dev) * 3000
offs8 = np.arange(n_utts + 1, dtype=np.int64) * n_samp
mfe = ssp.MelDbFrontEnd()
d_off = None


def run_librosa():
    mel_db, foffs, _ = mfe.bands.extract_device(pcm8, offs8)
    ceps = torch.empty((int(foffs[-1]), 13), dtype=torch.float32, device=dev)
    t_off = torch.as_tensor(foffs, device=dev)
    ssp._lib.check(mfe.lib.ssp_mel_db_post(ssp._lib.ptr(mel_db), ssp._lib.ptr(t_off), len(foffs) - 1, mel_db.shape[1], 13,
                                           ssp._lib.ptr(mfe.t_dct), mfe.top_db, ssp._lib.ptr(ceps), ssp._lib.stream_ptr()), "post")
    return ceps


ms = timed(run_librosa)
out["librosa_2048_512"] = {"ms": ms, "utts": n_utts, "samples_per_s": n_utts * n_samp / ms * 1e3}
errs = []
for n in (12000, 24000, 3000, 700):
    s8 = synth.synth_utterance(6, n % 11, n, 8000)
    errs.append(scaled_err(ssp.MFCC_lib(s8), ofe.mfcc_lib(s8)))
out["librosa_err"] = errs
print(json.dumps(out))

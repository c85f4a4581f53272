"""Tiny driver for ncu / timing: EM statistics passes (K=512, D=39) on synthetic frames, operand images reused."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_signal_processing_b200 as ssp
from speech_signal_processing_b200 import synth, _lib
from speech_signal_processing_b200.mixture import StatsWorkspace

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
k, d = 512, 39
w, mu, var = synth.synth_ubm(k, d, seed=0)
x = torch.randn((n, d), device="cuda")
ms = ssp.ModelSet(w, mu, var)
ws = StatsWorkspace(x.device)
seg = np.array([0, n])
ms.stats(x, seg, workspace=ws, reuse_images=True)   # builds the images
torch.cuda.synchronize()
_lib.load().ssp_reset_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ms.stats(x, seg, workspace=ws, reuse_images=True)
e1.record()
torch.cuda.synchronize()
ms_per = e0.elapsed_time(e1) / reps
print(f"{n} frames: {ms_per:.3f} ms per statistics call = {n / ms_per / 1e3:.1f} M frames/s; 36 M frames -> {ms_per * 36e6 / n:.1f} ms; launches {_lib.launch_log()}")

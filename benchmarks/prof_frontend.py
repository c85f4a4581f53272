"""Tiny driver for ncu / timing: the fused front-end (sidekit recipe, 39-d, CMVN) on N synthetic 3 s utterances."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_signal_processing_b200 as ssp
from benchmarks.configs import synth_pcm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
pcm = synth_pcm(n, 48000, torch.device("cuda"))
offs = np.arange(n + 1, dtype=np.int64) * 48000
fe = ssp.FrontEnd(ssp.sidekit_recipe(), delta_order=2, cmvn=True)
out = torch.empty((n * 298, 39), dtype=torch.float32, device="cuda")
for _ in range(3):
    fe.extract_device(pcm, offs, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    fe.extract_device(pcm, offs, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{n} utterances: {ms:.3f} ms = {n * 298 / ms / 1e3:.1f} M frames/s")

"""Tiny driver for ncu: a few EM statistics passes (K=512, D=39) on synthetic frames."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_signal_processing_b200 as ssp
from speech_signal_processing_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
k, d = 512, 39
w, mu, var = synth.synth_ubm(k, d, seed=0)
x = torch.randn((n, d), device="cuda")
ms = ssp.ModelSet(w, mu, var)
for _ in range(3):
    ms.stats(x, np.array([0, n]))
torch.cuda.synchronize()

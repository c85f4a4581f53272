"""Tiny driver for ncu: run the fused front-end on N synthetic 3 s utterances a few times."""
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
for _ in range(3):
    fe.extract_device(pcm, offs)
torch.cuda.synchronize()

import sys, os, numpy as np, torch
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import speech_signal_processing_b200 as ssp
from speech_signal_processing_b200 import synth
k, d, n_models = 2048, 39, 101
w, mu, var = synth.synth_ubm(k, d, seed=0)
spk = np.concatenate([synth.synth_speaker_means(mu, n_models - 1, seed=1, shift=0.25), mu[None]])
ms = ssp.ModelSet(np.tile(w, (n_models, 1)), spk, np.tile(var, (n_models, 1, 1)))
t, n = 298, 10000
x = torch.randn((n * t, d), device="cuda"); offs = np.arange(n + 1, dtype=np.int64) * t
for prec in ("tf32", "tf32x3"):
    for _ in range(2): ms.score(x, offs, precision=prec)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): ms.score(x, offs, precision=prec)
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get("SSP_B200_LIB", "default").split("/")[-1], prec, round(e0.elapsed_time(e1) / 3, 2), "ms")

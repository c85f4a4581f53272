import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import speech_signal_processing_b200 as ssp
from speech_signal_processing_b200 import synth
from oracle import gmm as ogmm
k, d = int(sys.argv[1]), int(sys.argv[2])
w, mu, var = synth.synth_ubm(k, d, seed=k + d)
x = synth.sample_gmm(w, mu * 0.8, var * 1.2, 300, seed=1)
ms = ssp.ModelSet(w, mu, var)
n, f, s, ll = ms.stats(torch.as_tensor(x, device="cuda"), np.array([0, 300]))
torch.cuda.synchronize()
rn, rf, rs, rll = ogmm.suff_stats(x.astype(np.float64), w, mu, var)
print("ll", float(ll[0]), rll)
print("lse[:4]", ms._keep[1][:4].cpu().numpy(), ogmm.score_samples(x[:4].astype(np.float64), w, mu, var))
print("n ", n.cpu().numpy()[0]); print("rn", rn)
print("f0", f.cpu().numpy()[0, 0]); print("rf0", rf[0])
print("s0", s.cpu().numpy()[0, 0]); print("rs0", rs[0])

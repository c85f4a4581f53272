"""Error of the shared-variance scoring kernel against the float64 oracle, per configuration: largest relative error of a
per-utterance score, largest absolute error of a log-likelihood ratio against the reference member (GMM_UBM.py:194),
largest relative error of a per-frame log-likelihood.  Sets the tolerances of tests/test_gpu_gmm.py.
    gpurun -- 'python tests/sv_precision_probe.py'   (lives under tests/: it calls the oracle)
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import speech_signal_processing_b200 as ssp  # noqa: E402
from oracle import gmm as ogmm  # noqa: E402
from speech_signal_processing_b200 import synth  # noqa: E402

for k, d, n_spk, shift, lens in ((1024, 39, 12, 0.25, [298] * 8 + [98, 64, 33, 1]), (2048, 39, 4, 0.25, [64, 298, 130, 257]),
                                 (64, 26, 10, 0.25, [298] * 6 + [50, 7]), (64, 26, 10, 1.0, [298] * 6), (512, 39, 6, 0.5, [298] * 6),
                                 (200, 13, 5, 0.25, [100] * 6)):
    w, mu, var = synth.synth_ubm(k, d, seed=61)
    spk_mu = np.concatenate([synth.synth_speaker_means(mu, n_spk, seed=62, shift=shift), mu[None]])
    utts = [synth.sample_gmm(w, spk_mu[i % n_spk], var, n, seed=900 + i) for i, n in enumerate(lens)]
    x = np.concatenate(utts)
    want = np.array([[ogmm.score(u, w, m, var) for m in spk_mu] for u in utts])
    want_llr = want[:, :n_spk] - want[:, n_spk:]
    row = {"K": k, "D": d, "models": n_spk + 1, "shift": shift}
    sms = ssp.SharedModelSet(w, var, spk_mu, ref_model=n_spk)
    feats, offs = ssp.mixture.concat_utterances(utts, sms.device)
    got, lse = sms.score(feats, offs, want_frame_lse=True)
    got, lse = got.cpu().numpy(), lse.cpu().numpy()
    row["sv_score_rel"] = float(np.abs(got / want - 1).max())
    row["sv_llr_abs"] = float(np.abs(got[:, :n_spk] - got[:, n_spk:] - want_llr).max())
    row["sv_frame_rel"] = float(max(np.abs(lse[i] / ogmm.score_samples(x, w, spk_mu[i], var) - 1).max() for i in (0, n_spk)))
    ms = sms.expand()
    for prec in ("tf32", "tf32x2"):
        g2 = ms.score(feats, offs, precision=prec)[0].cpu().numpy()
        row[prec + "_score_rel"] = float(np.abs(g2 / want - 1).max())
        row[prec + "_llr_abs"] = float(np.abs(g2[:, :n_spk] - g2[:, n_spk:] - want_llr).max())
    print(json.dumps(row), flush=True)

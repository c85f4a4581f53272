"""Multi-GPU paths over NCCL (needs >= 2 GPUs: run with `gpurun --gpus 2`): frame-sharded UBM EM with the
N/F/S all-reduce == single-GPU EM on the pooled frames; utterance-sharded identify == single-GPU identify."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys, warnings
import numpy as np, torch
sys.path.insert(0, {root!r})
import speech_signal_processing_b200 as ssp
from speech_signal_processing_b200 import synth
from speech_signal_processing_b200.dist import Comm, shard_range, fit_ubm_sharded, identify_sharded, map_enrol_sharded
comm = Comm("nccl")
k, d = 32, 13
w, mu, var = synth.synth_ubm(k, d, seed=3)
x = synth.sample_gmm(w, mu * 0.9, var * 1.2, 20001, seed=4)
lo, hi = shard_range(len(x), comm.rank, comm.world_size)
kw = dict(weights_init=w, means_init=mu, precisions_init=1.0 / var, max_iter=5, tol=0.0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    sharded = fit_ubm_sharded(comm, x[lo:hi], k, **kw)
    single = ssp.GaussianMixture(n_components=k, covariance_type="diag", **kw).fit(x)
assert sharded.n_iter_ == single.n_iter_ == 5
assert abs(sharded.lower_bound_ - single.lower_bound_) < 1e-6 * abs(single.lower_bound_)
np.testing.assert_allclose(sharded.means_, single.means_, atol=5e-5)  # FP32 accumulation order in TMEM differs with the block layout of a shard; 5 iterations
np.testing.assert_allclose(sharded.covariances_, single.covariances_, rtol=2e-4, atol=2e-6)
np.testing.assert_allclose(sharded.weights_, single.weights_, atol=1e-6)
# k-means initialisation is identical on every rank (seeds broadcast from rank 0, statistics all-reduced)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    g2 = fit_ubm_sharded(comm, x[lo:hi], 8, max_iter=3, random_state=0)
m = torch.as_tensor(g2.means_, device="cuda")
m0 = m.clone(); comm.broadcast(m0, 0)
assert torch.equal(m, m0)
# utterance-sharded identify
spk = synth.synth_speaker_means(mu, 5, seed=9, shift=0.4)
models = [ssp.GaussianMixture.from_params(w, spk[i], var) for i in range(5)]
utts = [synth.sample_gmm(w, spk[i % 5], var, 50 + 7 * i, seed=20 + i) for i in range(11)]
ubm = ssp.GaussianMixture.from_params(w, mu, var)
full, who = identify_sharded(comm, utts, models, ubm)
ref, who_ref = ssp.identify(utts, models, ubm)
np.testing.assert_allclose(full, ref, atol=2e-5)  # fp32 partial-sum order differs with the frame alignment
assert (who == who_ref).all() and (who == np.arange(11) % 5).all()
# speaker-sharded MAP enrolment (adapted means all-gathered) + utterance-sharded identify with the shared-variance kernel
enrol = [synth.sample_gmm(w, spk[i], var, 400 + 13 * i, seed=40 + i) for i in range(5)]
sms = map_enrol_sharded(comm, ubm, enrol, relevance=16.0)
sms_ref = ssp.map_enrol(ubm, enrol, relevance=16.0)
assert sms.n_models == 6 and sms.ubm_index == 5
np.testing.assert_allclose(sms._params[2].cpu().numpy(), sms_ref._params[2].cpu().numpy(), rtol=0, atol=5e-6)  # TMEM fp32 accumulation order differs with the chunking
full2, who2 = identify_sharded(comm, utts, sms)
ref2, who_ref2 = ssp.identify(utts, sms_ref)
np.testing.assert_allclose(full2, ref2, atol=2e-5)
assert (who2 == who_ref2).all() and (who2 == np.arange(11) % 5).all()
comm.barrier()
torch.distributed.destroy_process_group()
print("rank", comm.rank, "ok")
"""


def test_sharded_em_and_identify_nccl(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    port = 29700 + (os.getpid() % 200)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True))
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o[-3000:]
        assert f"rank {r} ok" in o

#!/usr/bin/env python
"""bench.py -- frames/sec of MFCC + GMM-UBM scoring (1024 components) on B200.

    python bench.py --gpus N --steps K --warmup W        # this repo's CUDA path
    python bench.py --impl reference ...                  # the reference's CPU path (oracle port + sklearn)

Workload (BASELINE.json configs[3], the one the metric is quoted on): per GPU, 10 000 test utterances
of 3 s @ 16 kHz int16 -> 39-d MFCC+delta+delta-delta with per-utterance CMVN (298 frames each, 2.98 M
frames) -> scored against 1 000 MAP-enrolled speaker models + the UBM (1 001 models x 1 024 diagonal
components) -> LLR argmax per utterance.  One "step" = one pass over that batch.  Multi-GPU: every rank
owns its own batch of utterances, all models replicated, no data-path collective (weak scaling).

`value`  = frames/s with the PCM already resident in HBM (front-end kernel + scoring kernels + argmax).
`e2e`    = frames/s through the public host-to-host call `ssp.identify_pcm` from pinned HOST PCM (H2D copy inside the
           timed region, its tail overlapped with the kernels of the head part) to the decisions read back on the host
           (D2H inside the timed region).
`roofline` = the tcgen05 scoring kernel against the tensor roofline: algorithmic 4*D*K FLOP per
           (frame, model) / its CUDA-event time / the measured peak.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec MFCC+GMM-UBM scoring (1024 comp)"
UNIT = "frames/s"
K_COMP, N_SPK, N_UTT, UTT_SAMPLES, DIM = 1024, 1000, 10000, 48000, 39
FRAMES_PER_UTT = 298


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--utts", type=int, default=N_UTT, help="test utterances per GPU (default = the named config)")
    ap.add_argument("--speakers", type=int, default=N_SPK)
    ap.add_argument("--components", type=int, default=K_COMP)
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"])
    ap.add_argument("--scorer", default="shared", choices=["shared", "general"],
                    help="shared: the shared-variance tensor kernel (mean-only MAP speakers keep the UBM's weights and "
                         "variances); general: the kernel for arbitrary model sets")
    ap.add_argument("--cpu-utts", type=int, default=0, help="reference arm: utterances per step (0 = auto-size)")
    ap.add_argument("--cpu-budget", type=float, default=100.0, help="reference arm: seconds of CPU work for the whole run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"config4: {a.components}-comp {DIM}-d GMM-UBM identify, {a.utts} utts x {FRAMES_PER_UTT} frames (3 s @16 kHz) "
            f"per GPU vs {a.speakers} MAP speakers + UBM")


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU path (sidekit-recipe MFCC restatement + sklearn GMM scoring)
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_worker_init(w, mu, var, ubm):
    from threadpoolctl import threadpool_limits

    _W["lim"] = threadpool_limits(1)
    _W["params"] = (w, mu, var, ubm)


def _cpu_score_task(task):
    """score one feature matrix against a slice of speaker models with sklearn (GMM_UBM.py:194)."""
    from sklearn.mixture import GaussianMixture

    feat, lo, hi = task
    w, mu, var, ubm = _W["params"]
    out = np.empty(hi - lo)
    gm = GaussianMixture(n_components=len(w), covariance_type="diag")
    for i in range(lo, hi):
        gm.weights_, gm.means_, gm.covariances_ = w, mu[i], var
        gm.precisions_cholesky_ = 1.0 / np.sqrt(var)
        out[i - lo] = gm.score(feat)
    return lo, out


def _cpu_frontend_task(sig):
    from oracle import frontend as ofe

    return ofe.features(sig, preset="sidekit", delta_order=2, cmvn=True)


def run_reference(a):
    """Times GMM_UBM.py's per-utterance recipe on the host cores: extract_feature (:89-93, sidekit
    restatement + delta x2 + scale) then GMM[i].score(x) - UBM.score(x) for every model (:191-197), on a
    bounded sample of the workload's utterances per step, all host cores busy."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from concurrent.futures import ProcessPoolExecutor

    from sklearn.mixture import GaussianMixture

    from speech_signal_processing_b200 import synth

    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    w, mu, var = synth.synth_ubm(a.components, DIM, seed=0)
    spk = synth.synth_speaker_means(mu, a.speakers, seed=1, shift=0.25)
    n_models = a.speakers
    ubm = GaussianMixture(n_components=a.components, covariance_type="diag")
    ubm.weights_, ubm.means_, ubm.covariances_, ubm.precisions_cholesky_ = w, mu, var, 1.0 / np.sqrt(var)

    with ProcessPoolExecutor(procs, initializer=_cpu_worker_init, initargs=(w, spk, var, None)) as pool:
        # calibrate: one utterance against 2 models per worker
        sig0 = synth.synth_utterance(0, 0, UTT_SAMPLES)
        t0 = time.perf_counter()
        f0 = _cpu_frontend_task(sig0)
        list(pool.map(_cpu_score_task, [(f0, i % (n_models - 1), i % (n_models - 1) + 2) for i in range(0, 2 * procs, 2)]))
        per_pair = (time.perf_counter() - t0) / 2.0  # wall seconds per (utt, model) per worker
        budget = a.cpu_budget / max(1, a.steps + a.warmup)   # whole run within a few minutes
        n_utt = a.cpu_utts or int(max(1, min(64, budget / max(1e-6, per_pair * (n_models + 1) / procs))))
        sigs = [synth.synth_utterance(s % 50, s // 50, UTT_SAMPLES) for s in range(n_utt)]
        chunk = max(1, (n_models + procs - 1) // procs)

        def step():
            feats = list(pool.map(_cpu_frontend_task, sigs))
            tasks = [(f, lo, min(lo + chunk, n_models)) for f in feats for lo in range(0, n_models, chunk)]
            pred = np.zeros((n_utt, n_models))
            for j, (lo, out) in enumerate(pool.map(_cpu_score_task, tasks)):
                pred[j // ((n_models + chunk - 1) // chunk), lo : lo + len(out)] = out
            base = np.array([ubm.score(f) for f in feats])
            return (pred - base[:, None]).argmax(axis=1), sum(len(f) for f in feats)

        for _ in range(a.warmup):
            step()
        t0 = time.perf_counter()
        frames = 0
        for _ in range(a.steps):
            _, nf = step()
            frames += nf
        dt = time.perf_counter() - t0
    val = frames / dt
    sample = f"{n_utt} of {a.utts} utterances per step x all {n_models}+1 models, {procs} worker processes (1 BLAS thread each)"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port", "sample": sample,
                         "host_cpu_count": cores},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


def cpu_baseline_subprocess(a):
    """The oracle-port CPU baseline, run in a child BEFORE this process touches CUDA (it forks workers)."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1", "--warmup", "1",
                              "--cpu-budget", "30", "--utts", str(a.utts), "--speakers", str(a.speakers), "--components", str(a.components)],
                             capture_output=True, text=True, timeout=600, env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})
        line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:  # the baseline is reported, never required
        return {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": f"failed: {e!r}"[:200]}


def run_b200(a):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cpu_base = cpu_baseline_subprocess(a)

    import torch

    import speech_signal_processing_b200 as ssp
    from speech_signal_processing_b200 import _lib, synth

    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    comm = None
    if world > 1:
        from speech_signal_processing_b200.dist import Comm

        comm = Comm("nccl")

    def barrier():
        if comm is not None:
            comm.barrier()
        torch.cuda.synchronize()

    K, S, N, D = a.components, a.speakers, a.utts, DIM
    # ---- models: synthetic UBM; speakers enrolled by MAP (means only, r = 16) from 10 utterances each
    w, mu, var = synth.synth_ubm(K, D, seed=0)
    ubm = ssp.GaussianMixture.from_params(w, mu, var)
    t_mu = torch.as_tensor(synth.synth_speaker_means(mu, S, seed=1, shift=0.25), device=dev)
    t_var = torch.as_tensor(var, device=dev)
    enrol_frames = 10 * FRAMES_PER_UTT
    labels = torch.arange(S, device=dev).repeat_interleave(enrol_frames)
    enrol = synth.synth_features_torch(S * enrol_frames, D, t_mu, t_var, labels, seed=7, device=dev)
    seg = np.arange(S + 1, dtype=np.int64) * enrol_frames
    sw, smu, svar = ssp.map_adapt(ubm, (enrol, seg), relevance=16.0)
    del enrol, labels, t_mu
    models = ssp.ModelSet(torch.cat([sw, torch.as_tensor(w, device=dev)[None]]),
                          torch.cat([smu, torch.as_tensor(mu, device=dev)[None]]),
                          torch.cat([svar, t_var[None]]), device=dev)  # model S is the UBM
    shared = a.scorer == "shared" and a.precision == "tf32"
    scorer = ssp.SharedModelSet(torch.as_tensor(w, device=dev), t_var, torch.cat([smu, torch.as_tensor(mu, device=dev)[None]]),
                                ref_model=S, device=dev) if shared else models
    del sw, smu, svar
    # ---- test audio: synthetic int16 PCM, distinct per rank, kept both in HBM and in pinned host memory
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    pcm = torch.empty(N * UTT_SAMPLES, dtype=torch.int16, device=dev)
    tt = torch.arange(UTT_SAMPLES, device=dev, dtype=torch.float32) / 16000.0
    for lo in range(0, N, 500):
        hi = min(N, lo + 500)
        f0 = 80 + 170 * torch.rand((hi - lo, 1), generator=g, device=dev)
        sig = torch.zeros((hi - lo, UTT_SAMPLES), device=dev)
        for h in range(1, 12):
            sig += torch.sin(2 * np.pi * h * f0 * tt[None]) / h * torch.rand((hi - lo, 1), generator=g, device=dev)
        sig += 0.3 * torch.randn((hi - lo, UTT_SAMPLES), generator=g, device=dev)
        sig *= 3000.0 / sig.pow(2).mean(dim=1, keepdim=True).sqrt()
        pcm[lo * UTT_SAMPLES : hi * UTT_SAMPLES] = sig.round().clamp(-32768, 32767).to(torch.int16).flatten()
        del sig
    host_pcm = torch.empty(N * UTT_SAMPLES, dtype=torch.int16, pin_memory=True)
    host_pcm.copy_(pcm)
    sample_offsets = np.arange(N + 1, dtype=np.int64) * UTT_SAMPLES
    fe = ssp.FrontEnd(ssp.sidekit_recipe(), delta_order=2, cmvn=True, device=dev)
    host_dec = torch.empty(N, dtype=torch.int64, pin_memory=True)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    score_ms = []

    def device_step(pcm_dev, timed=False):
        feats, foffs, _ = fe.extract_device(pcm_dev, sample_offsets)
        if timed:
            ev[2].record()
        scores, _ = scorer.score(feats, foffs, precision=a.precision)
        if timed:
            ev[3].record()
        llr = scores[:, :S] - scores[:, S:]
        return llr.argmax(dim=1), int(foffs[-1])

    def e2e_step():
        # the public host-to-host call: H2D of this step's PCM (the tail of the copy overlaps the kernels of the head),
        # front-end, scoring, LLR argmax, D2H of the decisions; returns once they have landed in host_dec
        _, nf = ssp.identify_pcm(host_pcm, sample_offsets, fe, scorer, ubm_index=S, precision=a.precision, out=host_dec)
        return nf

    for _ in range(a.warmup):
        device_step(pcm)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib = _lib.load()
    lib.ssp_reset_launch_count()
    # ---- timed region 1: inputs resident in HBM
    barrier()
    ev[0].record()
    frames = 0
    for _ in range(a.steps):
        dec, nf = device_step(pcm, timed=True)
        frames += nf
        torch.cuda.current_stream().synchronize()
        score_ms.append(ev[2].elapsed_time(ev[3]))
    ev[1].record()
    barrier()
    dev_ms = ev[0].elapsed_time(ev[1])
    launches = int(lib.ssp_launch_count())
    # ---- timed region 2: end to end from host buffers
    for _ in range(min(a.warmup, 2)):
        e2e_step()
    barrier()
    ev[0].record()
    for _ in range(a.steps):
        e2e_step()
    ev[1].record()
    barrier()
    e2e_ms = ev[0].elapsed_time(ev[1])
    clocks = sampler.stop() if rank == 0 else None
    # host-to-host decisions vs HBM-resident decisions.  The two paths cut the batch differently, so the FP32 partial
    # sums inside a warp group other frames of an utterance (the cross-warp accumulation is float64): scores agree to
    # ~1e-6 absolute and a decision can only differ where the top-2 LLR margin is below that.  Report the count and
    # the largest margin among differing utterances instead of a bare flag.
    e2e_dec = host_dec.to(dev)
    differ = (e2e_dec != dec).nonzero().flatten()
    e2e_check = {"utts": int(dec.numel()), "differ": int(differ.numel()), "largest_llr_margin_among_differing": 0.0}
    if differ.numel():
        feats, foffs, _ = fe.extract_device(pcm, sample_offsets)
        sc, _ = scorer.score(feats, foffs, precision=a.precision)
        llr = (sc[:, :S] - sc[:, S:])[differ]
        e2e_check["largest_llr_margin_among_differing"] = float(
            (llr.gather(1, dec[differ][:, None]) - llr.gather(1, e2e_dec[differ][:, None])).abs().max().item())
        del feats, sc, llr

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(frames)], dtype=torch.float64, device=dev)
    if comm is not None:
        comm.allreduce_max(t)
        comm.allreduce_sum(tot)
    dev_ms, e2e_ms = (float(v) for v in t.tolist())
    total_frames = float(tot.item())
    if rank != 0:
        return
    # GPU-vs-GPU sanity on a sub-sample: tensor-core decisions == FP32 CUDA-core decisions
    sub = 64
    feats, foffs, _ = fe.extract_device(pcm[: sub * UTT_SAMPLES], sample_offsets[: sub + 1])
    s_tc, _ = scorer.score(feats, foffs, precision=a.precision)
    s_fp, _ = models.score(feats, foffs, precision="fp32")
    dec_tc = (s_tc[:, :S] - s_tc[:, S:]).argmax(dim=1)
    dec_fp = (s_fp[:, :S] - s_fp[:, S:]).argmax(dim=1)
    rel = float(((s_tc - s_fp).abs() / s_fp.abs()).max().item())

    peaks, peak_src = measured_peaks()
    n_models = S + 1
    flop_per_launch = 4.0 * D * K * (frames / a.steps) * n_models
    k_ms = float(np.mean(score_ms))
    achieved = flop_per_launch / (k_ms * 1e-3) / 1e12
    if a.precision == "tf32":
        peak = peaks["bf16_tflops_sustained"] / 2.0
        peak_note = peak_src + ": bf16_tflops_sustained / 2 -- kind::tf32 MMAs issue at half the bf16 rate"
        bound = "tensor"
        try:  # a measured TF32 GEMM figure (benchmarks/measure_tf32_peak.py) beats the derived one
            with open(os.path.join(ROOT, "profiles", "tf32_peak.json")) as f:
                peak = float(json.load(f)["tf32_tflops_sustained"])
            peak_note = "measured (profiles/tf32_peak.json): torch.matmul TF32 8192^3 sustained"
        except (OSError, KeyError, ValueError):
            pass
    else:
        peak, peak_note, bound = 70.0, "nominal FP32 CUDA-core FMA peak (no measured figure)", "tensor"
    kernel = "gmm_score_sv_kernel" if shared else ("gmm_score_tc_kernel" if a.precision == "tf32" else "gmm_score_simt_kernel")
    roof_extra = {}
    if shared:
        # with the shared-variance shortcut the executed tensor work is smaller than the algorithmic 4DK:
        # per (frame, model) 2 * KS * Kp with KS = roundup(D + 2, 8), plus the common part once per 32 models
        ks, kp = (D + 2 + 7) // 8 * 8, (K + 63) // 64 * 64
        exec_flop = 2.0 * ks * kp * (frames / a.steps) * (n_models * (1.0 + 1.0 / 32.0) + 2.0)
        roof_extra = {"executed_tflops": exec_flop / (k_ms * 1e-3) / 1e12, "executed_frac": exec_flop / (k_ms * 1e-3) / 1e12 / peak,
                      "note": "achieved = algorithmic 4*D*K FLOP per (frame, model); the shared-variance kernel executes "
                              "2*(D+2 padded to 48)*K on the tensor pipe and is bound by the 3.05e12 exponentials of the "
                              "log-sum-exp (MUFU ex2 + an FMA-pipe polynomial share), see profiles/"}
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this size, from a committed ncu --set full capture
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(kernel)
        if t and (t["utts"], t["speakers"], t["components"]) == (N, S, K):
            traffic = {"bytes": t["dram_bytes"], "unit": "B per launch", "algorithmic_bytes": t.get("algorithmic_bytes"),
                       "source": t["source"]}
    except Exception:
        traffic = None
    line = {
        "metric": METRIC, "value": total_frames / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if a.precision == "tf32" else "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "parallelism": f"utterances sharded x{world}, models replicated",
                   "l2": "inputs per step (0.96 GB PCM, 0.2-0.33 GB model tiles) exceed the 126 MB L2", "scorer": a.scorer,
                   "frames_x_models_per_s": total_frames * n_models / (dev_ms * 1e-3)},
        "e2e": {"value": total_frames / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(host_pcm.numel() * 2),
                "d2h_bytes_per_step": int(host_dec.numel() * 8), "ms_per_step": e2e_ms / a.steps},
        "gpu_launches": launches,
        "roofline": {"bound": bound, "kernel": kernel, **roof_extra,
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                     "kernel_ms": k_ms, "kernel_share_of_step": k_ms / (dev_ms / a.steps), "peak_source": peak_note,
                     "frac_of_bf16_peak": achieved / peaks["bf16_tflops_sustained"]},
        "clocks": clocks,
        "check": {"tc_vs_fp32_max_rel": rel, "decisions_equal": bool((dec_tc == dec_fp).all().item()), "utts": sub,
                  "e2e_vs_device_decisions": e2e_check},
    }
    if cpu_base is not None:
        line["cpu_baseline"] = cpu_base
    print(json.dumps(line), flush=True)


def _shutdown():
    try:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
        _shutdown()

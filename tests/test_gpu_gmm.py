"""Parity of the GMM kernels (through the C ABI) with sklearn fixtures and the oracle."""
import os
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import speech_signal_processing_b200 as ssp  # noqa: E402
from oracle import gmm as ogmm  # noqa: E402
from speech_signal_processing_b200 import synth  # noqa: E402

# Stated tolerances (relative, per-utterance score = what GaussianMixture.score returns):
#   fp32 CUDA-core kernel: 2e-6 everywhere;
#   single-pass TF32 tensor kernel: 1e-4 at the named workloads (K = 1024, ~300 frames per utterance, see
#   test_tf32_matches_fp32_kernel_1024_components) and 5e-4 worst case for tiny models / very short utterances,
#   where the unbiased A-operand rounding averages over fewer frames and components.
#   3-pass TF32 (tf32x3, what "auto" picks below 512 components): FP32-grade, 2e-6; 2-pass (split model operand): the
#   frame-side rounding remains, same per-score bound as one pass but a 3-5x smaller LLR error.
#   shared-variance kernel (FP16 operands = TF32's significand; common part FP32 grade, speakers as differences from the
#   reference): 2e-4 worst case (1- and 2-frame utterances), 1e-4 / LLR 1e-3 from 31 frames on, the reference member 5e-6.
REL = {"fp32": 2e-6, "tf32": 5e-4, "tf32x2": 5e-4, "tf32x3": 2e-6, "sv": 2e-4}
# LLR within 1e-3 absolute, per-utterance score within 1e-4 relative
LLR_ATOL, SCORE_RTOL = 1e-3, 1e-4
# the kernels that must serve ssp_gmm_stats for D <= 39 (tcgen05), and the FP32 CUDA-core pair for wider features
EM_TENSOR_KERNELS = {"em_plan_kernel", "em_prep_kernel", "gmm_em_lse_kernel", "em_merge_kernel", "gmm_em_stats_kernel"}
EM_SIMT_KERNELS = {"gmm_score_simt_kernel", "gmm_stats_simt_kernel"}


@pytest.mark.parametrize("precision", ["fp32", "tf32", "tf32x2", "tf32x3"])
@pytest.mark.parametrize("tag", ["s", "m", "l"])
def test_score_matches_sklearn(golden, tag, precision):
    g = golden("sklearn_gmm.npz")
    gm = ssp.GaussianMixture.from_params(g[f"{tag}_w"], g[f"{tag}_mu"], g[f"{tag}_var"], precision=precision)
    x = g[f"{tag}_x"]
    want = g[f"{tag}_score_samples"]
    got = gm.score_samples(x)
    assert got.shape == want.shape
    # per frame: TF32 operand rounding (2^-12 relative on ~80 products of magnitude up to ~40) is a few 1e-2
    # absolute on |L| ~ 50; it is unbiased on the frame side and averages out in the utterance mean below
    np.testing.assert_allclose(got, want, rtol=5e-5 if precision in ("fp32", "tf32x3") else 3e-3, atol=0)
    s = gm.score(x)
    assert isinstance(s, float)
    assert abs(s - float(g[f"{tag}_score"])) <= REL[precision] * abs(float(g[f"{tag}_score"]))


@pytest.mark.parametrize("precision", ["fp32", "tf32", "tf32x2", "tf32x3"])
def test_score_matrix_ragged_matches_oracle(precision):
    k, d, n_spk = 64, 39, 5
    w, mu, var = synth.synth_ubm(k, d, seed=3)
    spk_mu = synth.synth_speaker_means(mu, n_spk, seed=4)
    lens = [1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 298, 500]
    utts = [synth.sample_gmm(w, spk_mu[i % n_spk], var, n, seed=100 + i) for i, n in enumerate(lens)]
    models = [ssp.GaussianMixture.from_params(w, spk_mu[i], var) for i in range(n_spk)]
    got = ssp.score_matrix(utts, models, precision=precision)
    want = np.array([[ogmm.score(u, w, spk_mu[i], var) for i in range(n_spk)] for u in utts])
    np.testing.assert_allclose(got, want, rtol=REL[precision] * (6 if precision in ("tf32", "tf32x2") else 1), atol=0)
    long_enough = np.array(lens) >= 31
    np.testing.assert_allclose(got[long_enough], want[long_enough], rtol=REL[precision], atol=0)
    assert (got.argmax(axis=1) == want.argmax(axis=1))[long_enough].all()


def test_tf32_matches_fp32_kernel_1024_components():
    """Property at scale: the tensor-core and CUDA-core kernels agree; argmax identical."""
    k, d, n_spk = 1024, 39, 24
    w, mu, var = synth.synth_ubm(k, d, seed=5)
    spk_mu = synth.synth_speaker_means(mu, n_spk, seed=6, shift=0.25)
    ms = ssp.ModelSet(np.tile(w, (n_spk, 1)), spk_mu, np.tile(var, (n_spk, 1, 1)))
    lens = [298] * 40 + [98, 777, 5]
    utts = [synth.sample_gmm(w, spk_mu[i % n_spk], var, n, seed=i) for i, n in enumerate(lens)]
    a = ssp.score_matrix(utts, ms, precision="fp32")
    b = ssp.score_matrix(utts, ms, precision="tf32")
    np.testing.assert_allclose(b[:42], a[:42], rtol=1e-4, atol=0)
    assert (a.argmax(axis=1) == b.argmax(axis=1))[:42].all()
    assert (a.argmax(axis=1)[:40] == np.arange(40) % n_spk).all()
    # oracle spot check on a few pairs (float64)
    for j in (0, 17, 41):
        for i in (0, 5):
            ref = ogmm.score(utts[j], w, spk_mu[i], var)
            assert abs(a[j, i] - ref) <= 2e-6 * abs(ref)
            assert abs(b[j, i] - ref) <= 1e-4 * abs(ref)


def test_tf32_stabiliser_redo_paths():
    """The tensor kernel carries its exp stabiliser from model to model; models whose likelihood for the same
    frames differs by thousands of nats (both directions) must take the redo path and still match float64."""
    k, d = 256, 13
    w, mu, var = synth.synth_ubm(k, d, seed=11)
    far = mu + 25.0          # every component ~ 25 sigma away: log-lik ~ -5000
    tight = var * 0.05       # sharp model: huge negative log-lik for most frames
    models_mu = np.stack([mu, far, mu, mu, far, mu * 0.5])
    models_var = np.stack([var, var, tight, var, var, var])
    ms = ssp.ModelSet(np.tile(w, (6, 1)), models_mu, models_var)
    utts = [synth.sample_gmm(w, mu, var, n, seed=200 + n) for n in (300, 257, 40)]
    got = ssp.score_matrix(utts, ms, precision="tf32")
    want = np.array([[ogmm.score(u, w, models_mu[i], models_var[i]) for i in range(6)] for u in utts])
    assert np.all(np.isfinite(got))
    np.testing.assert_allclose(got, want, rtol=5e-4)
    assert (got.argmax(axis=1) == want.argmax(axis=1)).all()


def test_split_utterance_property():
    """score(whole) == frame-weighted mean of score(parts) (mean of per-frame log-likelihoods)."""
    k, d = 128, 26
    w, mu, var = synth.synth_ubm(k, d, seed=8)
    x = synth.sample_gmm(w, mu, var, 1000, seed=9)
    ms = ssp.ModelSet(w, mu, var)
    for precision in ("fp32", "tf32"):
        whole = ssp.score_matrix([x], ms, precision=precision)[0, 0]
        parts = ssp.score_matrix([x[:300], x[300:], x[:0]], ms, precision=precision)[:, 0]
        assert abs(whole - (0.3 * parts[0] + 0.7 * parts[1])) < 1e-6 * abs(whole)
        assert parts[2] == 0.0  # empty utterance: no frames, score left at 0


@pytest.mark.parametrize("k,d", [(8, 5), (64, 26), (200, 39), (16, 13), (32, 60)])
def test_stats_match_oracle(k, d):
    w, mu, var = synth.synth_ubm(k, d, seed=k + d)
    lens = [0, 700, 64, 1, 130, 0, 2048 + 17]
    x = synth.sample_gmm(w, mu * 0.8, var * 1.2, sum(lens), seed=1)
    seg = np.concatenate([[0], np.cumsum(lens)])
    ms = ssp.ModelSet(w, mu, var)
    import torch

    from speech_signal_processing_b200 import _lib

    _lib.load().ssp_reset_launch_count()
    n, f, s, ll = ms.stats(torch.as_tensor(x, device="cuda"), seg)
    # the tensor-core kernels serve every D <= 39; only wider features take the FP32 CUDA-core path
    assert set(_lib.launch_log()) == (EM_TENSOR_KERNELS if d <= 39 else EM_SIMT_KERNELS), _lib.launch_log()
    n, f, s, ll = (t.cpu().numpy() for t in (n, f, s, ll))
    for i, ln in enumerate(lens):
        xs = x[seg[i] : seg[i + 1]].astype(np.float64)
        if ln == 0:
            assert np.all(n[i] == 0) and np.all(f[i] == 0) and ll[i] == 0
            continue
        rn, rf, rs_, rll = ogmm.suff_stats(xs, w, mu, var)
        np.testing.assert_allclose(n[i], rn, rtol=2e-4, atol=2e-4 * ln / k)
        np.testing.assert_allclose(f[i], rf, rtol=2e-4, atol=5e-4 * ln / k)
        np.testing.assert_allclose(s[i], rs_, rtol=2e-4, atol=1e-3 * ln / k)
        # the logits are 3 x BF16 (16-17 significant bits per product): the log-likelihood is good to ~1e-5 relative
        assert abs(ll[i] - rll) <= 1e-5 * abs(rll)
        assert abs(n[i].sum() - ln) < 1e-3 * ln  # posteriors sum to one per frame


@pytest.mark.parametrize("tag", ["s", "m"])
@pytest.mark.parametrize("iters", [1, 3, 100])
def test_em_trajectory_matches_sklearn(golden, tag, iters):
    """Same initial parameters in => same EM trajectory out."""
    import warnings

    g = golden("sklearn_gmm.npz")
    w, mu, var, x = g[f"{tag}_w"], g[f"{tag}_mu"], g[f"{tag}_var"], g[f"{tag}_x"]
    gm = ssp.GaussianMixture(n_components=len(w), covariance_type="diag", weights_init=w, means_init=mu,
                             precisions_init=1.0 / var, max_iter=iters)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gm.fit(x)
    assert gm.n_iter_ == int(g[f"{tag}_fit{iters}_niter"])
    assert gm.converged_ == bool(g[f"{tag}_fit{iters}_conv"])
    assert abs(gm.lower_bound_ - float(g[f"{tag}_fit{iters}_lb"])) < 5e-6 * abs(float(g[f"{tag}_fit{iters}_lb"]))
    # measured (tests/em_trajectory_probe.py): weights 7e-5 relative, means 1.4e-4 absolute, variances 1.2e-4 relative at
    # 1, 3 and 100 iterations alike -- the TF32 rounding of the posteriors in the statistics GEMM, which does not compound
    np.testing.assert_allclose(gm.weights_, g[f"{tag}_fit{iters}_w"], rtol=3e-4, atol=1e-6)
    np.testing.assert_allclose(gm.means_, g[f"{tag}_fit{iters}_mu"], rtol=0, atol=5e-4)
    np.testing.assert_allclose(gm.covariances_, g[f"{tag}_fit{iters}_var"], rtol=5e-4, atol=1e-6)
    np.testing.assert_allclose(gm.precisions_cholesky_, 1 / np.sqrt(gm.covariances_))


def test_fit_default_init_and_errors():
    w, mu, var = synth.synth_ubm(8, 6, seed=21, spread=4.0)
    x = synth.sample_gmm(w, mu, var * 0.3, 4000, seed=22)
    gm = ssp.GaussianMixture(n_components=8, covariance_type="diag", random_state=0).fit(x)
    assert gm.converged_ and gm.weights_.shape == (8,) and np.isclose(gm.weights_.sum(), 1.0)
    truth = ogmm.score(x, w, mu, var * 0.3)
    assert gm.score(x) > truth - 0.25  # well-separated clusters: EM from k-means gets near the truth
    assert all(b2 >= b1 - 1e-4 for b1, b2 in zip(gm.lower_bounds_, gm.lower_bounds_[1:]))  # EM is monotone
    with pytest.raises(ValueError, match="n_samples >= n_components"):
        ssp.GaussianMixture(n_components=64).fit(x[:10])
    with pytest.raises(NotImplementedError):
        ssp.GaussianMixture(n_components=2, covariance_type="full").fit(x)
    g2 = pickle.loads(pickle.dumps(gm))
    assert abs(g2.score(x) - gm.score(x)) < 1e-9
    sk = gm.to_sklearn()
    assert abs(sk.score(x.astype(np.float64)) - gm.score(x)) < 1e-4 * abs(gm.score(x))


def test_map_adapt_matches_oracle():
    k, d, n_spk = 32, 13, 6
    w, mu, var = synth.synth_ubm(k, d, seed=31)
    ubm = ssp.GaussianMixture.from_params(w, mu, var)
    spk_mu = synth.synth_speaker_means(mu, n_spk, seed=32, shift=0.5)
    frames = [synth.sample_gmm(w, spk_mu[i], var, 300 + 50 * i, seed=40 + i) for i in range(n_spk)]
    for adapt in (("means",), ("means", "weights", "variances")):
        ow, omu, ovar = (t.cpu().numpy() for t in ssp.map_adapt(ubm, frames, relevance=16.0, adapt=adapt))
        for i in range(n_spk):
            n, f, s, _ = ogmm.suff_stats(frames[i].astype(np.float64), w, mu, var)
            rw, rmu, rvar = ogmm.map_adapt(n, f, s, w, mu, var, len(frames[i]), 16.0, adapt)
            np.testing.assert_allclose(omu[i], rmu, atol=2e-4)
            np.testing.assert_allclose(ow[i], rw, atol=1e-5)
            np.testing.assert_allclose(ovar[i], rvar, rtol=1e-3, atol=1e-3)


def test_reference_pipeline_decisions(golden, tmp_path, monkeypatch):
    """GMM_UBM.GMM(model=True): models trained by the unmodified reference, scored on the GPU:
    the pred matrix matches GMM_UBM.py:194 and every argmax decision is identical."""
    from sklearn.mixture import GaussianMixture as SkGM

    g = golden("pipeline.npz")

    def sk(wt, m, v):
        e = SkGM(n_components=len(wt), covariance_type="diag")
        e.weights_, e.means_, e.covariances_ = wt, m, v
        e.precisions_cholesky_ = 1 / np.sqrt(v)
        return e

    monkeypatch.chdir(tmp_path)
    (tmp_path / "Model").mkdir()
    with open("Model/GMM_MFCC_model.pkl", "wb") as f:
        pickle.dump([sk(g["gmm_w"][i], g["gmm_mu"][i], g["gmm_var"][i]) for i in range(4)], f)
    with open("Model/UBM_MFCC_model.pkl", "wb") as f:
        pickle.dump(sk(g["ubm_w"], g["ubm_mu"], g["ubm_var"]), f)
    feats = [x.astype(np.float32) for x in g["feat_test"]]
    # "auto" (the default of ssp.GMM) is the 3-pass tensor rung at this model size: LLRs within the survey's 1e-3
    for precision, atol in (("fp32", 2e-4), ("auto", 2e-4), ("tf32x3", 2e-4), ("tf32x2", 1e-2), ("tf32", 2e-2)):
        acc_tr, acc, pred = ssp.GMM({}, feats, list(g["y_test"]), feats, list(g["y_test"]), n_components=4, model=True,
                                    precision=precision)
        np.testing.assert_allclose(pred, g["pred"], rtol=0, atol=atol)
        assert (pred.argmax(axis=1) == g["pred"].argmax(axis=1)).all()
        assert f"test acc {acc:.2%}" in str(g["printed"])


def test_GMM_trains_and_identifies(tmp_path, monkeypatch):
    """End to end on the GPU: audio -> extract_feature -> GMM() (train + identify), config 1 shape."""
    monkeypatch.chdir(tmp_path)
    x, y = synth.synth_corpus(4, 8, 16000)
    idx = np.random.RandomState(0).permutation(len(x))
    tr, te = idx[:22], idx[22:]
    train, f_tr, y_tr = ssp.extract_feature([x[i] for i in tr], [y[i] for i in tr], is_train=True)
    f_te, y_te = ssp.extract_feature([x[i] for i in te], [y[i] for i in te])
    acc_tr, acc, pred = ssp.GMM(train, f_tr, y_tr, f_te, y_te, n_components=4, random_state=0)
    assert acc_tr >= 0.9 and acc >= 0.6
    assert (tmp_path / "Model" / "GMM_MFCC_model.pkl").exists()
    with open(tmp_path / "Model" / "UBM_MFCC_model.pkl", "rb") as f:
        ubm = pickle.load(f)  # a stock sklearn estimator the reference GUIs can consume
    assert type(ubm).__module__.startswith("sklearn")
    assert np.isfinite(ubm.score(f_te[0].astype(np.float64)))


def test_main_from_wav_tree(tmp_path, monkeypatch):
    """GMM_UBM.main() equivalent: WAV tree -> load_data -> split -> extract_feature -> GMM()."""
    from scipy.io import wavfile

    from speech_signal_processing_b200 import ubm as subm

    root = tmp_path / "dataset" / "ASR_GMM"
    for s in range(3):
        d = root / f"spk{s}" / "session0"
        d.mkdir(parents=True)
        for u in range(7):
            sig = synth.synth_utterance(s, u, 16000)
            if u == 0:
                sig = np.stack([sig, sig], axis=1)  # a stereo file: first channel is used
            wavfile.write(str(d / f"utt{u}.wav"), 16000, sig)
    monkeypatch.chdir(tmp_path)
    subm.label_encoder.clear()
    x, y = ssp.load_data(str(root))
    assert len(x) == 21 and all(a.ndim == 1 and a.dtype == np.int16 for a in x) and sorted(set(y)) == [0, 1, 2]
    acc_tr, acc, pred = subm.main(str(root))
    assert pred.shape == (7, 3) and acc_tr >= 0.9
    subm.label_encoder.clear()


def test_full_size_identify_properties():
    """BASELINE config 4 at full size (10 000 utterances x 298 frames vs 1 000 speakers + UBM, K = 1024, D = 39)
    through size-independent properties: utterance permutation permutes rows, a sub-sample agrees with the FP32
    CUDA-core kernel and the float64 oracle, the planted speaker wins."""
    import torch

    k, d, s, n, t = 1024, 39, 1000, 10000, 298
    w, mu, var = synth.synth_ubm(k, d, seed=0)
    dev = torch.device("cuda")
    t_mu = torch.as_tensor(synth.synth_speaker_means(mu, s, seed=1, shift=0.25), device=dev)
    t_var = torch.as_tensor(var, device=dev)
    truth = torch.arange(n, device=dev) % s
    feats = synth.synth_features_torch(n * t, d, t_mu, t_var, truth.repeat_interleave(t), seed=3, device=dev)
    offs = np.arange(n + 1, dtype=np.int64) * t
    ms = ssp.ModelSet(torch.cat([torch.as_tensor(np.tile(w, (s, 1)), device=dev), torch.as_tensor(w, device=dev)[None]]),
                      torch.cat([t_mu, torch.as_tensor(mu, device=dev)[None]]),
                      torch.cat([t_var[None].expand(s, k, d), t_var[None]]))
    scores, _ = ms.score(feats, offs, precision="tf32")
    assert scores.shape == (n, s + 1) and bool(torch.isfinite(scores).all())
    llr = scores[:, :s] - scores[:, s:]
    assert bool((llr.argmax(dim=1) == truth).all())
    # permutation of whole utterances permutes the rows (bitwise up to the fp32 partial-sum order)
    perm = torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(0))
    feats_p = feats.view(n, t, d)[perm].reshape(n * t, d).contiguous()
    scores_p, _ = ms.score(feats_p, offs, precision="tf32")
    assert float((scores_p - scores[perm]).abs().max()) < 2e-5
    # sub-sample against the FP32 kernel and the oracle
    sub = [0, 1234, 9999]
    sub_feats = torch.cat([feats[i * t : (i + 1) * t] for i in sub])
    s32, _ = ms.score(sub_feats, np.arange(len(sub) + 1) * t, precision="fp32")
    rel = ((scores[sub] - s32).abs() / s32.abs()).max()
    assert float(rel) < 1e-4
    host_mu = t_mu.cpu().numpy()
    for r, i in enumerate(sub):
        x = feats[i * t : (i + 1) * t].cpu().numpy()
        for m in (int(truth[i]), 7):
            ref = ogmm.score(x, w, host_mu[m], var)
            assert abs(float(s32[r, m]) - ref) <= 2e-6 * abs(ref)
            assert abs(float(scores[i, m]) - ref) <= 1e-4 * abs(ref)
    # the shared-variance kernel (the bench's scorer) on the same full-size job: same scores, same decisions
    sms = ssp.SharedModelSet(w, t_var, torch.cat([t_mu, torch.as_tensor(mu, device=dev)[None]]), ref_model=s)
    sv, _ = sms.score(feats, offs)
    assert bool(torch.isfinite(sv).all())
    assert float(((sv - scores).abs() / scores.abs()).max()) < 1e-4      # the general kernel's single TF32 pass
    assert bool(((sv[:, :s] - sv[:, s:]).argmax(dim=1) == truth).all())
    assert float(((sv[sub] - s32).abs() / s32.abs()).max()) < 3e-5        # common part FP32-grade, differences in TF32
    llr32 = s32[:, :s] - s32[:, s:]
    assert float(((sv[sub][:, :s] - sv[sub][:, s:]) - llr32).abs().max()) < LLR_ATOL


# ------------------------------------------------------------------------------------------------
# shared-variance tensor kernel (mean-only MAP speaker sets): ssp_gmm_pack_shared / ssp_gmm_score_shared
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,d,n_spk", [(64, 39, 5), (100, 13, 3), (200, 26, 7), (1024, 39, 9)])
def test_shared_variance_scoring_matches_oracle(k, d, n_spk):
    import torch

    w, mu, var = synth.synth_ubm(k, d, seed=11)
    spk_mu = np.concatenate([synth.synth_speaker_means(mu, n_spk, seed=12, shift=0.3), mu[None]])  # last = UBM
    lens = [1, 2, 31, 32, 33, 127, 128, 129, 255, 256, 257, 298, 500]
    utts = [synth.sample_gmm(w, spk_mu[i % n_spk], var, n, seed=200 + i) for i, n in enumerate(lens)]
    sms = ssp.SharedModelSet(w, var, spk_mu)
    feats, offs = ssp.mixture.concat_utterances(utts, sms.device)
    got, lse = sms.score(feats, offs, want_frame_lse=True)
    got, lse = got.cpu().numpy(), lse.cpu().numpy()
    want = np.array([[ogmm.score(u, w, m, var) for m in spk_mu] for u in utts])
    long_enough = np.array(lens) >= 31
    # the common part is FP32 grade, only the speakers' differences from the reference carry 11-bit rounding
    np.testing.assert_allclose(got, want, rtol=REL["sv"], atol=0)
    np.testing.assert_allclose(got[long_enough], want[long_enough], rtol=SCORE_RTOL, atol=0)
    np.testing.assert_allclose(got[long_enough, :n_spk] - got[long_enough, n_spk:], want[long_enough, :n_spk] - want[long_enough, n_spk:],
                               rtol=0, atol=LLR_ATOL)
    np.testing.assert_allclose(got[:, n_spk], want[:, n_spk], rtol=5e-6, atol=0)   # the reference member itself
    assert (got[:, :n_spk].argmax(axis=1) == want[:, :n_spk].argmax(axis=1))[long_enough].all()
    # per-frame log-likelihoods (GaussianMixture.score_samples) of every model
    x = np.concatenate(utts)
    for i in (0, n_spk):
        ref = ogmm.score_samples(x, w, spk_mu[i], var)
        np.testing.assert_allclose(lse[i], ref, rtol=2e-3 if i < n_spk else 2e-5, atol=0)
    # and it is the same thing as the general (FP32) kernel on the expanded set
    gen, _ = sms.expand().score(feats, offs, precision="fp32")
    np.testing.assert_allclose(got[long_enough], gen.cpu().numpy()[long_enough], rtol=SCORE_RTOL, atol=0)
    torch.cuda.synchronize()


def test_shared_variance_scoring_rescues_models_far_from_the_reference():
    """A mean set nowhere near the reference model underflows the fixed per-frame stabiliser; the kernel marks the
    pair and the FP32 fix-up pass re-scores it, so the result is still right."""
    k, d = 64, 39
    w, mu, var = synth.synth_ubm(k, d, seed=21)
    far = mu + 40.0  # ~ -30000 nats per frame below the reference
    means = np.stack([mu + 0.1, far, mu])
    utts = [synth.sample_gmm(w, mu, var, n, seed=300 + n) for n in (64, 298)]
    sms = ssp.SharedModelSet(w, var, means)
    feats, offs = ssp.mixture.concat_utterances(utts, sms.device)
    got = sms.score(feats, offs)[0].cpu().numpy()
    want = np.array([[ogmm.score(u, w, m, var) for m in means] for u in utts])
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, want, rtol=REL["tf32"], atol=0)


def test_shared_variance_fp16_range_is_guarded():
    """The shared-variance kernel's operands are FP16.  Frames outside its range (|x| > 255: x^2 overflows) come back
    right through the FP32 fix-up pass; a model set outside it is refused by ssp_gmm_pack_shared and identify() takes
    the general kernel instead."""
    k, d = 64, 13
    w, mu, var = synth.synth_ubm(k, d, seed=23)
    means = np.stack([mu + 0.1, mu - 0.2, mu])
    utts = [synth.sample_gmm(w, mu, var, n, seed=310 + n) for n in (64, 130)]
    utts[1][5, 3] = 700.0    # one wild frame: its whole utterance is re-scored
    sms = ssp.SharedModelSet(w, var, means)
    feats, offs = ssp.mixture.concat_utterances(utts, sms.device)
    got = sms.score(feats, offs)[0].cpu().numpy()
    want = np.array([[ogmm.score(u, w, m, var) for m in means] for u in utts])
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, want, rtol=2e-5, atol=0)
    tiny = var * 1e-7        # (mu_s - mu_ref) / var ~ 1e6: not representable
    with pytest.raises(NotImplementedError):
        ssp.SharedModelSet(w, tiny, means)
    models = [ssp.GaussianMixture.from_params(w, m, tiny) for m in means]
    near = [m[:40] + 1e-5 for m in (mu, mu)]
    pred, who = ssp.identify(near, models[:2], models[2], precision="tf32")
    ref = np.array([[ogmm.score(u, w, m, tiny) - ogmm.score(u, w, mu, tiny) for m in means[:2]] for u in near])
    assert (who == ref.argmax(axis=1)).all()


def test_single_pass_rung_runs_from_fp16_images_and_guards_their_range():
    """precision="tf32" (one pass, 11-bit significands) streams the FP16 images of the pack: same tolerance as the TF32 images
    it replaces; a frame outside FP16's range is re-scored in FP32 by tc_fixup_kernel, a model set outside it falls back to
    the TF32 images."""
    from speech_signal_processing_b200 import _lib

    k, d = 200, 26
    w, mu, var = synth.synth_ubm(k, d, seed=33)
    spk = np.concatenate([synth.synth_speaker_means(mu, 3, seed=34, shift=0.3), mu[None]])
    utts = [synth.sample_gmm(w, spk[i % 3], var, n, seed=320 + i) for i, n in enumerate((298, 64, 130))]
    want = np.array([[ogmm.score(u, w, m, var) for m in spk] for u in utts])
    ms = ssp.ModelSet(np.tile(w, (4, 1)), spk, np.tile(var, (4, 1, 1)))
    feats, offs = ssp.mixture.concat_utterances(utts, ms.device)
    _lib.load().ssp_reset_launch_count()
    got = ms.score(feats, offs, precision="tf32")[0].cpu().numpy()
    assert _lib.launch_log() == {"gmm_score_tc_kernel": 1, "tc_fixup_kernel": 1}
    np.testing.assert_allclose(got, want, rtol=REL["tf32"], atol=0)
    wild = [u.copy() for u in utts]
    wild[1][7, 2] = -900.0
    want_w = np.array([[ogmm.score(u, w, m, var) for m in spk] for u in wild])
    f2, o2 = ssp.mixture.concat_utterances(wild, ms.device)
    got_w = ms.score(f2, o2, precision="tf32")[0].cpu().numpy()
    assert np.isfinite(got_w).all()
    np.testing.assert_allclose(got_w[1], want_w[1], rtol=5e-6, atol=0)        # the FP32 fix-up pass
    np.testing.assert_allclose(got_w[[0, 2]], want_w[[0, 2]], rtol=REL["tf32"], atol=0)
    tiny = var * 1e-7      # mu / var ~ 1e7: the pack kernel flags it, the TF32 images serve the call
    ms_t = ssp.ModelSet(np.tile(w, (4, 1)), spk, np.tile(tiny, (4, 1, 1)))
    near = [spk[0][:50] + 1e-5, spk[3][:70] - 1e-5]
    f3, o3 = ssp.mixture.concat_utterances(near, ms_t.device)
    _lib.load().ssp_reset_launch_count()
    got_t = ms_t.score(f3, o3, precision="tf32")[0].cpu().numpy()
    assert _lib.launch_log() == {"gmm_score_tc_kernel": 1}
    want_t = np.array([[ogmm.score(u, w, m, tiny) for m in spk] for u in near])
    assert (got_t.argmax(axis=1) == want_t.argmax(axis=1)).all()


def test_shared_variance_many_units_and_models():
    """More frames than one wave of 256-frame units x more models than accumulator slots; decisions equal FP32."""
    k, d, n_spk = 256, 39, 33
    w, mu, var = synth.synth_ubm(k, d, seed=31)
    spk_mu = np.concatenate([synth.synth_speaker_means(mu, n_spk, seed=32, shift=0.25), mu[None]])
    import torch

    dev = torch.device("cuda")
    n_utts, t = 400, 298  # 119 200 frames = 466 units > 148 SMs
    labels = torch.arange(n_utts, device=dev).repeat_interleave(t) % n_spk
    feats = synth.synth_features_torch(n_utts * t, d, torch.as_tensor(spk_mu[:n_spk], device=dev), torch.as_tensor(var, device=dev),
                                       labels, seed=5, device=dev)
    offs = np.arange(n_utts + 1, dtype=np.int64) * t
    sms = ssp.SharedModelSet(w, var, spk_mu)
    a = sms.score(feats, offs)[0].cpu().numpy()
    b = sms.expand().score(feats, offs, precision="fp32")[0].cpu().numpy()
    assert np.abs(a - b).max() <= 1e-4 * np.abs(b).max()
    assert (a[:, :n_spk].argmax(axis=1) == b[:, :n_spk].argmax(axis=1)).all()
    a2 = sms.score(feats, offs)[0].cpu().numpy()  # the workspace is left clean: a second call gives the same answer
    assert np.abs(a - a2).max() <= 1e-9 * np.abs(a).max()


def test_map_enrol_identify_equals_general_path():
    """map_enrol -> identify (one shared-variance launch) == map_adapt -> ModelSet -> identify with a separate UBM."""
    k, d, n_spk = 128, 39, 12
    w, mu, var = synth.synth_ubm(k, d, seed=41)
    spk_mu = synth.synth_speaker_means(mu, n_spk, seed=42, shift=0.4)
    ubm = ssp.GaussianMixture.from_params(w, mu, var)
    enrol = [synth.sample_gmm(w, spk_mu[i], var, 900, seed=500 + i) for i in range(n_spk)]
    tests = [synth.sample_gmm(w, spk_mu[i % n_spk], var, 298, seed=600 + i) for i in range(3 * n_spk)]
    sms = ssp.map_enrol(ubm, enrol, relevance=16.0)
    assert sms.n_models == n_spk + 1 and sms.ubm_index == n_spk
    pred_sv, who_sv = ssp.identify(tests, sms)
    aw, amu, avar = ssp.map_adapt(ubm, enrol, relevance=16.0)
    pred, who = ssp.identify(tests, ssp.ModelSet(aw, amu, avar), ubm, precision="fp32")
    assert pred_sv.shape == pred.shape == (len(tests), n_spk)
    # the shared-variance kernel rounds only the speakers' DIFFERENCES from the UBM: LLRs within the survey's 1e-3
    np.testing.assert_allclose(pred_sv, pred, rtol=0, atol=LLR_ATOL)
    assert (who_sv == who).all() and (who == np.arange(len(tests)) % n_spk).all()
    # the general tensor kernel in one TF32 pass rounds the full means of every model: 1e-4 relative on |score| ~ 55
    pred_tc, who_tc = ssp.identify(tests, ssp.ModelSet(aw, amu, avar), ubm, precision="tf32")
    np.testing.assert_allclose(pred_tc, pred, rtol=0, atol=1.2e-2)
    assert (who_sv == who_tc).all()


def test_identify_routes_shared_base_model_lists_to_the_shared_variance_kernel():
    """A list of models that all carry the UBM's weights and covariances (mean-only MAP, e.g. unpickled) goes through
    ssp_gmm_score_shared; the result equals the general path's."""
    from speech_signal_processing_b200 import _lib

    k, d, n_spk = 64, 26, 6
    w, mu, var = synth.synth_ubm(k, d, seed=51)
    spk_mu = synth.synth_speaker_means(mu, n_spk, seed=52, shift=0.4)
    ubm = ssp.GaussianMixture.from_params(w, mu, var)
    models = [ssp.GaussianMixture.from_params(w, spk_mu[i], var) for i in range(n_spk)]
    tests = [synth.sample_gmm(w, spk_mu[i % n_spk], var, 200 + 11 * i, seed=700 + i) for i in range(2 * n_spk)]
    _lib.load().ssp_reset_launch_count()
    pred, who = ssp.identify(tests, models, ubm, precision="tf32")
    assert _lib.launch_log() == {"gmm_pack_sv_kernel": 1, "gmm_score_sv_kernel": 1, "sv_fixup_kernel": 1}
    ref, who_ref = ssp.identify(tests, models, ubm, precision="fp32")
    np.testing.assert_allclose(pred, ref, rtol=0, atol=2e-3)     # (the general kernel's single TF32 pass: 1.2e-2, below)
    assert (who == who_ref).all() and (who == np.arange(len(tests)) % n_spk).all()
    # a model with its own variances switches the whole call to the general kernel
    odd = ssp.GaussianMixture.from_params(w, spk_mu[0], var * 1.1)
    pred2, _ = ssp.identify(tests, models[:-1] + [odd], ubm, precision="tf32")
    np.testing.assert_allclose(pred2[:, :-1], ref[:, :-1], rtol=0, atol=1.2e-2)
    # default precision at K = 64: the 3-pass general tensor kernel for speakers and UBM, LLRs at the survey's bar
    _lib.load().ssp_reset_launch_count()
    pred3, who3 = ssp.identify(tests, models, ubm)
    assert _lib.launch_log() == {"gmm_pack_kernel": 2, "gmm_score_tc_kernel<3 passes>": 2, "tc_fixup_kernel": 2}   # (FP16 pieces + range net)
    np.testing.assert_allclose(pred3, ref, rtol=0, atol=LLR_ATOL)
    assert (who3 == who_ref).all()
    # a UBM handed over as a ModelSet: pred is still the LLR, whichever kernel is selected (speaker score - UBM score)
    ums = ssp.ModelSet(w, mu, var)
    for precision in ("auto", "tf32", "fp32"):
        pred4, who4 = ssp.identify(tests, models, ums, precision=precision)
        np.testing.assert_allclose(pred4, ref, rtol=0, atol=2e-3 if precision == "tf32" else LLR_ATOL)
        assert (who4 == who_ref).all()


@pytest.mark.parametrize("split", [False, True])
def test_identify_pcm_equals_separate_calls(split):
    """identify_pcm (pinned host PCM -> decisions on the host, H2D of the tail overlapped with the kernels of the head)
    == front-end + scoring + argmax called one after the other on the whole batch; also with ragged utterance lengths
    and the UBM in the middle of the model set."""
    import torch

    k, d, n_spk = 128, 39, 6
    w, mu, var = synth.synth_ubm(k, d, seed=3, spread=0.8)
    spk = synth.synth_speaker_means(mu, n_spk, seed=4, shift=0.3)
    means = np.concatenate([spk[:2], mu[None], spk[2:]])  # UBM is model 2
    sms = ssp.SharedModelSet(w, var, means, ref_model=2)
    sigs = [synth.synth_utterance(s % 5, s, 16000 + 800 * (s % 7)) for s in range(40)]
    fe = ssp.FrontEnd(ssp.sidekit_recipe(), delta_order=2, cmvn=True)
    host, offs = fe.pack_host(sigs)
    feats, foffs, _ = fe.extract(sigs)
    scores, _ = sms.score(feats, foffs)
    llr = (scores - scores[:, 2:3]).cpu().numpy()
    llr[:, 2] = -np.inf
    want = llr.argmax(axis=1)
    got, total = ssp.identify_pcm(host, offs, fe, sms, ubm_index=2, min_split_frames=0 if split else 1 << 40, head_fraction=0.3)
    assert total == int(foffs[-1]) and got.dtype == torch.int64 and got.shape == (len(sigs),)
    assert (got.numpy() == want).all()
    # numpy PCM (not pinned) and no UBM: plain argmax over the scores
    got2, _ = ssp.identify_pcm(host.numpy().copy(), offs, fe, sms, min_split_frames=0 if split else 1 << 40)
    assert (got2.numpy() == scores.cpu().numpy().argmax(axis=1)).all()


def test_config1_reference_pipeline_at_full_size(golden, config1_corpus, tmp_path, monkeypatch):
    """BASELINE configs[0] at its stated size (10 speakers x 30 utterances x 3 s, 26-d, 64 components): audio ->
    extract_feature on the GPU -> GMM(model=True) with the models the UNMODIFIED reference trained (fixture) -> the LLR
    matrix of GMM_UBM.py:191-194, every decision and the printed accuracy equal the reference's."""
    import zlib

    from sklearn.mixture import GaussianMixture as SkGM

    g = golden("config1.npz")
    _, x_te, _, y_te = config1_corpus
    assert zlib.crc32(np.concatenate(x_te).tobytes()) == int(g["audio_crc"])
    feats, labels = ssp.extract_feature(x_te, list(y_te))
    assert len(feats) == 90 and feats[0].shape == (298, 26)
    tol = lambda want: 1e-4 * np.maximum(1.0, np.abs(want))  # noqa: E731  
    # CMVN divides by a per-dimension std of ~1e-1..1e1: allow the front-end tolerance after that scaling
    assert (np.abs(feats[0] - g["feat_first"]) <= 5 * tol(g["feat_first"])).all()
    assert (np.abs(feats[-1] - g["feat_last"]) <= 5 * tol(g["feat_last"])).all()
    np.testing.assert_allclose([np.abs(f).mean() for f in feats], g["feat_abs_mean"], rtol=1e-4)

    def sk(wt, m, v):
        e = SkGM(n_components=len(wt), covariance_type="diag")
        e.weights_, e.means_, e.covariances_ = wt, m.astype(np.float64), v.astype(np.float64)
        e.precisions_cholesky_ = 1 / np.sqrt(e.covariances_)
        return e

    monkeypatch.chdir(tmp_path)
    (tmp_path / "Model").mkdir()
    with open("Model/GMM_MFCC_model.pkl", "wb") as f:
        pickle.dump([sk(g["gmm_w"][i], g["gmm_mu"][i], g["gmm_var"][i]) for i in range(10)], f)
    with open("Model/UBM_MFCC_model.pkl", "wb") as f:
        pickle.dump(sk(g["ubm_w"], g["ubm_mu"], g["ubm_var"]), f)
    # An LLR is the difference of two scores of magnitude 30 .. 100.  The default rung ("auto" -> 3-pass TF32 at K = 64)
    # and the FP32 kernel see only the front-end's float32 differences; the single-pass TF32 kernel is good to 5e-4
    # relative per score at this model size (measured: 4.4e-4 of the LLR at worst) and stays an opt-in.
    default_pred = None
    for precision, atol in ((None, 2e-3), ("fp32", 2e-3), ("tf32x3", 2e-3), ("tf32x2", 2e-2), ("tf32", 5e-2)):
        kw = {} if precision is None else {"precision": precision}
        acc_tr, acc, pred = ssp.GMM({}, feats, labels, feats, labels, n_components=64, model=True, **kw)
        np.testing.assert_allclose(pred, g["pred"], rtol=0, atol=atol)
        if precision is None:
            default_pred = pred
        elif precision == "fp32":  # tensor path vs CUDA-core path on the SAME features: the survey's LLR bar
            np.testing.assert_allclose(default_pred, pred, rtol=0, atol=LLR_ATOL)
        assert (pred.argmax(axis=1) == g["pred"].argmax(axis=1)).all()
        assert acc == 1.0 and "test acc 100.00%" in str(g["printed"])
    assert float(g["min_top2_margin"]) > 100 * 5e-2  # decisions are far from the tolerance


@pytest.mark.parametrize("feature_type", ["MFCC", "PLP", "MFCC_PLP"])
def test_chunked_gui_path_matches_oracle(feature_type):
    """records() + _GMM_test() of the final GUI (UI/tmp.py:301-349): 1-s chunks -> mfcc / plp / both side by side ->
    scale -> GMM[i].score - UBM.score -> the GUI's probability and argmax, batched."""
    from oracle import frontend as ofe

    audio = np.concatenate([synth.synth_utterance(s, 7, 16000 + 2200 * s) for s in range(3)])  # 3.4 s -> 3 chunks
    feats = ssp.chunk_features(audio, feature_type)
    assert len(feats) == len(audio) // 16000 == 3
    want = []
    for i in range(3):
        ch = audio[i * 16000 : (i + 1) * 16000]
        parts = []
        if "MFCC" in feature_type:
            parts.append(ofe.sidekit_mfcc(ch)[0])
        if "PLP" in feature_type:
            parts.append(ofe.sidekit_plp(ch)[0])
        want.append(ofe.scale(np.hstack(parts)))
    for f, w_ in zip(feats, want):
        assert f.shape == w_.shape == (98, 13 * (2 if feature_type == "MFCC_PLP" else 1))
        # scale divides by per-column std (PLP columns: ~0.02 .. 0.5)
        np.testing.assert_allclose(f, w_, rtol=0, atol=5e-4 if feature_type == "MFCC" else 2e-2)
    d = feats[0].shape[1]
    models = []
    for i in range(3):
        wt, mu, var = synth.synth_ubm(8, d, seed=70 + i, spread=0.7)
        models.append(ssp.GaussianMixture.from_params(wt, mu, var))
    wt, mu, var = synth.synth_ubm(8, d, seed=80, spread=0.7)
    ubm = ssp.GaussianMixture.from_params(wt, mu, var)
    pred, prob, who = ssp.chunk_identify(feats, models, ubm, precision="fp32")
    ref = np.array([[ogmm.score(f.astype(np.float64), m.weights_, m.means_, m.covariances_)
                     - ogmm.score(f.astype(np.float64), ubm.weights_, ubm.means_, ubm.covariances_) for m in models] for f in feats])
    np.testing.assert_allclose(pred, ref, rtol=0, atol=2e-4)
    np.testing.assert_allclose(prob, np.exp(ref.max(axis=1)) / np.exp(ref).sum(axis=1), rtol=1e-3)
    assert (who == ref.argmax(axis=1)).all()
    assert ssp.chunk_features(audio[:15999], feature_type) == []
    with pytest.raises(NameError):
        ssp.chunk_features(audio, "LPCC")


# ------------------------------------------------------------------------------------------------
# the named configs of BASELINE.json at their own model sizes (K = 2048 scoring, K = 512 statistics)
# ------------------------------------------------------------------------------------------------
def test_config5_2048_components_scoring_matches_oracle():
    """BASELINE configs[4]: K = 2048, D = 39 -- 16 component tiles per model in the general tensor kernel, 32 in the
    shared-variance one.  Utterance scores, per-frame log-likelihoods and decisions against the float64 oracle."""
    k, d, n_spk = 2048, 39, 4
    w, mu, var = synth.synth_ubm(k, d, seed=61)
    spk_mu = np.concatenate([synth.synth_speaker_means(mu, n_spk, seed=62, shift=0.25), mu[None]])  # last = UBM
    lens = [64, 298, 130, 257, 98, 2998 // 4]
    utts = [synth.sample_gmm(w, spk_mu[i % n_spk], var, n, seed=900 + i) for i, n in enumerate(lens)]
    x = np.concatenate(utts)
    want = np.array([[ogmm.score(u, w, m, var) for m in spk_mu] for u in utts])
    want_llr = want[:, :n_spk] - want[:, n_spk:]
    ms = ssp.ModelSet(np.tile(w, (n_spk + 1, 1)), spk_mu, np.tile(var, (n_spk + 1, 1, 1)))
    feats, offs = ssp.mixture.concat_utterances(utts, ms.device)
    for precision, rtol, llr_atol in (("tf32", SCORE_RTOL, 5e-3), ("tf32x3", 2e-6, 1e-4), ("fp32", 2e-6, 1e-4)):
        got, lse = ms.score(feats, offs, precision=precision, want_frame_lse=True)
        got, lse = got.cpu().numpy(), lse.cpu().numpy()
        np.testing.assert_allclose(got, want, rtol=rtol, atol=0)
        np.testing.assert_allclose(got[:, :n_spk] - got[:, n_spk:], want_llr, rtol=0, atol=llr_atol)
        assert ((got[:, :n_spk]).argmax(axis=1) == want[:, :n_spk].argmax(axis=1)).all()
        for i in (0, n_spk):
            np.testing.assert_allclose(lse[i], ogmm.score_samples(x, w, spk_mu[i], var), rtol=3e-3 if precision == "tf32" else 5e-5)
    sms = ssp.SharedModelSet(w, var, spk_mu)
    got, lse = sms.score(feats, offs, want_frame_lse=True)
    got, lse = got.cpu().numpy(), lse.cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=3e-5, atol=0)
    np.testing.assert_allclose(got[:, :n_spk] - got[:, n_spk:], want_llr, rtol=0, atol=LLR_ATOL)
    assert (got[:, :n_spk].argmax(axis=1) == want[:, :n_spk].argmax(axis=1)).all()
    for i in (1, n_spk):
        np.testing.assert_allclose(lse[i], ogmm.score_samples(x, w, spk_mu[i], var), rtol=1e-3 if i < n_spk else 2e-5)


def test_config3_512_component_statistics_match_oracle():
    """BASELINE configs[2] at its model size: N / F / S and the log-likelihood under a K = 512, D = 39 UBM over 240 000
    frames in 4 segments (one empty) -- thousands of frame blocks accumulated in TMEM per CTA -- against the float64
    oracle evaluated chunk by chunk."""
    import torch

    from speech_signal_processing_b200 import _lib

    k, d = 512, 39
    w, mu, var = synth.synth_ubm(k, d, seed=71, spread=1.0)
    lens = [150_000, 0, 60_001, 29_999]
    x = synth.sample_gmm(w, mu * 0.9, var * 1.1, sum(lens), seed=72)
    seg = np.concatenate([[0], np.cumsum(lens)])
    ms = ssp.ModelSet(w, mu, var)
    _lib.load().ssp_reset_launch_count()
    n, f, s, ll = (t.cpu().numpy() for t in ms.stats(torch.as_tensor(x, device="cuda"), seg))
    assert set(_lib.launch_log()) == EM_TENSOR_KERNELS, _lib.launch_log()
    for i, ln in enumerate(lens):
        if ln == 0:
            assert np.all(n[i] == 0) and np.all(f[i] == 0) and np.all(s[i] == 0) and ll[i] == 0
            continue
        rn, rf, rs_, rll = np.zeros(k), np.zeros((k, d)), np.zeros((k, d)), 0.0
        for lo in range(seg[i], seg[i + 1], 20_000):
            part = ogmm.suff_stats(x[lo : min(lo + 20_000, seg[i + 1])].astype(np.float64), w, mu, var)
            rn += part[0]; rf += part[1]; rs_ += part[2]; rll += part[3]
        # gamma is TF32-rounded (2^-11 relative, unbiased) before the statistics GEMM: the error of a sum over n_c
        # frames grows like sqrt(n_c), far below these bounds at n_c ~ ln / k
        np.testing.assert_allclose(n[i], rn, rtol=1e-4, atol=1e-4 * ln / k)
        np.testing.assert_allclose(f[i], rf, rtol=1e-4, atol=3e-4 * ln / k)
        np.testing.assert_allclose(s[i], rs_, rtol=1e-4, atol=6e-4 * ln / k)
        assert abs(ll[i] - rll) <= 1e-5 * abs(rll)
        assert abs(n[i].sum() - ln) < 1e-5 * ln
    # the M-step the EM loop would take from these statistics equals the oracle's
    tot = [a.sum(axis=0) for a in (n, f, s)]
    rw, rmu, rvar = ogmm.m_step(*(np.asarray(t, dtype=np.float64) for t in tot))
    gw, gmu, gvar = (torch.empty(sh, dtype=torch.float64, device="cuda") for sh in ((k,), (k, d), (k, d)))
    tn, tf_, ts = (torch.as_tensor(t, device="cuda") for t in tot)
    _lib.check(_lib.load().ssp_gmm_mstep(_lib.ptr(tn), _lib.ptr(tf_), _lib.ptr(ts), 1, k, d, 1e-6, 10 * np.finfo(np.float64).eps,
                                         _lib.ptr(gw), _lib.ptr(gmu), _lib.ptr(gvar), _lib.stream_ptr()), "ssp_gmm_mstep")
    np.testing.assert_allclose(gw.cpu().numpy(), rw, rtol=1e-12)
    np.testing.assert_allclose(gmu.cpu().numpy(), rmu, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gvar.cpu().numpy(), rvar, rtol=1e-9)


def test_install_runs_the_reference_call_pattern(tmp_path, monkeypatch):
    """a stand-in module with the names GMM_UBM.py:16-20,53 binds, ssp.install() on it, then
    the reference's own extract_feature -> GMM call pattern (GMM_UBM.py:86-93, :154-170, :182-197) through those names."""
    import types

    from oracle import frontend as ofe

    mod = types.ModuleType("GMM_UBM")
    for name in ("GaussianMixture", "preprocessing", "mfcc", "plp", "delta"):
        setattr(mod, name, None)
    assert ssp.install(mod) is mod
    x, y = synth.synth_corpus(3, 6, 16000)
    feats = []
    for sig in x:                                        # GMM_UBM.py:86-93
        c = mod.mfcc(sig)
        assert isinstance(c, np.ndarray) and c.shape == (98, 13)
        dl = mod.delta(c, 2)
        feats.append(mod.preprocessing.scale(np.hstack((c, dl))))
    want = ofe.features(x[0], preset="sidekit", delta_order=1, cmvn=True)
    np.testing.assert_allclose(feats[0], want, rtol=0, atol=5e-4)
    p = mod.plp(x[0])
    assert p.shape[0] == 98 and np.isfinite(p).all()
    train = {s: np.vstack([f for f, lab in zip(feats, y) if lab == s]) for s in range(3)}
    gmms = [mod.GaussianMixture(n_components=4, covariance_type="diag", random_state=0).fit(train[s]) for s in range(3)]
    ubm = mod.GaussianMixture(n_components=4, covariance_type="diag", random_state=0).fit(np.vstack([train[s] for s in range(3)]))
    pred = np.array([[gmms[i].score(f) - ubm.score(f) for i in range(3)] for f in feats])   # GMM_UBM.py:182-185
    assert (pred.argmax(axis=1) == np.array(y)).mean() >= 0.9
    # the same numbers as the batched entry point
    ref, _ = ssp.identify(feats, gmms, ubm, precision="fp32")
    np.testing.assert_allclose(pred, ref, rtol=0, atol=1e-4)


def test_shared_variance_model_groups_in_l2(tmp_path):
    """Large speaker sets are scored in L2-resident model groups (the per-frame stabilisers of the first group are kept in
    the workspace for the later ones).  Forced here with a tiny group footprint (SSP_SV_GROUP_MB is read once per process):
    70 models in 3 groups over 5 frame units give the scores of the single-group order and of the FP32 kernel."""
    import subprocess
    import sys

    script = tmp_path / "groups.py"
    script.write_text(r"""
import os, sys
import numpy as np, torch
sys.path.insert(0, %r)
import speech_signal_processing_b200 as ssp
from speech_signal_processing_b200 import synth
k, d, n_spk = 64, 39, 69
w, mu, var = synth.synth_ubm(k, d, seed=31)
spk_mu = np.concatenate([synth.synth_speaker_means(mu, n_spk, seed=32, shift=0.25), mu[None]])
lens = [298, 1, 300, 255, 257, 129]
utts = [synth.sample_gmm(w, spk_mu[i %% n_spk], var, n, seed=5 + i) for i, n in enumerate(lens)]
sms = ssp.SharedModelSet(w, var, spk_mu, ref_model=n_spk)
feats, offs = ssp.mixture.concat_utterances(utts, sms.device)
got, lse = sms.score(feats, offs, want_frame_lse=True)
assert (sms._ws is not None) == (os.environ["SSP_SV_GROUP_MB"] != "0"), "workspace <=> more than one group"
ref = sms.expand().score(feats, offs, precision="fp32", want_frame_lse=True)
np.save(sys.argv[1], np.stack([got.cpu().numpy(), ref[0].cpu().numpy()]))
assert float((lse - ref[1]).abs().max() / ref[1].abs().max()) < 3e-3
""" % ROOT)
    out = {}
    for mb in ("0.4", "0"):
        path = str(tmp_path / f"scores_{mb}.npy")
        r = subprocess.run([sys.executable, str(script), path], env=dict(os.environ, SSP_SV_GROUP_MB=mb), capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        out[mb] = np.load(path)
    grouped, fp32 = out["0.4"]
    single = out["0"][0]
    long_enough = np.array([298, 1, 300, 255, 257, 129]) >= 31
    np.testing.assert_allclose(grouped[long_enough], fp32[long_enough], rtol=REL["tf32"], atol=0)
    # same stabilisers, same tiles, same order of the partial sums within a model: the grouping changes nothing
    np.testing.assert_array_equal(grouped, single)


def test_fit_batch_equals_the_per_speaker_loop():
    """GMM_UBM.py:154-165 trains one GMM per speaker in a Python loop; fit_batch runs the EM iterations of all of them
    as one segmented call per iteration (segment s under model s).  Same initial parameters in => per-model n_iter_,
    convergence flags, lower bounds and parameters of the looped fits out -- also when the models need different numbers
    of iterations -- with O(iterations) statistics launches instead of O(speakers x iterations)."""
    import warnings

    from speech_signal_processing_b200 import _lib

    k, d, n_spk = 16, 26, 5
    xs, inits = [], []
    for i in range(n_spk):
        w, mu, var = synth.synth_ubm(k, d, seed=90 + i, spread=1.0 + 0.3 * i)
        xs.append(synth.sample_gmm(w, mu, var, 1500 + 211 * i, seed=95 + i))       # ragged: 1500 .. 2344 frames (odd block counts)
        w0, mu0, var0 = synth.synth_ubm(k, d, seed=190 + i, spread=1.0)
        inits.append((w0, mu0 + 0.2, np.ones_like(var0)))
    kw = dict(covariance_type="diag", max_iter=40, tol=1e-3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        looped = [ssp.GaussianMixture(n_components=k, weights_init=a, means_init=b, precisions_init=1.0 / c, **kw).fit(x)
                  for x, (a, b, c) in zip(xs, inits)]
        _lib.load().ssp_reset_launch_count()
        batched = ssp.fit_batch(xs, n_components=k, weights_init=[a for a, _, _ in inits], means_init=[b for _, b, _ in inits],
                                precisions_init=[1.0 / c for _, _, c in inits], **kw)
    log = _lib.launch_log()
    iters = max(g.n_iter_ for g in batched)
    assert log["gmm_em_stats_kernel"] == iters and log["gmm_mstep_kernel"] == iters   # not n_spk x iterations
    assert len({g.n_iter_ for g in looped}) > 1, "the case should exercise per-model convergence"
    for a, b in zip(looped, batched):
        assert (a.n_iter_, a.converged_) == (b.n_iter_, b.converged_)
        # The two runs chunk the frames differently (FP32 accumulation in TMEM per chunk) and add the float64 partial
        # statistics in a different order; over dozens of iterations that moves the trajectory by ~1e-5 relative.
        assert abs(a.lower_bound_ - b.lower_bound_) <= 2e-5 * abs(a.lower_bound_)
        np.testing.assert_allclose(b.weights_, a.weights_, rtol=2e-3, atol=1e-6)
        np.testing.assert_allclose(b.means_, a.means_, rtol=0, atol=2e-3)
        np.testing.assert_allclose(b.covariances_, a.covariances_, rtol=1e-2, atol=1e-5)
    # ... while ONE iteration from the same parameters is the same arithmetic up to the summation order
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        one_l = [ssp.GaussianMixture(n_components=k, weights_init=a, means_init=b, precisions_init=1.0 / c, covariance_type="diag",
                                     max_iter=1).fit(x) for x, (a, b, c) in zip(xs, inits)]
        one_b = ssp.fit_batch(xs, n_components=k, weights_init=[a for a, _, _ in inits], means_init=[b for _, b, _ in inits],
                              precisions_init=[1.0 / c for _, _, c in inits], covariance_type="diag", max_iter=1)
    for a, b in zip(one_l, one_b):
        assert abs(a.lower_bound_ - b.lower_bound_) <= 1e-7 * abs(a.lower_bound_)
        np.testing.assert_allclose(b.weights_, a.weights_, rtol=1e-6, atol=1e-9)
        np.testing.assert_allclose(b.means_, a.means_, rtol=0, atol=1e-6)
        np.testing.assert_allclose(b.covariances_, a.covariances_, rtol=1e-5, atol=1e-8)
    # default initialisation (k-means per model, then batched EM) trains usable models
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        auto = ssp.fit_batch(xs, n_components=k, covariance_type="diag", random_state=0)
    for i, g in enumerate(auto):
        assert g.score(xs[i]) > max(g.score(xs[j]) for j in range(n_spk) if j != i)

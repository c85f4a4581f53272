"""Parity of the VAD kernels (through the C ABI) with the unmodified reference's outputs (tests/golden/vad.npz)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import vad as ov  # noqa: E402
from speech_signal_processing_b200 import vad  # noqa: E402

# Stated tolerances: zero-crossing counts and decisions exact; energy 1e-6 relative (double accumulation of
# float32-rounded samples on the frame-matrix path, exact samples on the PCM path); spectral entropy 2e-4 absolute
# (float32 FFT against numpy's float64 one).
ENT_ATOL = 2e-4


def _cases(g):
    return range(int(g["n_cases"]))


def test_vad_batch_matches_reference(golden):
    g = golden("vad.npz")
    sigs = [g[f"sig{i}"] for i in _cases(g)]
    speech, foffs, (zg, p, e) = vad.vad_batch(sigs)
    speech, zg, p, e = speech.cpu().numpy(), zg.cpu().numpy(), p.cpu().numpy(), e.cpu().numpy()
    for i in _cases(g):
        sl = slice(foffs[i], foffs[i + 1])
        assert foffs[i + 1] - foffs[i] == len(g[f"zcr{i}"])                      # ceil(N / 128) frames
        assert np.array_equal(zg[sl], g[f"zcr{i}"][:, 0])
        np.testing.assert_allclose(p[sl], g[f"power{i}"][:, 0], rtol=1e-12)
        np.testing.assert_allclose(e[sl], g[f"entropy{i}"][:, 0], rtol=0, atol=ENT_ATOL)
        assert np.array_equal(speech[sl], g[f"det{i}"][:, 0].astype(np.uint8))
        clear = np.abs(g[f"entropy{i}"][:, 0] - 0.4) > 10 * ENT_ATOL
        assert np.array_equal(vad.VAD_frequency(e[sl])[clear], g[f"freq{i}"][:, 0][clear])
    # other thresholds, same launches
    speech_b, _, _ = vad.vad_batch(sigs, zcr_gate=25, ampl=1.0, amph=8)
    speech_b = speech_b.cpu().numpy()
    for i in _cases(g):
        assert np.array_equal(speech_b[foffs[i] : foffs[i + 1]], g[f"det_b{i}"][:, 0].astype(np.uint8))


def test_vad_reference_entry_points(golden):
    """enframe / ZCR / energy / spectrum_entropy / feature / VAD_detection with the reference's shapes."""
    g = golden("vad.npz")
    for i in _cases(g):
        wave = ov.wav_normalise(g[f"sig{i}"])
        frames = vad.enframe(wave)
        if f"frames{i}" in g:
            assert np.array_equal(frames, g[f"frames{i}"])
        n = frames.shape[1]
        z = vad.ZCR(frames)
        assert z.shape == (n, 1) and np.array_equal(z, g[f"zcr_raw{i}"])
        np.testing.assert_allclose(vad.energy(frames), g[f"power{i}"], rtol=1e-6)
        np.testing.assert_allclose(vad.spectrum_entropy(frames), g[f"entropy{i}"], rtol=0, atol=ENT_ATOL)
        zg, p, e = vad.feature(frames)
        assert np.array_equal(zg, g[f"zcr{i}"])
        np.testing.assert_allclose(p, g[f"power{i}"], rtol=1e-6)
        assert np.array_equal(vad.VAD_detection(g[f"zcr{i}"], g[f"power{i}"]), g[f"det{i}"])
        assert np.array_equal(vad.VAD_detection(g[f"zcr{i}"], g[f"power{i}"], zcr_gate=25, ampl=1.0, amph=8), g[f"det_b{i}"])


def test_vad_detector_quirks_match_oracle():
    """Random feature tracks (many runs, runs at both ends, wrapped backward search) against the oracle's restatement."""
    rs = np.random.RandomState(0)
    for trial in range(40):
        n = int(rs.randint(1, 300))
        power = np.abs(rs.standard_normal(n)) * rs.choice([0.05, 1.0, 20.0], size=n, p=[0.3, 0.3, 0.4])
        run = rs.randint(0, max(1, n - 1))
        power[run : run + rs.randint(0, 60)] = 30.0
        if trial % 5 == 0:
            power[0] = 0.01   # the backward search always finds a quiet frame before wrapping past -n
        zcr = rs.randint(0, 80, size=n).astype(np.float64) * (power > 0.1)
        try:
            want = ov.detect(zcr, power)
        except IndexError:
            continue           # the reference itself raises here
        got = vad.VAD_detection(zcr, power)
        assert np.array_equal(got, want), trial

"""Deviation of the EM trajectory from sklearn's (fixtures of tests/golden/sklearn_gmm.npz) per iteration count: sets the
tolerances of test_em_trajectory_matches_sklearn.   gpurun -- 'python tests/em_trajectory_probe.py'"""
import json
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import speech_signal_processing_b200 as ssp  # noqa: E402

g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sklearn_gmm.npz"))
for tag in ("s", "m"):
    w, mu, var, x = g[f"{tag}_w"], g[f"{tag}_mu"], g[f"{tag}_var"], g[f"{tag}_x"]
    for iters in (1, 3, 100):
        gm = ssp.GaussianMixture(n_components=len(w), covariance_type="diag", weights_init=w, means_init=mu, precisions_init=1.0 / var,
                                 max_iter=iters)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gm.fit(x)
        rw, rmu, rvar = g[f"{tag}_fit{iters}_w"], g[f"{tag}_fit{iters}_mu"], g[f"{tag}_fit{iters}_var"]
        print(json.dumps({"tag": tag, "K": len(w), "D": int(x.shape[1]), "frames": int(len(x)), "iters": iters, "n_iter": int(gm.n_iter_),
                          "lb_rel": abs(gm.lower_bound_ / float(g[f"{tag}_fit{iters}_lb"]) - 1),
                          "w_abs": float(np.abs(gm.weights_ - rw).max()), "w_rel": float(np.abs(gm.weights_ / rw - 1).max()),
                          "mu_abs": float(np.abs(gm.means_ - rmu).max()), "var_rel": float(np.abs(gm.covariances_ / rvar - 1).max())}))

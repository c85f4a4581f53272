"""CPU oracle for the MFCC -> diag-GMM-UBM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``speech_signal_processing_b200/``
imports this package.  The only allowed importers are ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs, and there only as the checker / CPU baseline.

What is restated, and how each piece is pinned
----------------------------------------------
* ``oracle.frontend.processing_mfcc``  restates /root/reference/utils/processing.py:19-144
  (enframe, mfccInitFilterBanks, stMFCC, MFCC).  PINNED: the reference file is
  imported unchanged (with import shims) by ``tests/golden/make_golden.py`` and
  its outputs are committed under ``tests/golden/``; ``tests/test_oracle.py``
  checks the restatement against them.
* ``oracle.frontend.delta`` restates /root/reference/GMM_UBM.py:53-69.  PINNED the
  same way (reference ``delta`` imported and run).
* ``oracle.frontend.scale`` restates sklearn.preprocessing.scale as called at
  /root/reference/GMM_UBM.py:93.  PINNED against sklearn 1.9.0 outputs.
* ``oracle.gmm`` restates sklearn.mixture.GaussianMixture (diag) E-step, M-step,
  score and the EM loop as called at /root/reference/GMM_UBM.py:158-170,185,194.
  PINNED against sklearn 1.9.0 (golden fixtures + live comparison in the tests).
* ``oracle.vad`` restates /root/reference/VAD.py (framing, zero-crossing count, energy, spectral entropy, the
  double-threshold detector and the entropy gate).  PINNED: the unmodified file is imported (shims for
  ``bayes_opt`` / ``seaborn``) and run on synthetic signals -> ``tests/golden/vad.npz``.
* ``oracle.frontend.sidekit_mfcc`` restates SIDEKIT 1.3.x ``frontend/features.py``
  ``mfcc`` (the function /root/reference/GMM_UBM.py:20,89 binds).  SIDEKIT is an
  un-vendored, un-pinned dependency (requirements.txt:5) that is absent from
  this container: **parity unpinned** except for the reference's own witnesses
  (98x13 cepstra for 1 s @ 16 kHz, report/final.pdf p.5; ``[0]`` is the cepstra,
  UI/GMM_UBM_GUI.py:91).
* ``oracle.frontend.sidekit_plp`` restates SIDEKIT ``plp`` (GMM_UBM.py:20, :94-99), a port of rastamat's
  ``rastaplp``; same absent package: **parity unpinned**.
* ``oracle.frontend.psf_mfcc`` restates python_speech_features 0.6 ``mfcc``
  (imported, never called, at /root/reference/GMM_UBM.py:13; BASELINE config 1
  quotes its 26-filter default).  Absent here: **parity unpinned**.
* ``oracle.frontend.librosa_mfcc`` / ``mfcc_lib`` restate ``librosa.feature.mfcc`` as /root/reference/MFCC_DTW.py:27-30
  calls it.  librosa is absent and un-pinned: **parity unpinned**; cross-checked (not pinned) against two
  independent implementations of the same conventions (torchaudio's MFCC transform, ``transformers.audio_utils``) in
  ``tests/test_oracle.py``.
* ``oracle.gmm.map_adapt`` is Reynolds/Quatieri/Dunn (2000) mean-only (and full)
  relevance MAP.  The reference has NO MAP code: **parity unpinned**,
  defined by formula.
"""

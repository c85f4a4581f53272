# shared-variance scoring kernel: parity tests, then timing at a few polynomial shares.  Fails fast: a hung kernel
# must not burn the GPU budget.
cd $GRAFT_REPO_ROOT
timeout 180 python -m pytest tests/test_gpu_gmm.py -x -q -k "shared_variance" 2>&1 | tail -15
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "TESTS FAILED"; exit 1; }
timeout 120 python benchmarks/prof_score_sv.py 2000 1000 1024 2>&1 | tail -1 || exit 1
for cfg in "0 4" "4 4" "8 4" "6 3" "8 3"; do
  set -- $cfg
  SV_COMPARE=0 SSP_SV_POLY_PAIRS=$1 SSP_SV_POLY_DEG=$2 timeout 120 python benchmarks/prof_score_sv.py 2000 1000 1024 2>&1 | tail -1
done

# quick A/B of the shared-variance kernel: parity tests, then timing (2000 utts x 1001 models) at a few settings
cd $GRAFT_REPO_ROOT
timeout 180 python -m pytest tests/test_gpu_gmm.py -x -q -k "shared_variance or map_enrol" 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "TESTS FAILED"; exit 1; }
for cfg in ${SV_CFGS:-"4 4" "6 4" "6 3"}; do
  set -- $cfg
  SV_COMPARE=0 SSP_SV_POLY_PAIRS=$1 SSP_SV_POLY_DEG=$2 timeout 120 python benchmarks/prof_score_sv.py 2000 1000 1024 2>&1 | tail -1 | cut -c 1-200
done

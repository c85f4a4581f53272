# EM / MAP statistics kernels: parity tests, then config 3 timing (quarter scale) with the pipelined and the single-buffered kernels
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_gmm.py -x -q -k "stats or em_trajectory or map_adapt or fit_default or map_enrol or GMM_trains or reference_pipeline" 2>&1 | tail -6
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "TESTS FAILED"; exit 1; }
timeout 200 python benchmarks/configs.py --only 3 --scale ${EM_SCALE:-0.25} 2>&1 | tail -1 | cut -c 1-200
SSP_EM_LSE=1 timeout 200 python benchmarks/configs.py --only 3 --scale ${EM_SCALE:-0.25} 2>&1 | tail -1 | cut -c 1-200
SSP_EM_LSE=1 SSP_EM_STATS=1 timeout 200 python benchmarks/configs.py --only 3 --scale ${EM_SCALE:-0.25} 2>&1 | tail -1 | cut -c 1-200

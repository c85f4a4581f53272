# A/B of the FMA-pipe polynomial share of the exponentials in gmm_score_sv_kernel (env knobs SSP_SV_POLY_PAIRS / SSP_SV_POLY_DEG),
# bench.py config 4 with the float64 oracle check.   gpurun -- 'bash benchmarks/sv_poly_ab.sh "4 4" "6 4" "6 3" ...'
cd $GRAFT_REPO_ROOT
for pd in "$@"; do
  set -- $pd
  SSP_SV_POLY_PAIRS=$1 SSP_SV_POLY_DEG=$2 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); c=l['check']; print('pairs $1 deg $2', 'kernel_ms', round(l['roofline']['kernel_ms'],1), 'sm_mhz', l['clocks']['sm_mhz'], 'W', l['clocks'].get('power_w_max'), 'rel', '%.3g' % c['oracle_max_rel'], 'llr', '%.3g' % c['oracle_max_llr_abs'])"
done

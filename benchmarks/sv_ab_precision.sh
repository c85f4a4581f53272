cd $GRAFT_REPO_ROOT
for lib in libssp_q05.so libssp_q15.so libssp_q0f.so default; do
  if [ "$lib" = "default" ]; then unset SSP_B200_LIB; else export SSP_B200_LIB=$GRAFT_REPO_ROOT/benchmarks/bin/$lib; fi
  timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); c=l['check']; print('$lib', 'kernel_ms', round(l['roofline']['kernel_ms'],1), 'sm_mhz', l['clocks']['sm_mhz'], 'rel', c['oracle_max_rel'], 'llr', c['oracle_max_llr_abs'], 'scoring_rel', c['oracle_scoring_only_max_rel'])"
done

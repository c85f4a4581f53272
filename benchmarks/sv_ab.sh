# A/B of shared-variance scoring kernel builds on ONE box: bench.py (config 4, no secondary legs, no CPU legs) per library.
#   gpurun -- 'bash benchmarks/sv_ab.sh lib1.so lib2.so ...'   (libraries under benchmarks/bin/; "default" = the in-tree build)
cd $GRAFT_REPO_ROOT
for lib in "$@"; do
  if [ "$lib" = "default" ]; then unset SSP_B200_LIB; else export SSP_B200_LIB=$GRAFT_REPO_ROOT/benchmarks/bin/$lib; fi
  for mb in ${GROUPS_MB:-24}; do
    SSP_SV_GROUP_MB=$mb timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-secondary --oracle-utts 0 2>&1 | tail -1 | python -c "import sys,json; l=json.loads(sys.stdin.read()); print('$lib', 'group_mb', '$mb', 'kernel_ms', round(l['roofline']['kernel_ms'],1), 'sm_mhz', l['clocks']['sm_mhz'], 'acc', round(l['check']['planted_speaker_accuracy'],4))"
  done
done

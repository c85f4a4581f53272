# Round check on one B200: GPU parity tests, the bench (both scorers), secondary configs, the ncu launch list of
# the bench command and full captures of the hot kernels.  Every step has its own timeout.
#   gpurun -- 'bash benchmarks/gpu_round_check.sh TAG [quick]'
cd $GRAFT_REPO_ROOT
TAG=${1:-r1d}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
grep -q "pytest exit 0" gpurun_out/${TAG}_pytest_gpu.log || exit 1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 2600 gpurun_out/${TAG}_bench_n1.json
timeout 600 python bench.py --steps 3 --warmup 3 --scorer general --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_general.json 2>> gpurun_out/${TAG}_bench_n1.err; cut -c 1-300 gpurun_out/${TAG}_bench_n1_general.json
[ "$2" = "quick" ] && exit 0
timeout 600 python benchmarks/configs.py --only 2,3,5 > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err; cut -c 1-260 gpurun_out/${TAG}_configs.jsonl | head -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gmm_|frontend|sv_fixup" -c 60 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gmm_score_sv_kernel -s 1 -c 1 -o gpurun_out/${TAG}_score_sv -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_score_sv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:frontend512_kernel -s 1 -c 1 -o gpurun_out/${TAG}_frontend512 -f python benchmarks/prof_frontend.py 4000 > gpurun_out/${TAG}_ncu_fe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gmm_em_ -s 2 -c 2 -o gpurun_out/${TAG}_em -f python benchmarks/prof_em.py 2000000 > gpurun_out/${TAG}_ncu_em.log 2>&1
ls -la gpurun_out | tail -16

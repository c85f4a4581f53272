# Round-2 check on one B200: GPU parity tests, smoke, the bench line, its ncu launch list, a full capture of the scoring
# launches of one step (DRAM traffic), sanitizer passes.   gpurun -- 'bash benchmarks/gpu_r2_check.sh TAG'
cd $GRAFT_REPO_ROOT
TAG=${1:-r2l}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 1500 gpurun_out/${TAG}_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gmm_|frontend|sv_fixup|em_" -c 80 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --oracle-utts 0 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
SV_COMPARE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gmm_score_sv_kernel -s 8 -c 4 -o gpurun_out/${TAG}_score_sv -f python benchmarks/prof_score_sv.py 10000 > gpurun_out/${TAG}_ncu_score_sv.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_score_sv.log
ls -la gpurun_out | grep ${TAG}

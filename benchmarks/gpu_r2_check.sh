# Round-2 check on one B200: GPU parity tests, smoke, the bench line, its ncu launch list, a full capture of the first scoring
# launch of a step and the DRAM traffic of all of its launches, sanitizer passes over smoke().
#   gpurun -- 'bash benchmarks/gpu_r2_check.sh TAG'
cd $GRAFT_REPO_ROOT
TAG=${1:-r2m}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 600 gpurun_out/${TAG}_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gmm_|frontend|sv_fixup|em_" -c 80 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --oracle-utts 0 > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
SV_COMPARE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gmm_score_sv_kernel -s 8 -c 1 -o gpurun_out/${TAG}_score_sv -f python benchmarks/prof_score_sv.py 10000 > gpurun_out/${TAG}_ncu_score_sv.log 2>&1; tail -2 gpurun_out/${TAG}_ncu_score_sv.log
SV_COMPARE=0 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:gmm_score_sv_kernel -s 8 -c 4 --csv --log-file gpurun_out/${TAG}_sv_dram_per_launch.csv python benchmarks/prof_score_sv.py 10000 > /dev/null 2>&1; tail -5 gpurun_out/${TAG}_sv_dram_per_launch.csv | cut -c 1-200
timeout 600 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "COMPUTE-SANITIZER|smoke ok|ERROR SUMMARY|Invalid|error" | head -8 > gpurun_out/${TAG}_sanitizer_memcheck.log; cat gpurun_out/${TAG}_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "COMPUTE-SANITIZER|smoke ok|RACECHECK SUMMARY|hazard|error" | head -8 > gpurun_out/${TAG}_sanitizer_racecheck.log; cat gpurun_out/${TAG}_sanitizer_racecheck.log
ls -la gpurun_out | grep ${TAG}

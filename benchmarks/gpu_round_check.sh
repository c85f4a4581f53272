set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1b_pytest_gpu.log
tail -5 gpurun_out/r1b_pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r1b_bench_n1.json 2> gpurun_out/r1b_bench_n1.err; tail -c 3000 gpurun_out/r1b_bench_n1.json
timeout 600 python benchmarks/configs.py --only 2,3,5 > gpurun_out/r1b_configs.jsonl 2> gpurun_out/r1b_configs.err; cut -c 1-400 gpurun_out/r1b_configs.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1b_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1b_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gmm_score_tc_kernel -s 1 -c 1 -o gpurun_out/r1b_score_tc -f python bench.py --steps 1 --warmup 1 --utts 2000 --no-cpu-baseline > gpurun_out/r1b_ncu_score.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:frontend512_kernel -s 1 -c 1 -o gpurun_out/r1b_frontend512 -f python benchmarks/prof_frontend.py 4000 > gpurun_out/r1b_ncu_fe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gmm_em_ -s 2 -c 2 -o gpurun_out/r1b_em_tc -f python benchmarks/prof_em.py 2000000 > gpurun_out/r1b_ncu_em.log 2>&1
ls -la gpurun_out

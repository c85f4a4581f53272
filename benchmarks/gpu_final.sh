# Round-end check on one B200 within a small GPU budget: parity tests, the bench line (with the CPU baseline leg),
# smoke(), the ncu launch list of the bench command.
# Budget note (round 1): this script ran 60 s on the box and was charged 177 s of the gpurun budget (acquire + push +
# overheads count); pytest-only calls ran 16-20 s and were charged 37-67 s; a 2-GPU call of 35 s was charged 127 s.
cd $GRAFT_REPO_ROOT
TAG=${1:-r1j}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 150 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; cat gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gmm_|frontend|sv_fixup" -c 40 --csv --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_launches_bench.csv | cut -c 1-200

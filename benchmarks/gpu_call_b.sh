cd $GRAFT_REPO_ROOT
TAG=${1:-r1g}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv,noheader
timeout 600 python -m pytest tests -m gpu -q --tb=short > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -30 gpurun_out/${TAG}_pytest_gpu.log
timeout 200 python tests/tools/new_paths_check.py > gpurun_out/${TAG}_new_paths.json 2> gpurun_out/${TAG}_new_paths.err; cat gpurun_out/${TAG}_new_paths.json; tail -5 gpurun_out/${TAG}_new_paths.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_nocpu.json 2> gpurun_out/${TAG}_bench_n1.err; cat gpurun_out/${TAG}_bench_n1_nocpu.json; tail -5 gpurun_out/${TAG}_bench_n1.err
# A/B of a second build of the library: SSP_B200_LIB=/path/to/other/libssp_b200.so python bench.py ...  (r1g compared a build
# with the in-warp partial sums carried in double this way; see DESIGN.md 4.1 "Reproducibility")

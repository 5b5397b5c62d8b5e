set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -s -k "real_lcm or batched_driver" > gpurun_out/r2_cfg5_tests.log 2>&1; echo "cfg5 tests rc $?"
tail -n 6 gpurun_out/r2_cfg5_tests.log
timeout 900 python bench.py --steps 5 --precision bf16 --no-longform --no-micro --no-batch1 > gpurun_out/r2_bench_cfg5.json 2> gpurun_out/r2_bench_cfg5.err; echo "bench rc $?"
tail -n 5 gpurun_out/r2_bench_cfg5.err

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -s -k "conv1d_layer or hybrid" > gpurun_out/r2_dit_tests.log 2>&1; echo "dit tests rc $?"; tail -n 12 gpurun_out/r2_dit_tests.log
timeout 900 python bench.py --steps 4 --precision bf16 --no-cpu --no-longform --no-micro --no-batch1 > gpurun_out/r2_bench_cfg5b.json 2> gpurun_out/r2_bench_cfg5b.err; echo "bench rc $?"; tail -n 5 gpurun_out/r2_bench_cfg5b.err

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lw in 20 100000; do
  ALCM_ACTPRO=0 ALCM_LANE_WAVES=$lw timeout 600 python bench.py --steps 4 --precision bf16 --no-cpu --no-longform --no-micro --no-config5 --no-batch1 > gpurun_out/r2_bench_lanes$lw.json 2> gpurun_out/r2_bench_lanes$lw.err; echo "bench lanes=$lw rc $?"
done
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "lcm_step" > gpurun_out/r2_lcm_step_test.log 2>&1; echo "lcm rc $?"; tail -n 3 gpurun_out/r2_lcm_step_test.log

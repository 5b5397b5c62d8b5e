set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/r2_ops_tests.log 2>&1; echo "ops rc $?"
timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -x -q -s > gpurun_out/r2_model_tests.log 2>&1; echo "models rc $?"
tail -n 3 gpurun_out/r2_ops_tests.log; tail -n 3 gpurun_out/r2_model_tests.log
for fs in 0 56 48 32 63; do
  ALCM_FUSE_STAGES=$fs timeout 600 python bench.py --steps 4 --precision bf16 --no-cpu --no-longform --no-micro > gpurun_out/r2_bench_fuse$fs.json 2> gpurun_out/r2_bench_fuse$fs.err; echo "fuse $fs rc $?"
done
for p in bf16 tf32; do
  python tools/one_act.py 64 24 160000 $p 2 > gpurun_out/r2_one_act_$p.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:act1d -s 3 -c 1 -f -o gpurun_out/prof_r2_act_v7_$p python tools/one_act.py 64 24 160000 $p 2 > gpurun_out/r2_one_act_${p}_ncu.log 2>&1
  echo "ncu $p rc $?"
done

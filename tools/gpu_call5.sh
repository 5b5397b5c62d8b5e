set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "act_conv1d or every_kernel_form" > gpurun_out/r2_actpro_tests.log 2>&1; echo "actpro ops rc $?"
tail -n 15 gpurun_out/r2_actpro_tests.log
timeout 1200 python -m pytest tests/test_gpu_models.py -m gpu -x -q -s -k "full_config or small_vs or forced or batch_and" > gpurun_out/r2_actpro_model_tests.log 2>&1; echo "actpro models rc $?"
tail -n 8 gpurun_out/r2_actpro_model_tests.log
for ap in 1 0; do
  ALCM_ACTPRO=$ap timeout 600 python bench.py --steps 4 --precision both --no-cpu --no-longform --no-micro --no-config5 > gpurun_out/r2_bench_actpro$ap.json 2> gpurun_out/r2_bench_actpro$ap.err; echo "bench actpro=$ap rc $?"
done
for v in 7 9; do ALCM_ACT_VARIANT=$v timeout 300 python tools/bench_act.py bf16,tf32 > gpurun_out/r2b_bench_act_v$v.log 2>&1; done
tail -n 4 gpurun_out/r2b_bench_act_v7.log gpurun_out/r2b_bench_act_v9.log

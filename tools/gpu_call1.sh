set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2_gpu.txt
timeout 1500 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/r2_ops_tests.log 2>&1; echo "ops rc $?" 
timeout 1500 python -m pytest tests/test_gpu_models.py -m gpu -x -q -s > gpurun_out/r2_model_tests.log 2>&1; echo "models rc $?"
for v in 3 7 8; do ALCM_ACT_VARIANT=$v timeout 300 python tools/bench_act.py bf16,tf32 > gpurun_out/r2_bench_act_v$v.log 2>&1; done
ALCM_UNPADDED=0 ALCM_ACT_VARIANT=3 timeout 300 python tools/bench_act.py bf16 > gpurun_out/r2_bench_act_v3_padded.log 2>&1
timeout 900 python bench.py --steps 5 > gpurun_out/r2_bench_first.json 2> gpurun_out/r2_bench_first.err; echo "bench rc $?"
tail -3 gpurun_out/r2_ops_tests.log gpurun_out/r2_model_tests.log

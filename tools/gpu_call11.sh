set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "activation" > gpurun_out/r2_act_tests.log 2>&1; echo "act tests rc $?"; tail -n 3 gpurun_out/r2_act_tests.log
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "full_config or forced or guard" > gpurun_out/r2_model_tests.log 2>&1; echo "models rc $?"; tail -n 3 gpurun_out/r2_model_tests.log
for v in 0 1 2; do ALCM_ACT_VARIANT=$v timeout 300 python tools/bench_act.py bf16,tf32 > gpurun_out/r2d_bench_act_v$v.log 2>&1; done
paste -d'|' <(cut -c1-75 gpurun_out/r2d_bench_act_v0.log) <(cut -c50-75 gpurun_out/r2d_bench_act_v1.log) <(cut -c50-75 gpurun_out/r2d_bench_act_v2.log)
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 4 --precision both --no-cpu --no-longform --no-config5 > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err; echo "bench $tag rc $?"; }
run actv1 ALCM_ACT_VARIANT=1
run actv2 ALCM_ACT_VARIANT=2

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/r2_ops_tests.log 2>&1; echo "ops rc $?"; tail -n 3 gpurun_out/r2_ops_tests.log
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "full_config or small_vs or forced or batch_and or batch64 or guard" > gpurun_out/r2_model_tests.log 2>&1; echo "models rc $?"; tail -n 3 gpurun_out/r2_model_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 4 --precision both --no-cpu --no-longform --no-micro --no-config5 > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err; echo "bench $tag rc $?"; }
run wres1 ALCM_W_RESIDENT=1
run wres0 ALCM_W_RESIDENT=0

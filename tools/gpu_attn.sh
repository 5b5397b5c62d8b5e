set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attn" > gpurun_out/r2_attn_tests.log 2>&1; echo "attn ops rc $?"; tail -n 15 gpurun_out/r2_attn_tests.log
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "vae or full_path or guard" > gpurun_out/r2_attn_model_tests.log 2>&1; echo "attn models rc $?"; tail -n 8 gpurun_out/r2_attn_model_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 4 --precision both --no-cpu --no-longform --no-micro --no-config5 > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err; echo "bench $tag rc $?"; }
run attn1 ALCM_ATTN_TC=1
run attn0 ALCM_ATTN_TC=0

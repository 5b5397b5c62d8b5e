set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -s -k "long_sequence" > gpurun_out/r2_long_tests.log 2>&1; echo "long tests rc $?"; tail -n 12 gpurun_out/r2_long_tests.log
timeout 900 python bench.py --steps 3 --precision bf16 --no-cpu --no-micro --no-batch1 --no-config5 > gpurun_out/r2_bench_long.json 2> gpurun_out/r2_bench_long.err; echo "bench rc $?"; tail -n 5 gpurun_out/r2_bench_long.err
python -c "import json; d=json.load(open('gpurun_out/r2_bench_long.json')); print(d['longform'])"

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > gpurun_out/r2_ops_tests.log 2>&1; echo "ops rc $?"; tail -n 3 gpurun_out/r2_ops_tests.log
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -k "full_config or small_vs or forced or batch_and or batch64" > gpurun_out/r2_model_tests.log 2>&1; echo "models rc $?"; tail -n 3 gpurun_out/r2_model_tests.log
run() { tag=$1; shift; env "$@" timeout 600 python bench.py --steps 4 --precision both --no-cpu --no-longform --no-micro --no-config5 > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err; echo "bench $tag rc $?"; }
run as1 ALCM_AS_TILES=1
run as0 ALCM_AS_TILES=0
for v in 0 1; do ALCM_ACT_VARIANT=$v timeout 300 python tools/bench_act.py bf16,tf32 > gpurun_out/r2c_bench_act_v$v.log 2>&1; done
for p in bf16 tf32; do
  python tools/one_decode.py $p 64 > gpurun_out/r2_one_decode_${p}_64.log 2>&1 && \
  ncu --nvtx --nvtx-include "alcm_decode/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2_${p}_b64.csv python tools/one_decode.py $p 64 > gpurun_out/r2_one_decode_${p}_64_ncu.log 2>&1
  echo "ncu launches $p rc $?"
done

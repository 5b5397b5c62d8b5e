set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
ALCM_TRACE=1 timeout 300 python tools/bench_conv.py bf16 0 5 > gpurun_out/r2_conv_narrow_trace.log 2>&1
ALCM_W_RESIDENT=0 timeout 300 python tools/bench_conv.py bf16 0,1,2,3 5 > gpurun_out/r2_conv_narrow_dbg_wres0.log 2>&1
ALCM_W_RESIDENT=1 timeout 300 python tools/bench_conv.py bf16 0,1,2,3 5 > gpurun_out/r2_conv_narrow_dbg_wres1.log 2>&1
ALCM_PERSIST=0 timeout 300 python tools/bench_conv.py bf16 0 5 > gpurun_out/r2_conv_narrow_nopersist.log 2>&1
cat gpurun_out/r2_conv_narrow_trace.log gpurun_out/r2_conv_narrow_dbg_wres0.log gpurun_out/r2_conv_narrow_dbg_wres1.log gpurun_out/r2_conv_narrow_nopersist.log

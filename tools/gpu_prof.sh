set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# top kernels, --set full, random operands
python tools/one_conv.py 8 768 768 2500 11 1 bf16 3 > gpurun_out/r2_one_conv8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 3 -c 1 -f -o gpurun_out/prof_r2_conv_s1k11_b8 python tools/one_conv.py 8 768 768 2500 11 1 bf16 3 > gpurun_out/r2_one_conv8_ncu.log 2>&1; echo "ncu conv b8 rc $?"
python tools/one_conv.py 1 768 768 2500 11 1 bf16 3 > gpurun_out/r2_one_conv1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 3 -c 1 -f -o gpurun_out/prof_r2_conv_s1k11_b1 python tools/one_conv.py 1 768 768 2500 11 1 bf16 3 > gpurun_out/r2_one_conv1_ncu.log 2>&1; echo "ncu conv b1 rc $?"
python tools/one_conv.py 64 24 24 160000 11 5 bf16 3 > gpurun_out/r2_one_conv_s6.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 3 -c 1 -f -o gpurun_out/prof_r2_conv_s6k11_b64 python tools/one_conv.py 64 24 24 160000 11 5 bf16 3 > gpurun_out/r2_one_conv_s6_ncu.log 2>&1; echo "ncu conv s6 rc $?"
for p in bf16 tf32; do
  python tools/one_act.py 64 24 160000 $p 2 > gpurun_out/r2_one_act_$p.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:act1d -s 3 -c 1 -f -o gpurun_out/prof_r2_act_$p python tools/one_act.py 64 24 160000 $p 2 > gpurun_out/r2_one_act_${p}_ncu.log 2>&1
  echo "ncu act $p rc $?"
done
# launch list of one batch-1 decode (configs[1])
python tools/one_decode.py bf16 1 > gpurun_out/r2_one_decode_bf16_1.log 2>&1 && \
ncu --nvtx --nvtx-include "alcm_decode/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_r2_bf16_b1.csv python tools/one_decode.py bf16 1 > gpurun_out/r2_one_decode_bf16_1_ncu.log 2>&1; echo "ncu launches b1 rc $?"

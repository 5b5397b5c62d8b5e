# bench.py at N GPUs of one box (usage: bash tools/gpu_scale.sh N), as the driver launches it
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
nvidia-smi -L > gpurun_out/r2_gpus_n$N.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc $?"
tail -n 3 gpurun_out/r2_bench_n$N.err

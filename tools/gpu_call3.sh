set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus_n2.txt
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q -s -k "nccl or guard or two_threads or preplanned" > gpurun_out/r2_nccl_tests.log 2>&1; echo "nccl tests rc $?"
tail -n 5 gpurun_out/r2_nccl_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc $?"
tail -n 3 gpurun_out/r2_bench_n2.err

set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc $?"; tail -n 4 gpurun_out/r2_smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests_full.log 2>&1; echo "gpu tests rc $?"; tail -n 4 gpurun_out/r2_gpu_tests_full.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc $?"; tail -n 3 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc $?"
python - <<'PY'
import torch, time
x = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
for _ in range(3): x.zero_()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): x.zero_()
e1.record(); torch.cuda.synchronize()
print("pure write (memset 2 GiB): %.0f GB/s" % (10 * x.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9))
y = torch.empty_like(x)
for _ in range(3): y.copy_(x)
torch.cuda.synchronize()
e0.record()
for _ in range(10): y.copy_(x)
e1.record(); torch.cuda.synchronize()
print("copy (read+write bytes): %.0f GB/s" % (10 * 2 * x.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9))
PY

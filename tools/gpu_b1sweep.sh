set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat > /tmp/b1.py <<'PY'
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
from bench import build_pipe, T_LAT, clip_latent, time_steps
prec = sys.argv[1]
pipe = build_pipe(prec, "cuda:0")
z = torch.from_numpy(clip_latent(0)).to("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
pipe.plan(1, T_LAT); pipe.decode_tensor(z); torch.cuda.synchronize()
s = time_steps(lambda: pipe.decode_tensor(z), 30, 5, flush)
print(f"{prec} batch-1 decode: {1e3 * s / 30:.4f} ms  env: " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("ALCM_")))
PY
for p in bf16 tf32; do
  python /tmp/b1.py $p
  ALCM_PDL=1 python /tmp/b1.py $p
  ALCM_SMEM_BUDGET_1W=100000 python /tmp/b1.py $p
  ALCM_SMEM_BUDGET_1W=160000 python /tmp/b1.py $p
  ALCM_ACT_VARIANT=2 python /tmp/b1.py $p
  ALCM_ACT_VARIANT=0 python /tmp/b1.py $p
  ALCM_LANES=0 python /tmp/b1.py $p
  ALCM_NT192=192 python /tmp/b1.py $p
  ALCM_W_RESIDENT=0 python /tmp/b1.py $p
done 2>&1 | grep "batch-1 decode" > gpurun_out/r2_b1_sweep.log
cat gpurun_out/r2_b1_sweep.log

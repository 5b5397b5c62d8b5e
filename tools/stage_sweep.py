"""Per-stage, per-class milliseconds PER CLIP at several batch sizes (eager replay, one event pair per kernel).
    python tools/stage_sweep.py [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_pipe, T_LAT  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
pipe = build_pipe(prec, "cuda:0")
rows = {}
Bs = tuple(int(b) for b in os.environ.get("SWEEP_B", "2,4,8,16,64").split(","))
for B in Bs:
    pipe.profile_stages(B, T_LAT, 1)
    for st, cl in pipe.profile_stages(B, T_LAT, 3).items():
        for name, v in cl.items():
            rows.setdefault((st, name), {})[B] = v["ms"] / B
print(f"{prec}: ms per clip   " + "".join(f"B={B:<8d}" for B in Bs))
for (st, name), v in rows.items():
    print(f"{st:>9s} {name:>5s}  " + "".join(f"{v.get(B, 0):<10.4f}" for B in Bs))

#!/bin/bash
# A/B sweep of the engine's tuning knobs on the batch-1 decode (run on a B200: bash tools/sweep_env.sh VAR=val ...)
run() { env "$@" timeout 150 python bench.py --steps 30 --warmup 3 --no-cpu --no-batch64 > gpurun_out/bench_sw.json 2>/dev/null; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_sw.json')); print(' '.join(sys.argv[1:]), d['value'], d['ms_per_step'])" "$@"; }
if [ $# -gt 0 ]; then for kv in "$@"; do run $kv; done; exit 0; fi
run BASE=1
run ALCM_SMEM_BUDGET_1W=120000
run ALCM_SMEM_BUDGET_1W=150000
run ALCM_SMEM_BUDGET_MW=120000
run ALCM_SMEM_BUDGET_MW=75000
run ALCM_SMEM_BUDGET_1W=120000 ALCM_TPG=3
run BASE=2

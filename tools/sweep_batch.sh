run() { env $1 timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu --no-batch64 --batch $2 > gpurun_out/bench_sw.json 2>/dev/null; python -c "
import json,sys; d=json.load(open('gpurun_out/bench_sw.json')); print(sys.argv[1], 'B=',sys.argv[2], d['value'], d['ms_per_step'])" $1 $2; }
for b in 8 64; do for w in 128 16; do run ALCM_WIDE_NT=$w $b; done; done

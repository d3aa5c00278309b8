set -x
O=gpurun_out/c7b; mkdir -p $O
export HGPU_STRUCT=1
B="timeout 60 python bench.py --no-cpu-baseline --no-e2e --steps 60 --warmup 5"
HGPU_GENERIC_COST=4 $B > $O/bench_c4.json 2> $O/bench_c4.err
HGPU_GENERIC_COST=6.5 $B > $O/bench_c65.json 2> $O/bench_c65.err
HGPU_GENERIC_COST=10 $B > $O/bench_c10.json 2> $O/bench_c10.err
for f in bench_c4 bench_c65 bench_c10; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", round(d["value"]/1e9,3), round(d["ms_per_step"],4), 'kernel', round(d["roofline"]["kernel_ms"],4), 'frac', round(d["roofline"]["frac"],3), d["layout"]["ntiles"], d["layout"]["struct_tiles"])
except Exception as e: print("$f", "ERR", e); print(open("$O/$f.err").read()[-600:])
PY
done

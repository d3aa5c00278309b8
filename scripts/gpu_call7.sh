set -x
O=gpurun_out/c7; mkdir -p $O
export HGPU_STRUCT=1
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -k "structured" > $O/pytest_struct.log 2>&1; echo "rc=$?" >> $O/pytest_struct.log
tail -n 6 $O/pytest_struct.log
B="timeout 70 python bench.py --no-cpu-baseline --no-e2e --steps 100"
HGPU_GENERIC_COST=2.5 $B > $O/bench_c25.json 2> $O/bench_c25.err
HGPU_STRUCT=0 $B > $O/bench_nostruct.json 2> $O/bench_nostruct.err
for f in bench_c25 bench_nostruct; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", round(d["value"]/1e9,3), round(d["ms_per_step"],4), 'kernel', round(d["roofline"]["kernel_ms"],4), 'frac', round(d["roofline"]["frac"],3), d["layout"]["grid_ctas"], d["layout"]["struct_tiles"])
except Exception as e: print("$f", "ERR", e); print(open("$O/$f.err").read()[-1500:])
PY
done

set -x
O=gpurun_out/c2d; mkdir -p $O
B="timeout 200 python bench.py --no-cpu-baseline --no-e2e"
HGPU_DYNAMIC=0 HGPU_GENERIC_COST=2.6 $B > $O/bench_static_morton.json 2> $O/bench_static_morton.err
HGPU_DYNAMIC=0 HGPU_GENERIC_COST=2.6 HGPU_ORDER=level $B > $O/bench_static_level.json 2> $O/bench_static_level.err
HGPU_ORDER=level $B > $O/bench_dyn_level.json 2> $O/bench_dyn_level.err
HGPU_ORDER=level HGPU_GENERIC_COST=3.2 $B > $O/bench_dyn_level_c32.json 2> $O/bench_dyn_level_c32.err
HGPU_STRUCT=0 HGPU_ORDER=level $B > $O/bench_nostruct_level.json 2> $O/bench_nostruct_level.err
HGPU_STRUCT=0 $B > $O/bench_nostruct_morton.json 2> $O/bench_nostruct_morton.err
HGPU_ORDER=level timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "structured or adaptive_workload or whole_run" > $O/pytest_level.log 2>&1; echo "rc=$?" >> $O/pytest_level.log; tail -n 3 $O/pytest_level.log
N="--set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1"
HGPU_DYNAMIC=0 HGPU_GENERIC_COST=2.6 timeout 300 ncu $N -o $O/prof_static_morton python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu1.log 2>&1
HGPU_ORDER=level timeout 300 ncu $N -o $O/prof_dyn_level python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu2.log 2>&1
for f in bench_static_morton bench_static_level bench_dyn_level bench_dyn_level_c32 bench_nostruct_level bench_nostruct_morton; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", round(d["value"]/1e9,3), round(d["ms_per_step"],4), 'kernel', round(d["roofline"]["kernel_ms"],4), 'frac', round(d["roofline"]["frac"],3))
except Exception as e: print("$f", "ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done

set -x
O=gpurun_out/c2c; mkdir -p $O
timeout 420 python -m pytest tests/test_gpu_parity.py -x -q -k "structured or properties or adaptive_workload or whole_run or wpass" > $O/pytest_struct.log 2>&1; echo "rc=$?" >> $O/pytest_struct.log
tail -5 $O/pytest_struct.log
HGPU_DYNAMIC=0 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "structured or adaptive_workload" > $O/pytest_struct_static.log 2>&1; echo "rc=$?" >> $O/pytest_struct_static.log
tail -3 $O/pytest_struct_static.log
timeout 200 python bench.py --no-cpu-baseline > $O/bench_dyn.json 2> $O/bench_dyn.err; echo rc=$?
HGPU_GENERIC_COST=3.0 timeout 200 python bench.py --no-cpu-baseline --no-e2e > $O/bench_dyn_c30.json 2> $O/bench_dyn_c30.err; echo rc=$?
HGPU_DYNAMIC=0 HGPU_GENERIC_COST=2.2 timeout 200 python bench.py --no-cpu-baseline --no-e2e > $O/bench_static_c22.json 2> $O/bench_static_c22.err; echo rc=$?
HGPU_DYNAMIC=0 HGPU_GENERIC_COST=2.8 timeout 200 python bench.py --no-cpu-baseline --no-e2e > $O/bench_static_c28.json 2> $O/bench_static_c28.err; echo rc=$?
HGPU_DYNAMIC=0 HGPU_GENERIC_COST=3.6 timeout 200 python bench.py --no-cpu-baseline --no-e2e > $O/bench_static_c36.json 2> $O/bench_static_c36.err; echo rc=$?
HGPU_STRUCT=0 timeout 200 python bench.py --no-cpu-baseline --no-e2e > $O/bench_nostruct.json 2> $O/bench_nostruct.err; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o $O/prof_struct python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_struct.log 2>&1; echo rc=$?
for f in bench_dyn bench_dyn_c30 bench_static_c22 bench_static_c28 bench_static_c36 bench_nostruct; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", round(d["value"]/1e9,3), round(d["ms_per_step"],4), 'kernel', round(d["roofline"]["kernel_ms"],4), 'frac', round(d["roofline"]["frac"],3), 'e2e', d["e2e"] and round(d["e2e"]["value"]/1e9,3))
except Exception as e: print("$f", "ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log; tail -4 $O/pytest_all.log

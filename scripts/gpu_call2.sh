set -x
O=gpurun_out/c2b; mkdir -p $O
timeout 420 python -m pytest tests/test_gpu_parity.py -x -q -k "structured or properties or adaptive_workload or whole_run or wpass" > $O/pytest_struct.log 2>&1; echo "rc=$?" >> $O/pytest_struct.log
tail -15 $O/pytest_struct.log
timeout 200 python bench.py --no-cpu-baseline > $O/bench_struct.json 2> $O/bench_struct.err; echo rc=$?
HGPU_STRUCT=0 timeout 200 python bench.py --no-cpu-baseline --no-e2e > $O/bench_nostruct.json 2> $O/bench_nostruct.err; echo rc=$?
HGPU_GENERIC_COST=2.2 timeout 200 python bench.py --no-cpu-baseline --no-e2e > $O/bench_struct_cost22.json 2> $O/bench_struct_cost22.err; echo rc=$?
HGPU_GENERIC_COST=1.2 timeout 200 python bench.py --no-cpu-baseline --no-e2e > $O/bench_struct_cost12.json 2> $O/bench_struct_cost12.err; echo rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o $O/prof_struct python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_struct.log 2>&1; echo rc=$?
for f in bench_struct bench_nostruct bench_struct_cost22 bench_struct_cost12; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"] and d["e2e"]["value"]/1e9, d["layout"])
except Exception as e: print("$f", "ERR", e)
PY
done
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_all.log 2>&1; echo "rc=$?" >> $O/pytest_all.log; tail -5 $O/pytest_all.log

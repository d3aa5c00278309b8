set -x
O=gpurun_out/c4; mkdir -p $O
export HGPU_STRUCT=1
timeout 420 python -m pytest tests/test_gpu_parity.py -x -q -k "structured or properties or adaptive_workload or whole_run" > $O/pytest_struct.log 2>&1; echo "rc=$?" >> $O/pytest_struct.log
tail -n 4 $O/pytest_struct.log
B="timeout 200 python bench.py --no-cpu-baseline --no-e2e"
HGPU_GENERIC_COST=2.6 $B > $O/bench_c26.json 2> $O/bench_c26.err
HGPU_GENERIC_COST=3.2 $B > $O/bench_c32.json 2> $O/bench_c32.err
HGPU_GENERIC_COST=4.0 $B > $O/bench_c40.json 2> $O/bench_c40.err
HGPU_GENERIC_COST=5.0 $B > $O/bench_c50.json 2> $O/bench_c50.err
HGPU_STRUCT=0 $B > $O/bench_nostruct.json 2> $O/bench_nostruct.err
N="--set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1"
HGPU_GENERIC_COST=3.2 timeout 300 ncu $N -o $O/prof_struct python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu1.log 2>&1
for f in bench_c26 bench_c32 bench_c40 bench_c50 bench_nostruct; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", round(d["value"]/1e9,3), round(d["ms_per_step"],4), 'kernel', round(d["roofline"]["kernel_ms"],4), 'frac', round(d["roofline"]["frac"],3))
except Exception as e: print("$f", "ERR", e); print(open("$O/$f.err").read()[-800:])
PY
done
unset HGPU_STRUCT
timeout 600 python -m pytest tests/test_integration.py -x -q > $O/pytest_integration.log 2>&1; echo "rc=$?" >> $O/pytest_integration.log; tail -n 5 $O/pytest_integration.log

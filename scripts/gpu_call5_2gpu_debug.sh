set -x
O=gpurun_out/c5c; mkdir -p $O
export BENCH_WATCHDOG_S=100
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 150 $TR bench.py --gpus 2 --no-cpu-baseline --no-e2e > $O/a_default.json 2> $O/a_default.err; echo "a rc=$?"
HGPU_COOPERATIVE=0 timeout 150 $TR bench.py --gpus 2 --no-cpu-baseline --no-e2e > $O/b_nocoop.json 2> $O/b_nocoop.err; echo "b rc=$?"
timeout 150 $TR bench.py --gpus 2 --no-cpu-baseline --no-e2e --steps 120 --warmup 5 > $O/c_120.json 2> $O/c_120.err; echo "c rc=$?"
grep -h "bench rank\|Error\|error\|File \"/tmp/code\|line [0-9]* in" $O/*.err | cut -c1-200 | head -90
for f in a_default b_nocoop c_120; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", d["value"]/1e9, d["ms_per_step"], d["parity_check"] and d["parity_check"]["rel_l2"], d["phases_ms_per_step"])
except Exception as e: print("$f", "ERR", e)
PY
done

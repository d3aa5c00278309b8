# 2 GPUs: why does the default mode of bench.py --gpus 2 hang?  small meshes, watchdog, variants
set -x
O=gpurun_out/c5; mkdir -p $O
export BENCH_WATCHDOG_S=100
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
B="bench.py --gpus 2 --no-cpu-baseline --no-e2e --edge 128 --steps 50 --warmup 5"
timeout 150 $TR $B > $O/a_default.json 2> $O/a_default.err; echo "a rc=$?"
HGPU_COOPERATIVE=0 timeout 150 $TR $B > $O/b_nocoop.json 2> $O/b_nocoop.err; echo "b rc=$?"
timeout 150 $TR $B --no-parity-check > $O/c_noparity.json 2> $O/c_noparity.err; echo "c rc=$?"
HGPU_COOPERATIVE=0 timeout 150 $TR $B --no-parity-check > $O/d_nocoop_noparity.json 2> $O/d_nocoop_noparity.err; echo "d rc=$?"
timeout 150 $TR $B --tail-overlap > $O/e_tail.json 2> $O/e_tail.err; echo "e rc=$?"
grep -h "bench rank\|Error\|error\|File \"/tmp/code" $O/*.err | cut -c1-220 | head -80
for f in a_default b_nocoop c_noparity d_nocoop_noparity e_tail; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", d["value"]/1e9, d["ms_per_step"], d["parity_check"] and d["parity_check"]["rel_l2"], d["phases_ms_per_step"])
except Exception as e: print("$f", "ERR", e)
PY
done

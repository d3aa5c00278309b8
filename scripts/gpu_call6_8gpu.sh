# 8 GPUs: multi-rank parity on real devices at 3 / 4 / 8 ranks (NCCL + peer memory), then the strong-scaling
# series on ONE adaptive mesh (configs[2]'s 96.6 M elements as 8 columns) and configs[3] at scale (773 M)
set -x
O=gpurun_out/c6; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
export BENCH_WATCHDOG_S=240
timeout 420 python -m pytest tests/test_multirank_gpu.py -x -q -k "np8 or np4 or np3" > $O/pytest_multirank_8gpu.log 2>&1; echo "rc=$?" >> $O/pytest_multirank_8gpu.log; tail -n 4 $O/pytest_multirank_8gpu.log
timeout 200 python -m pytest tests/test_integration.py -x -q -k multirank > $O/pytest_integration_multirank.log 2>&1; echo "rc=$?" >> $O/pytest_integration_multirank.log; tail -n 3 $O/pytest_integration_multirank.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29512"
S="bench.py --no-cpu-baseline --workload adaptive --strong --columns 8 --edge 256 --steps 60 --warmup 5"
timeout 200 $TR --nproc-per-node 8 $S --gpus 8 > $O/strong97M_n8.json 2> $O/strong97M_n8.err; echo "rc=$?"
timeout 200 $TR --nproc-per-node 4 $S --gpus 4 > $O/strong97M_n4.json 2> $O/strong97M_n4.err; echo "rc=$?"
timeout 400 $TR --nproc-per-node 8 bench.py --no-cpu-baseline --workload adaptive --strong --columns 8 --edge 512 --steps 40 --warmup 5 --gpus 8 > $O/strong773M_n8.json 2> $O/strong773M_n8.err; echo "rc=$?"
for f in strong97M_n8 strong97M_n4 strong773M_n8; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", round(d["value"]/1e9,2), round(d["ms_per_step"],4), d["scaling"], d["e2e"] and round(d["e2e"]["value"]/1e9,2), d["parity_check"] and d["parity_check"]["rel_l2"], d["config"]["global_elements"], d["phases_ms_per_step"])
except Exception as e: print("$f", "ERR", e); print(open("$O/$f.err").read()[-1200:])
PY
done

# 2 GPUs: multi-rank parity on REAL devices (NCCL + peer-memory transports), then benches with parity_check
set -x
O=gpurun_out/c3; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
export HGPU_TEST_TAIL_OVERLAP=1
timeout 700 python -m pytest tests/test_multirank_gpu.py -x -q > $O/pytest_multirank_2gpu.log 2>&1; echo "rc=$?" >> $O/pytest_multirank_2gpu.log; tail -6 $O/pytest_multirank_2gpu.log
timeout 400 python -m pytest tests/test_integration.py -x -q -k multirank > $O/pytest_integration_multirank.log 2>&1; echo "rc=$?" >> $O/pytest_integration_multirank.log; tail -4 $O/pytest_integration_multirank.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 2 --no-cpu-baseline > $O/bench_n2.json 2> $O/bench_n2.err; echo rc=$?
timeout 300 $TR bench.py --gpus 2 --no-cpu-baseline --tail-overlap > $O/bench_n2_tail.json 2> $O/bench_n2_tail.err; echo rc=$?
timeout 300 $TR bench.py --gpus 2 --no-cpu-baseline --halo nccl --no-e2e > $O/bench_n2_nccl.json 2> $O/bench_n2_nccl.err; echo rc=$?
timeout 700 $TR bench.py --gpus 2 --no-cpu-baseline --workload adaptive --edge 384 --strong --steps 50 --warmup 5 > $O/bench_adaptive384_strong_n2.json 2> $O/bench_adaptive384_strong_n2.err; echo rc=$?
for f in bench_n2 bench_n2_tail bench_n2_nccl bench_adaptive384_strong_n2; do python - <<PY
import json
try:
    d=json.load(open("$O/$f.json")); print("$f", d["value"]/1e9, d["ms_per_step"], d["e2e"] and d["e2e"]["value"]/1e9, d["parity_check"] and d["parity_check"]["rel_l2"], d["phases_ms_per_step"])
except Exception as e: print("$f", "ERR", e); print(open("$O/$f.err").read()[-1500:])
PY
done

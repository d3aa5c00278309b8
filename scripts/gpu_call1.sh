set -x
mkdir -p gpurun_out/c1
nvidia-smi -L > gpurun_out/c1/gpus.txt
export HGPU_TEST_WPASS=1 HGPU_TEST_TAIL_OVERLAP=1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c1/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c1/pytest.log
tail -5 gpurun_out/c1/pytest.log
timeout 400 python bench.py > gpurun_out/c1/bench_n1.json 2> gpurun_out/c1/bench_n1.err; echo rc=$?
timeout 200 python bench.py --wpass --no-cpu-baseline > gpurun_out/c1/bench_n1_wpass.json 2> gpurun_out/c1/bench_n1_wpass.err; echo rc=$?
timeout 500 python bench.py --workload adaptive --edge 512 --no-cpu-baseline --steps 100 > gpurun_out/c1/bench_adaptive512.json 2> gpurun_out/c1/bench_adaptive512.err; echo rc=$?
timeout 700 python bench.py --workload basin --edge 1024 --damping bkt --no-cpu-baseline --steps 40 > gpurun_out/c1/bench_basin1024.json 2> gpurun_out/c1/bench_basin1024.err; echo rc=$?
head -c 600 gpurun_out/c1/bench_n1.json; echo; head -c 400 gpurun_out/c1/bench_n1_wpass.json

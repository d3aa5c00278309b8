# the round's last GPU seconds: first timing of the conventional-stiffness (DENSE) step kernel, then the default
# kernel on the same mesh in the same call (edge 160: 4.1 M elements, working set 1.3 GB per step >> L2)
O=gpurun_out/c11; mkdir -p $O
timeout 13 python bench.py --edge 160 --stiffness conventional --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_conventional_160.json 2> $O/bench_conventional_160.err
echo "conv rc=$?"; cat $O/bench_conventional_160.json | cut -c1-300
timeout 13 python bench.py --edge 160 --steps 30 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_effective_160.json 2> $O/bench_effective_160.err
echo "eff rc=$?"; cat $O/bench_effective_160.json | cut -c1-300

O=gpurun_out/c8; mkdir -p $O
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file $O/launches.csv python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
echo "ncu rc=$?"; tail -n 3 $O/launches.csv | cut -c1-200

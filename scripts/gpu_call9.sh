# last 80 s of the round's GPU time: the device-plane tests only (tests/test_zz_planes_gpu.py)
O=gpurun_out/c9; mkdir -p $O
timeout 60 python -m pytest tests/test_zz_planes_gpu.py -x -q > $O/pytest_planes.log 2>&1
echo "pytest rc=$?"; tail -n 15 $O/pytest_planes.log

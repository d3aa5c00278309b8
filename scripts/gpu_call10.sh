# device planes as the single-rank default of psolve_gpu: the plane tests of both files
O=gpurun_out/c10; mkdir -p $O
timeout 40 python -m pytest tests/test_zz_planes_gpu.py tests/test_integration.py -k planes -x -q > $O/pytest_planes.log 2>&1
echo "pytest rc=$?"; tail -n 15 $O/pytest_planes.log

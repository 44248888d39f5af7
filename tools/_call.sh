set -x
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -12 | tee gpurun_out/r2_pytest_gpu_sharded_n2_final.log

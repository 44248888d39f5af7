set -x
timeout 900 python -m pytest tests/test_gpu_color.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py -m gpu -x -q 2>&1 | tail -4

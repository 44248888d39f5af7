set -x
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_slabs_one_device.py tests/test_gpu_color.py tests/test_gpu_baseline_sizes.py -m gpu -x -q 2>&1 | tail -2
echo "== skip on"; timeout 120 python tools/stage_times.py 2>&1 | grep -E "preprocess|icp_track|raycast_tracked|integrate_tracked|sum of"
echo "== skip off"; HKF_RAY_SKIP=0 timeout 120 python tools/stage_times.py 2>&1 | grep -E "raycast_tracked|integrate_tracked|sum of"

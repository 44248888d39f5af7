set -x
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_slabs_one_device.py tests/test_gpu_color.py -m gpu -x -q 2>&1 | tail -3
echo "== default"; timeout 120 python tools/stage_times.py 2>&1 | grep -E "preprocess|icp_track|raycast_tracked|integrate_tracked|sum of|N_upd"
timeout 120 python tools/stage_times.py 256 2>&1 | grep -E "integrate_tracked|N_upd"

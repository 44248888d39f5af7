set -x
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_baseline_sizes.py -m gpu -x -q 2>&1 | tail -6
echo "== default"; timeout 120 python tools/stage_times.py 2>&1 | grep -E "preprocess|icp_track|raycast_tracked|integrate_tracked|sum of|N_upd"
HKF_LIB=hybkinectfu_b200/libhkf_b200_stamps.so timeout 120 python tools/icp_stamps.py > gpurun_out/r2_icp_frame_stamps_e.txt 2>&1; cat gpurun_out/r2_icp_frame_stamps_e.txt

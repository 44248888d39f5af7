set -x
timeout 120 python tools/stage_times.py 2>&1 | grep -E "integrate_tracked|raycast_tracked|icp_track|preprocess|sum of"

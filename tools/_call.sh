set -x
echo "== 16x16 tile, 256 threads"; timeout 120 python tools/stage_times.py 2>&1 | grep -E "raycast_tracked|sum of"
echo "== 16x8 tile, 128 threads"; HKF_LIB=hybkinectfu_b200/libhkf_b200_ray8.so timeout 120 python tools/stage_times.py 2>&1 | grep -E "raycast_tracked|sum of"
echo "== 16x4 tile, 64 threads"; HKF_LIB=hybkinectfu_b200/libhkf_b200_ray4.so timeout 120 python tools/stage_times.py 2>&1 | grep -E "raycast_tracked|sum of"

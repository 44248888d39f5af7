#!/bin/bash
# One-shot check of the off-by-default device paths on a GPU box:
#   gpurun --timeout 400 -- 'bash tools/try_experimental.sh > gpurun_out/experimental.log 2>&1; tail -40 gpurun_out/experimental.log'
# 1. their bit-exactness tests, 2. per-stage times with each switch, 3. the default for comparison.
cd "$(dirname "$0")/.." || exit 1
echo "== gated tests"
HKF_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -x -q 2>&1 | tail -15
echo "== default"
timeout 60 python tools/stage_times.py 2>&1 | grep -E "preprocess|icp_track|raycast_tracked|integrate_tracked|sum of"
echo "== HKF_RAY_SKIP=1"
HKF_RAY_SKIP=1 timeout 60 python tools/stage_times.py 2>&1 | grep -E "icp_track|raycast_tracked|integrate_tracked|sum of"
echo "== HKF_PRE_V2=1"
HKF_PRE_V2=1 timeout 60 python tools/stage_times.py 2>&1 | grep -E "preprocess|icp_track|raycast_tracked|integrate_tracked|sum of"
echo "== HKF_INT_SPAN=3"
HKF_INT_SPAN=3 timeout 60 python tools/stage_times.py 2>&1 | grep -E "icp_track|raycast_tracked|integrate_tracked|sum of"
for cs in 8 16; do
    echo "== HKF_ICP_CLUSTER=$cs"
    HKF_ICP_CLUSTER=$cs timeout 60 python tools/stage_times.py 2>&1 | grep -E "icp_track|raycast_tracked|integrate_tracked|sum of"
done
echo "== all switches (cluster 16)"
HKF_RAY_SKIP=1 HKF_PRE_V2=1 HKF_INT_SPAN=3 HKF_ICP_CLUSTER=16 timeout 60 python tools/stage_times.py 2>&1 | grep -E "preprocess|icp_track|raycast_tracked|integrate_tracked|sum of"
HKF_RAY_SKIP=1 HKF_PRE_V2=1 HKF_INT_SPAN=3 HKF_ICP_CLUSTER=16 timeout 120 python bench.py --no-cpu-baseline 2>/dev/null | tail -c 700

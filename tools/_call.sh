set -x
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_pipeline.py tests/test_gpu_slabs_one_device.py -m gpu -x -q 2>&1 | tail -4
echo "== default"; timeout 120 python tools/stage_times.py 2>&1 | grep -E "preprocess|icp_track|raycast_tracked|integrate_tracked|sum of|N_upd"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_integrate -s 16 -c 2 -o gpurun_out/r2_prof_integrate_b python tools/prof_run.py 512 12 > gpurun_out/r2_ncu_b.log 2>&1; tail -2 gpurun_out/r2_ncu_b.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-c4-single > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_c.json').read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms_mean'], d['roofline'].get('steady'))"

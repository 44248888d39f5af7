set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_c4_n8_v3.json 2> gpurun_out/r2_c4_n8_v3.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c4_n8_v3.json').read().strip().splitlines()[-1])
print(8, d['value'], d['e2e']['value'], d['shard']['stage_us_per_rank'], d['shard']['owned_planes_per_rank'], d['shard']['other_icp_mode'])
PY
grep "c4 balance" gpurun_out/r2_c4_n8_v3.err; tail -2 gpurun_out/r2_c4_n8_v3.err
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2_pytest_gpu_sharded_n8_push.log

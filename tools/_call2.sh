set -x
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu_sharded_n8.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r2_c4_n$n.json 2> gpurun_out/r2_c4_n$n.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c4_n$n.json').read().strip().splitlines()[-1])
print($n, d['value'], d['e2e']['value'], d['shard']['stage_us_per_rank'], d['shard'].get('rank0_composite_us'), d['shard']['amdahl'])
PY
done
HKF_SHARD_ICP=replicated timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_c4_n8_repl.json 2> gpurun_out/r2_c4_n8_repl.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r2_c4_n8_repl.json').read().strip().splitlines()[-1])
print('repl', d['value'], d['e2e']['value'], d['shard']['stage_us_per_rank'])
PY

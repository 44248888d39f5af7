set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -s 2>&1 | tail -15 | tee gpurun_out/r2_pytest_gpu_sharded_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_c4_n2.json 2> gpurun_out/r2_c4_n2.err; tail -c 3000 gpurun_out/r2_c4_n2.json; tail -5 gpurun_out/r2_c4_n2.err
HKF_SHARD_ICP=replicated timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2_c4_n2_repl.json 2> gpurun_out/r2_c4_n2_repl.err; tail -c 1500 gpurun_out/r2_c4_n2_repl.json

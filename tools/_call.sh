set -x
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; tail -c 600 gpurun_out/r2_bench_e.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_e_ref.json 2> gpurun_out/r2_bench_e_ref.err; cat gpurun_out/r2_bench_e_ref.json | cut -c1-1500

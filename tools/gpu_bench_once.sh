mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3_r02a.json 2> gpurun_out/bench_c3_r02a.err; tail -c 3500 gpurun_out/bench_c3_r02a.json; tail -3 gpurun_out/bench_c3_r02a.err

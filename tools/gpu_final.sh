# What the driver runs at round end, on one B200: GPU test tier, smoke, bench (both arms).  gpurun --timeout 1200 -- "bash tools/gpu_final.sh r02"
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log
tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$tag.log; tail -4 gpurun_out/smoke_$tag.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; tail -c 600 gpurun_out/bench_ref_$tag.json
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err; tail -c 2500 gpurun_out/bench_c3_$tag.json

# Single-GPU measurement set of a round: gpurun --timeout 2400 -- "bash tools/gpu_measure.sh"; then python tools/summarize_ncu.py rNN
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python tools/prof_kernels.py 1000 1000 3 > gpurun_out/prof_kernels_c3.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 2600 gpurun_out/bench_c3.json
timeout 600 python bench.py --workload c2 --no-cpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -c 1200 gpurun_out/bench_c2.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 800 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
for k in edge_bwd2_kernel edge_fwd3_kernel mlp3_bwd_tc_kernel mlp3_fwd2_tc_kernel segment_sum_batch_kernel node_gemm_tc_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r01_full_$k python tools/prof_kernels.py 1000 1000 1 > gpurun_out/ncu_$k.log 2>&1
done
timeout 300 python tools/bench_train_step.py > gpurun_out/train_step.md 2> gpurun_out/train_step.err; tail -12 gpurun_out/train_step.md
timeout 300 python bench.py --workload c1 --no-cpu --steps 5 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; tail -c 600 gpurun_out/bench_c1.json
ls -la gpurun_out | head -30

# Single-GPU measurement set of a round: gpurun --timeout 2400 -- "bash tools/gpu_measure.sh r02"; then python tools/summarize_ncu.py r02
tag=${1:-r02}
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt; nproc >> gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$tag.log
tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$tag.log 2>&1; tail -3 gpurun_out/smoke_$tag.log
timeout 300 python tools/prof_kernels.py 1000 1000 15 > gpurun_out/prof_kernels_c3_$tag.txt 2>&1
timeout 900 python bench.py > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err; tail -c 1500 gpurun_out/bench_c3_$tag.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; tail -c 900 gpurun_out/bench_ref_$tag.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
# one capture per launch form the model makes (MGN_PROF_ONLY runs just that function of tools/prof_kernels.py; the regex
# names its kernel, so the set-up launches of the script are not what gets captured)
for kv in bwd2_dst:edge_bwd2 eblk_h1:edge_fwd3 nodefwd:mlp3_fwd2 bwd:mlp3_bwd csr:segment_sum lin_p:node_gemm wgrad:wgrad_tc; do
  k=${kv%%:*}; rx=${kv##*:}
  MGN_PROF_ONLY=$k timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -f -o gpurun_out/${tag}_full_$k python tools/prof_kernels.py 1000 1000 1 > gpurun_out/ncu_$k.log 2>&1
done
timeout 600 python tools/bench_c5.py > gpurun_out/c5_$tag.md 2> gpurun_out/c5_$tag.err; tail -5 gpurun_out/c5_$tag.md
timeout 300 python tools/bench_train_step.py > gpurun_out/train_step_$tag.md 2> gpurun_out/train_step_$tag.err; tail -12 gpurun_out/train_step_$tag.md
timeout 300 python bench.py --workload c1 --no-cpu --no-extra --steps 5 > gpurun_out/bench_c1_$tag.json 2> gpurun_out/bench_c1_$tag.err; tail -c 600 gpurun_out/bench_c1_$tag.json
ls -la gpurun_out | wc -l

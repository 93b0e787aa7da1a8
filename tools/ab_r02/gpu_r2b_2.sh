# gpurun --timeout 900 -- "bash tools/gpu_r2b_2.sh": per-CTA cycles / nanoseconds of the two edge kernels (debug build)
mkdir -p gpurun_out
MGN_NVCC_EXTRA="-DMGN_DEBUG_HOOKS" timeout 300 python -m modulus_b200.build > /dev/null || echo BUILD FAILED
(nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_throttle_reasons.active --format=csv -lms 200 > gpurun_out/r2b2_smi.csv &) 
MGN_NVCC_EXTRA="-DMGN_DEBUG_HOOKS" timeout 300 python tools/prof_kernels.py 1000 1000 15 > gpurun_out/r2b2.txt 2>&1
grep -E "eblk|bwd edge|per-CTA|FWD3|BWD2" gpurun_out/r2b2.txt | cut -c1-260
sort gpurun_out/r2b2_smi.csv | uniq -c | sort -rn | head -12
timeout 300 python -m modulus_b200.build > /dev/null

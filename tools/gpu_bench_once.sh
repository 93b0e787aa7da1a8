# One full single-GPU bench line + stand-alone kernel timings: gpurun --timeout 1500 -- "bash tools/gpu_bench_once.sh TAG"
tag=${1:-x}
mkdir -p gpurun_out
timeout 300 python tools/prof_kernels.py 1000 1000 15 2>&1 | head -12 | tee gpurun_out/prof_kernels_$tag.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c3_$tag.json 2> gpurun_out/bench_c3_$tag.err
python - <<PY
import json
d=[json.loads(l) for l in open("gpurun_out/bench_c3_$tag.json") if l.startswith("{")][-1]
print("c2 graph:", d["other_configs"]["c2"].get("cuda_graph")); print("c3:", round(d["ms_per_step"],2), "ms/step", round(d["value"]/1e6,2), "M edges/s | e2e", round(d["e2e"]["ms_per_step"],2), "| hbm_frac", round(d["whole_step"]["hbm_frac"],3), "| parity ok:", d["parity"]["ok"], "| c2:", round(d["other_configs"]["c2"]["ms_per_step"],2), "ms", round(d["other_configs"]["c2"]["value"]/1e6,1))
print("shares:", d["kernel_shares"])
print("roofline:", {k: d["roofline"][k] for k in ("kernel","frac","avg_launch_ms","share_of_step")}, "clocks", d["clocks"])
PY
tail -2 gpurun_out/bench_c3_$tag.err

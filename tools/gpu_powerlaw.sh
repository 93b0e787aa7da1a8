# gpurun --timeout 900 -- "bash tools/gpu_powerlaw.sh": hub threshold A/B of the segmented sum on the power-law graph
mkdir -p gpurun_out
for v in "" "-DMGN_LONG_SEG=128" "-DMGN_LONG_SEG=512"; do
  echo "=== variant '$v'"
  MGN_NVCC_EXTRA="$v" timeout 300 python -m modulus_b200.build > /dev/null || { echo BUILD FAILED; continue; }
  MGN_NVCC_EXTRA="$v" timeout 300 python tools/prof_powerlaw.py 128 10 2>&1 | tail -6
done
timeout 300 python -m modulus_b200.build > /dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/powerlaw_launches.csv python tools/prof_powerlaw.py 128 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open("gpurun_out/powerlaw_launches.csv") if not l.startswith("==")))
ix = {h: i for i, h in enumerate(rows[0])}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) > ix["Metric Value"] and r[ix["Metric Name"]] == "gpu__time_duration.sum":
        d = agg.setdefault(r[ix["Kernel Name"]][:70], [])
        d.append(float(r[ix["Metric Value"]].replace(",", "")) / 1e3)
for k, v in agg.items():
    if "segment" in k or "gather" in k or "copy" in k.lower():
        print(f"{k:70s} n={len(v):3d} last us: {[round(x, 1) for x in v[-4:]]}")
PY
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_rollout.py -m gpu -x -q 2>&1 | tail -3

"""Summarise gpurun_out ncu artefacts into profiles/ (tracked): launch list shares, per-kernel ncu --set full metrics,
DRAM traffic per launch (read back by bench.py as roofline.traffic)."""
import csv, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

def launch_list(path, title, out_md):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", "")); unit = r[ix["Metric Unit"]]
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        d = agg.setdefault(r[ix["Kernel Name"]], [0, 0.0]); d[0] += 1; d[1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(out_md, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are cold-cache and "
                f"serialised: compare shares, not absolutes.\n\ntotal {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches\n\n"
                "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
            f.write(f"| `{k[:100]}` | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |\n")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active"]

def full(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        if h in WANT or h == "Kernel Name":
            d[h] = (v, u)
    return d

def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

if __name__ == "__main__":
    ll = os.path.join(OUT, "launches_c3.csv")
    if os.path.exists(ll):
        launch_list(ll, f"ncu launch list, `bench.py --steps 1 --warmup 3 --no-cpu` (c3, {tag})", os.path.join(PROF, f"{tag}_launches_c3.md"))
    traffic = {"c3": {}}
    # capture name (tools/gpu_measure.sh: MGN_PROF_ONLY=<function of tools/prof_kernels.py>) -> C-ABI entry point
    entry = {"bwd2_dst": "mgn_edge_block_bwd_tc", "eblk_h1": "mgn_edge_block_fwd_tc", "nodefwd": "mgn_node_block_fwd_tc",
             "bwd": "mgn_mlp3_bwd_tc", "csr": "mgn_segment_sum", "lin_p": "mgn_node_gemm_tc", "wgrad": "mgn_wgrad_tc"}
    with open(os.path.join(PROF, f"{tag}_ncu_full_summary.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, one launch each at the c3 size (MGN_PROF_ONLY=<fn> tools/prof_kernels.py 1000 1000 1: the launch the model makes per layer), {tag}\n\n")
        for k, sym in entry.items():
            rep = os.path.join(OUT, f"{tag}_full_{k}.ncu-rep")
            if not os.path.exists(rep):
                continue
            d = full(rep)
            f.write(f"## {d.get('Kernel Name', (k,))[0][:120]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for h in WANT:
                if h in d:
                    f.write(f"| {h} | {d[h][0]} | {d[h][1]} |\n")
            rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
            f.write(f"| DRAM traffic per launch | {(rd + wr) / 1e9:.3f} | GB |\n\n")
            traffic["c3"][sym] = rd + wr
    traffic["note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at the c3 size, the form the model launches per "
                       "layer (edge kernels: 5 992 002 edge rows incl. the fused destination sums / the stored h1; node kernels: "
                       "1 000 000 rows), ncu --set full; bench.py reports it as roofline.traffic")
    json.dump(traffic, open(os.path.join(PROF, f"{tag}_ncu_traffic.json"), "w"), indent=1)
    print(open(os.path.join(PROF, f"{tag}_ncu_full_summary.md")).read())

"""Summarise `ncu --set full` reports: per launch duration, DRAM read + write bytes, tensor-pipe activity, and write
profiles/<name>.txt; with --traffic also profiles/r2_ncu_traffic.json ({bench kernel tag: dram bytes of ONE launch}) that
bench.py reads for `roofline.traffic`.  Usage: python tools/ncu_traffic.py out_name rep1.ncu-rep [rep2 ...] [--traffic]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
name, reps = args[0], args[1:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}
lines, traffic = [], {}
for rep in reps:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        kname = r[ix["Kernel Name"]]
        def val(k):
            if k not in ix or r[ix[k]] == "":
                return None
            return float(r[ix[k]].replace(",", "")) * UNIT.get(units[ix[k]], 1)
        ms, rd, wr = val("gpu__time_duration.sum"), val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        lines.append(f"{os.path.basename(rep)}: {kname[:60]}  grid {r[ix['launch__grid_size']]} x {r[ix['launch__block_size']]}  regs {r[ix['launch__registers_per_thread']]}")
        lines.append(f"    duration {ms:.4f} ms   dram read {rd / 1e9:.3f} GB  write {wr / 1e9:.3f} GB  -> {(rd + wr) / ms / 1e9:.2f} TB/s under ncu (cold caches, serialised)")
        for k in want[3:8]:
            if k in ix:
                lines.append(f"    {k} = {r[ix[k]]} {units[ix[k]]}")
        traffic.setdefault(kname.split("(")[0].split("::")[-1], []).append((ms, rd + wr))
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", name + ".txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
if "--traffic" in sys.argv:
    tags = {}
    fw = sorted(traffic.get("chain_fwd_ts2_kernel", []))          # by duration: coarse (64 samples) < fine (128 samples)
    if len(fw) >= 2:
        tags["chain_fwd_coarse"], tags["chain_fwd_fine"] = fw[0][1], fw[-1][1]
    for k, tag in (("trunk_bwd_kernel", "trunk_bwd"), ("fused_bwd_kernel", "fused_bwd heads1"), ("wgrad_kernel", "wgrad")):
        if k in traffic:
            tags[tag] = max(traffic[k])[1]
    json.dump(tags, open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w"), indent=1)
    print(tags)

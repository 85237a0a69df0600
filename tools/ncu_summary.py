#!/usr/bin/env python
"""Turn gpurun_out ncu artefacts into the small text summaries kept under profiles/.
usage: python tools/ncu_summary.py <tag>   (reads gpurun_out/launches.csv and gpurun_out/prof_fwd.ncu-rep)"""
import collections
import csv
import io
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
rep = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/prof_fwd.ncu-rep"
os.makedirs("profiles", exist_ok=True)

if os.path.exists("gpurun_out/launches.csv"):
    lines = [l for l in open("gpurun_out/launches.csv") if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else v * 1e3 if row["Metric Unit"] == "ms" else v
        agg.setdefault(row["Kernel Name"].split("(")[0][-70:], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(f"profiles/{tag}_launches.md", "w") as f:
        f.write(f"# ncu launch list ({tag}): `ncu --metrics gpu__time_duration.sum --clock-control none` on `python bench.py --steps 3 --warmup 3 --no-cpu`\n\n")
        f.write("Cold-cache, serialised per-launch times: compare shares, not absolutes.\n\n| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n")
        for k, v in agg.items():
            f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.2f} | {sum(v):.1f} | {100*sum(v)/tot:.1f}% |\n")
    print(open(f"profiles/{tag}_launches.md").read())

if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tensor_subpipe_hmma.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "smsp__cycles_active.avg"]
    with open(f"profiles/{tag}_{os.path.basename(rep).replace('.ncu-rep','')}_metrics.md", "w") as f:
        f.write(f"# ncu --set full --clock-control none ({tag}), `{os.path.basename(rep)}`, one row per captured launch\n\n")
        for r in rows[2:]:
            f.write("| metric | value | unit |\n|---|---|---|\n")
            for i, h in enumerate(hdr):
                base = h.split(".TriageCompute.")[-1]
                if any(base == w or h == w for w in want):
                    f.write(f"| {base} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    print(open(f"profiles/{tag}_{os.path.basename(rep).replace('.ncu-rep','')}_metrics.md").read()[:3000])

    # per-launch DRAM traffic of the conv kernel for bench.py's roofline.traffic (captured at the bench's default batch)
    import json
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in rows[2:] if "bk_forward_tc" in r[ik]]
    if vals:
        json.dump({"bk_forward_tc_kernel": {"batch": int(os.environ.get("BK_NCU_BATCH", "4096")), "dram_bytes_per_launch": sum(vals) / len(vals),
                                            "launches": len(vals), "source": f"profiles/{tag}_{os.path.basename(rep).replace('.ncu-rep','')}_metrics.md"}},
                  open("profiles/ncu_traffic.json", "w"), indent=1)

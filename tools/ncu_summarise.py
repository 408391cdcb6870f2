"""Summarise `ncu --set full` captures (.ncu-rep, read here with `ncu -i ... --page raw --csv`) into
profiles/rNN_kernel_summaries.json (read by bench.py for `roofline.traffic`) and a Markdown table.

usage: python tools/ncu_summarise.py OUT_PREFIX rep1.ncu-rep[:note[:algorithmic_bytes]] [rep2.ncu-rep...]
Per captured kernel (the launch with the longest duration per kernel name is kept):
  duration_us, dram_bytes (= dram__bytes_read.sum + dram__bytes_write.sum), achieved DRAM GB/s, tensor / fp64 pipe activity,
  registers, grid/block, occupancy, L2->SM throughput."""
import csv
import io
import json
import os
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
WANT = {
    "duration_us": "gpu__time_duration.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "dram_pct_of_peak": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pipe_pct_elapsed": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "tensor_pipe_pct_active": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "fp64_pipe_pct_elapsed": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "registers": "launch__registers_per_thread",
    "occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smem_per_block": "launch__shared_mem_per_block_dynamic",
    "issue_active_pct": "sm__inst_issued.avg.pct_of_peak_sustained_active" ,
}


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("unnamed>::", "").replace("gpk::", "").replace("(anonymous namespace)::", "")
    return name.strip()


def read(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": short(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for k, m in WANT.items():
            if m in col and r[col[m]] != "":
                v = float(r[col[m]].replace(",", ""))
                d[k] = v * UNIT.get(units[col[m]], 1.0) if k in ("duration_us", "dram_read", "dram_write") else v
        if "dram_read" in d:
            d["dram_bytes"] = d.pop("dram_read") + d.pop("dram_write")
            d["dram_gbs"] = d["dram_bytes"] / d["duration_us"] / 1e3
        res.append(d)
    return res


def main():
    prefix, reps = sys.argv[1], sys.argv[2:]
    summary = {}
    for item in reps:
        parts = item.split(":")
        rep, note = parts[0], (parts[1] if len(parts) > 1 else "")
        alg = float(parts[2]) if len(parts) > 2 and parts[2] else None
        for d in read(rep):
            if alg is not None:
                d["algorithmic_bytes"] = alg
            d["source"] = f"ncu --set full --clock-control none, {os.path.basename(rep)}" + (f" ({note})" if note else "")
            k = d["kernel"]
            if k not in summary or d.get("duration_us", 0) > summary[k].get("duration_us", 0):
                summary[k] = d
    path = prefix + "_kernel_summaries.json"
    old = {}
    if os.path.exists(path):
        old = json.load(open(path))
    old.update(summary)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    with open(prefix + "_kernel_summaries.md", "w") as fh:
        fh.write("| kernel | grid x block | regs | duration us | DRAM MB | DRAM GB/s (% of peak) | tensor pipe % elapsed | fp64 pipe % | SM % | occupancy % |\n")
        fh.write("|---|---|---|---|---|---|---|---|---|---|\n")
        for k, d in sorted(old.items()):
            g = lambda key, f="{:.1f}": f.format(d[key]) if key in d else "-"
            fh.write(f"| `{k}` | {d.get('grid')} x {d.get('block')} | {g('registers', '{:.0f}')} | {g('duration_us')} | "
                     f"{d.get('dram_bytes', 0) / 1e6:.1f} | {g('dram_gbs', '{:.0f}')} ({g('dram_pct_of_peak')}) | {g('tensor_pipe_pct_elapsed')} | "
                     f"{g('fp64_pipe_pct_elapsed')} | {g('sm_throughput_pct')} | {g('occupancy_pct')} |\n")
    print(json.dumps(summary, indent=1))


if __name__ == "__main__":
    main()

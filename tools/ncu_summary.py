"""Summarise an .ncu-rep (raw page) into the few metrics the roofline discussion needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [out.json]
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_bytes.sum", "sm__cycles_active.avg",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        item = {"kernel": d.get("Kernel Name"), "id": d.get("ID")}
        for k in KEYS:
            if k in d:
                item[k] = "%s %s" % (d[k], units[hdr.index(k)])
        stalls = {}
        for k in hdr:
            if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued"):
                try:
                    stalls[k[len("smsp__pcsamp_warps_issue_stalled_"):]] = int(float(d[k].replace(",", "")))
                except ValueError:
                    pass
        item["stall_samples"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        out.append(item)
    text = json.dumps(out, indent=1)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text + "\n")


if __name__ == "__main__":
    main()

"""Summarise `ncu --set full` captures into the small JSON / markdown files committed under profiles/.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > gpurun_out/X_raw.csv
    python tools/ncu_summary.py gpurun_out/X_raw.csv [more.csv ...] --json profiles/r02_force_ncu.json --md profiles/r02_ncu_summary.md

One entry per distinct kernel (the LAST launch of it in the capture).  bench.py reads `dram_bytes_read` / `dram_bytes_write` of
the dominant kernel from the JSON (roofline.traffic); nothing here is a bench value - ncu serialises kernels and flushes caches.
"""
import argparse
import csv
import json

KEYS = {
    "time_us": "gpu__time_duration.sum",
    "dram_bytes_read": "dram__bytes_read.sum",
    "dram_bytes_write": "dram__bytes_write.sum",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_throughput_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "registers": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "dyn_smem_kb": "launch__shared_mem_per_block_dynamic",
    "warp_inst": "smsp__inst_executed.sum",
    "threads_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
    "global_ld_requests": "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "global_ld_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "smem_ld_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "smem_ld_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "stall_long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "stall_short_scoreboard": "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "stall_barrier": "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "stall_mio_throttle": "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "stall_not_selected": "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "stall_wait": "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "cycles": "sm__cycles_elapsed.max",
}
UNIT_SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        e = {"name": name, "source": path}
        for k, m in KEYS.items():
            if m in d and d[m] != "":
                v = float(d[m].replace(",", ""))
                u = units[hdr.index(m)]
                if k in ("dram_bytes_read", "dram_bytes_write", "time_us") and u in UNIT_SCALE:
                    v *= UNIT_SCALE[u]
                e[k] = v
        out[name.split("(")[0]] = e          # last launch of each kernel wins
    return list(out.values())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv", nargs="+")
    ap.add_argument("--json")
    ap.add_argument("--md")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    kernels = []
    for p in a.csv:
        kernels += load(p)
    doc = {"what": "ncu --set full --clock-control none, one entry per kernel (last launch in the capture); dram bytes per launch",
           "note": a.note, "kernels": kernels}
    if a.json:
        json.dump(doc, open(a.json, "w"), indent=1)
    if a.md:
        with open(a.md, "w") as f:
            f.write("# ncu summary\n\n%s\n\n" % a.note)
            cols = ["time_us", "dram_bytes_read", "dram_bytes_write", "issue_active_pct", "l1tex_throughput_pct", "warps_active_pct",
                    "registers", "warp_inst", "smem_ld_wavefronts", "smem_ld_bank_conflicts", "stall_long_scoreboard", "stall_barrier"]
            f.write("| kernel | " + " | ".join(cols) + " |\n|---|" + "---|" * len(cols) + "\n")
            for k in kernels:
                f.write("| `%s` | " % k["name"].split("(")[0][:60] + " | ".join(
                    ("%.4g" % k[c]) if c in k else "-" for c in cols) + " |\n")
    for k in kernels:
        print("%-60s %8.2f us  dram %6.1f MB  issue %4.1f%%  L1 %4.1f%%  warps %4.1f%%" % (
            k["name"].split("(")[0][:60], k.get("time_us", 0), (k.get("dram_bytes_read", 0) + k.get("dram_bytes_write", 0)) / 1e6,
            k.get("issue_active_pct", 0), k.get("l1tex_throughput_pct", 0), k.get("warps_active_pct", 0)))


if __name__ == "__main__":
    main()

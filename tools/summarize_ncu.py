"""Turn the ncu artefacts a gpurun call brings back into the small tracked summaries under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/<tag>_launches.csv profiles/<name>.csv
        per-launch list (id, kernel, grid, block, gpu__time_duration in us) of a `--metrics gpu__time_duration.sum` pass
    python tools/summarize_ncu.py full gpurun_out/<tag>_full.ncu-rep profiles/<name>.json
        key metrics of every captured launch of a `--set full` report (read with `ncu -i ... --page raw --csv`)
"""
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "memory_l1_wavefronts_shared_ideal",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    ix = {k: j for j, k in enumerate(rows[h])}
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "grid", "block", "gpu_time_us"])
        for r in rows[h + 1:]:
            if len(r) < len(rows[h]) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
                continue
            v, u = float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]]
            v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v * 1e6 if u == "s" else v
            w.writerow([r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0][:80], r[ix["Grid Size"]].replace(" ", ""),
                        r[ix["Block Size"]].replace(" ", ""), "%.2f" % v])


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {k: j for j, k in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        e = {"kernel": r[ix["Kernel Name"]].split("(")[0][:80], "id": r[ix["ID"]]}
        for k in KEYS:
            if k in ix and r[ix[k]] not in ("", "no data"):
                e[k] = ("%s %s" % (r[ix[k]], units[ix[k]])).strip()
        out.append(e)
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])

"""Turn raw ncu output (gpurun_out/, scratch) into the committed summaries under profiles/.

    python profiles/summarize.py launches <launches.csv> <out.txt> "<command that was profiled>"
    python profiles/summarize.py full <report.ncu-rep> <out.txt>
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct"]


def launches(path, out, cmd):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    with open(out, "w") as f:
        f.write("# ncu launch list summary; per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"# command: {cmd}\n")
        f.write(f"total_us {T:.1f} launches {sum(cnt.values())}\n")
        for k, v in tot.most_common(60):
            f.write(f"{v / T * 100:6.2f}% {v:12.1f}us n={cnt[k]:6d} avg={v / cnt[k]:9.2f}us  {k[:110]}\n")


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; selected metrics per captured launch ({rep})\n")
        for r in rows[2:]:
            f.write(f"\n== {r[hdr.index('Kernel Name')][:150]}\n")
            for k in KEYS:
                if k in hdr:
                    f.write(f"{k:70s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])

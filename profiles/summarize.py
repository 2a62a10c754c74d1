"""Turn the ncu exports under gpurun_out/ into the small, committed summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv profiles/r01_launches_<tag>.md
  python profiles/summarize.py raw gpurun_out/gemm.ncu-rep profiles/r01_gemm_<tag>.md [traffic.json]
"""
import collections
import csv
import json
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", "")) * {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — per-launch times are cold-cache and "
                "serialised: compare SHARES, not absolutes.\n\n")
        f.write(f"total kernel time {T:.3f} ms over {sum(cnt.values())} launches\n\n| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"| {v:.3f} | {100 * v / T:.1f}% | {cnt[k]} | `{k[:110]}` |\n")
    print("wrote", dst)


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem"]


def raw(rep, dst, traffic_json=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ci = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({rep})\n\n`ncu --set full --clock-control none --import-source on`; one column per captured launch.\n\n")
        f.write("| metric | " + " | ".join(f"launch {i}" for i in range(len(data))) + " | unit |\n|---|" + "---:|" * len(data) + "---|\n")
        f.write("| kernel | " + " | ".join(r[ci["Kernel Name"]][:40] for r in data) + " | |\n")
        for k in KEYS:
            if k in ci:
                f.write(f"| {k} | " + " | ".join(r[ci[k]] for r in data) + f" | {units[ci[k]]} |\n")
    if traffic_json:
        tr = []
        for r in data:
            u = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
            rd = float(r[ci["dram__bytes_read.sum"]].replace(",", "")) * u[units[ci["dram__bytes_read.sum"]]]
            wr = float(r[ci["dram__bytes_write.sum"]].replace(",", "")) * u[units[ci["dram__bytes_write.sum"]]]
            tr.append(rd + wr)
        json.dump({"dram_bytes_per_launch": sum(tr) / len(tr), "per_launch": tr, "source": rep}, open(traffic_json, "w"))
    print("wrote", dst)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        raw(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)

"""Summarises ncu outputs brought back from the GPU box into small text files under profiles/.
  python tools/ncu_summary.py launches gpurun_out/launches.csv            -> per-kernel totals / shares
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep                -> key metrics per captured launch"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = row["Metric Unit"]
        ns = v * 1e3 if unit.startswith("us") else (v * 1e6 if unit.startswith("ms") else v)
        name = row["Kernel Name"]
        key = re.sub(r"<.*", "", name.split("(")[0]).replace("void ", "").replace("b200::", "")
        m = re.search(r"igemm_kernel<(\d+)", name)
        if m:
            key = f"igemm_kernel<BN={m.group(1)}>"
        m = re.search(r"attn_kernel<(\d+)", name)
        if m:
            key = f"attn_kernel<D={m.group(1)}>"
        agg[key][0] += 1
        agg[key][1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"# per-kernel device time of ONE eager SDXL UNet forward (B=8, 1024^2), ncu --metrics gpu__time_duration.sum")
    print(f"# (cold-cache, serialised launches: compare SHARES, not absolutes). total {tot / 1e6:.3f} ms, {sum(v[0] for v in agg.values())} launches")
    fam = collections.defaultdict(float)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:40s} n={v[0]:5d} {v[1] / 1e6:9.3f} ms {100 * v[1] / tot:6.2f}%")
        fam["igemm" if k.startswith("igemm") else "attention" if k.startswith("attn") else "layernorm" if "layernorm" in k
            else "groupnorm" if k.startswith("gn_") else "other"] += v[1]
    print("# by family: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])))


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__cluster_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "gpc__cycles_elapsed.avg.per_second",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full --clock-control none, {path}")
    for r in rows[2:]:
        print("----", r[idx["Kernel Name"]][:110])
        for w in WANT:
            if w in idx:
                print(f"  {w:78s} {r[idx[w]][:24]:>24s} {units[idx[w]]}")


def traffic(path):
    """csv of `--metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:igemm` -> JSON for bench.py's roofline.traffic"""
    import json
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.defaultdict(float)
    for row in csv.DictReader(lines):
        if not row["Metric Name"].startswith("dram__bytes"):
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
        per[row["ID"]] += v * mult
    n = len(per)
    tot = sum(per.values())
    print(json.dumps({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:igemm, one eager SDXL forward "
                                "(B=8, 1024^2)", "launches": n, "dram_bytes_total": tot, "dram_bytes_per_launch": round(tot / n)}))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])

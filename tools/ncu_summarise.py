"""Turn ncu outputs (launch-list csv, .ncu-rep) into small text summaries kept under profiles/."""
import collections
import csv
import subprocess
import sys


def launch_table(path, top=30):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[row["Metric Unit"]]
        a = agg.setdefault(row["Kernel Name"][:90], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    out = [f"total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches (ncu: serialised, cold cache; compare SHARES)"]
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        out.append(f"{v:9.3f} ms {100 * v / tot:5.1f}%  x{n:<4d} {k}")
    return "\n".join(out)


KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def rep_metrics(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        out.append(f"kernel: {name}")
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS:
                out.append(f"  {h} = {v} {u}")
    return "\n".join(out)


def sass_mix(path, units=None, top=24):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    by, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in rows[2:]:
        try:
            n = int(r[ix["Instructions Executed"]])
        except (ValueError, IndexError):
            continue
        src = r[ix["Source"]].strip().split()
        op = src[1] if src and src[0].startswith("@") else (src[0] if src else "?")
        op = ".".join(op.split(".")[:3])
        by[op] += n
        tot += n
        samp[op] += int(r[ix["# Samples"]] or 0)
    st = sum(samp.values()) or 1
    out = [f"warp-instructions executed: {tot}" + (f"  ({tot / units:.1f} per work unit)" if units else "")]
    for op, n in by.most_common(top):
        out.append(f"  {op:28s} {100 * n / tot:5.1f}% of instr   {100 * samp[op] / st:5.1f}% of stall samples"
                   + (f"   {n / units:7.2f}/unit" if units else ""))
    return "\n".join(out)


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    if mode == "launches":
        print(launch_table(path))
    elif mode == "rep":
        print(rep_metrics(path))
        units = float(sys.argv[3]) if len(sys.argv) > 3 else None
        print(sass_mix(path, units))

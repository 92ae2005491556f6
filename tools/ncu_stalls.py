"""Summarise an ncu report: pipe utilisation + the instructions / source lines where warps stall.
    python tools/ncu_stalls.py report.ncu-rep [top_n]"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[2] if len(rows) > 2 else rows[1]
    pat = re.compile(r"(gpu__time_duration\.sum$|sm__inst_executed_pipe_(xu|alu|fma|fmaheavy|lsu|tensor\w*|uniform|tmem|tma)\.avg\.pct_of_peak_sustained_active"
                     r"|sm__issue_active\.avg\.pct|sm__warps_active\.avg\.pct|dram__bytes_(read|write)\.sum$|launch__registers_per_thread$"
                     r"|lts__t_bytes\.sum$|sm__pipe_tensor_cycles_active\w*\.avg\.pct_of_peak_sustained_active|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$"
                     r"|local_op_(ld|st)\.sum$|smsp__cycles_active\.avg$|sm__cycles_active\.avg$|lts__t_sectors_srcunit_tex_op_read\.sum$)")
    for h, v in zip(hdr, vals):
        if pat.search(h):
            print(f"{h} = {v}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print(f"\ntotal samples {tot}")
    agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
        print(f"  {k:26s} {100 * v / tot:5.1f}%")
    print()
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:top]:
        st = {s: int(r[ix[s]] or 0) for s in stalls}
        main_s = max(st, key=st.get)
        print(f"{r[ix['Address']][-6:]} {int(r[ix['# Samples']]):7d} {100 * int(r[ix['# Samples']]) / tot:5.1f}% {r[ix['Instructions Executed']]:>10s} {main_s:22s} {r[ix['Source']][:80]}")


if __name__ == "__main__":
    main()

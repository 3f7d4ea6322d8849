"""Summarise an `ncu --page raw --csv` export: per-launch table (markdown) and per-kernel-instance
average DRAM traffic (json, consumed by bench.py's roofline.traffic)."""
import csv, json, re, sys
from collections import OrderedDict

raw, out_md, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def g(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(",", ""))
    except Exception:
        return default
def unit(name):
    return units[col[name]] if name in col else ""
def to_bytes(v, u):
    u = u.lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u.lower(), 1)
lines = ["| # | kernel | time us | DRAM rd MB | DRAM wr MB | DRAM GB/s | tensor pipe % | L2 hit % | regs |", "|---|---|---|---|---|---|---|---|---|"]
agg = OrderedDict()
for i, r in enumerate(data):
    name = r[col["Kernel Name"]]
    m = re.search(r"conv3x3_tc_kernel<(?:\(int\))?(\d+), (?:\(int\))?(\d+), (?:\(bool\))?(\d), (?:\(bool\))?(\d), (?:\(int\))?(\d+)(?:, (?:\(int\))?(\d+))?(?:, (?:\(int\))?(\d+))?>", name)
    short = (f"conv3x3_tc_kernel<{m.group(1)},{m.group(2)}> cta2={m.group(4)} mask={m.group(5)} pipe={m.group(7)}"
             if m else name.split("(")[0][-40:])
    key = f"conv3x3_tc_kernel<{m.group(1)},{m.group(2)}>" if m else ("first_conv_kernel" if "first_conv" in name else "final_conv_kernel" if "final_conv" in name else short)
    t = to_us(g(r, "gpu__time_duration.sum"), unit("gpu__time_duration.sum"))
    rd = to_bytes(g(r, "dram__bytes_read.sum"), unit("dram__bytes_read.sum"))
    wr = to_bytes(g(r, "dram__bytes_write.sum"), unit("dram__bytes_write.sum"))
    tp = g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    hit = g(r, "lts__t_sector_hit_rate.pct")
    regs = g(r, "launch__registers_per_thread")
    lines.append(f"| {i} | {short} | {t:.1f} | {rd/1e6:.1f} | {wr/1e6:.1f} | {(rd+wr)/t/1e3:.0f} | {tp:.1f} | {hit:.1f} | {regs:.0f} |")
    a = agg.setdefault(key, {"bytes": 0.0, "n": 0, "us": 0.0})
    a["bytes"] += rd + wr; a["n"] += 1; a["us"] += t
open(out_md, "w").write("\n".join(lines) + "\n")
json.dump({k: v["bytes"] / v["n"] for k, v in agg.items()}, open(out_json, "w"), indent=1)
print("\n".join(lines[:6])); print({k: (round(v["bytes"] / v["n"] / 1e6, 1), v["n"], round(v["us"], 1)) for k, v in agg.items()})

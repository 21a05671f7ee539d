"""Summarise an ncu launch list (csv of `--metrics gpu__time_duration.sum`) into per-kernel totals and shares.
usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/<name>_summary.txt"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
seq = []
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("udb::", "").strip()
    v = float(r[iv].replace(",", ""))
    us = v / 1000.0 if r[iu] in ("ns", "nsecond") else (v if r[iu] in ("us", "usecond") else v * 1000.0)
    t = tot.setdefault(name, [0.0, 0])
    t[0] += us
    t[1] += 1
    seq.append((name, us))
total = sum(t[0] for t in tot.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none, one eager infer step (tools/profile_step.py"
      + (" " + " ".join(sys.argv[2:]) if len(sys.argv) > 2 else "") + ")")
print(f"total {total:.1f} us over {len(seq)} launches (cold-cache, serialised: compare shares)")
for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{us:10.1f} us {100 * us / total:5.1f}% x{n:4d}  {name}")
first = next((i for i, (n, _) in enumerate(seq) if n.startswith("void attn_fwd") or n.startswith("attn_fwd")), None)
if first is not None:
    print("first encoder block:")
    for n, us in seq[max(0, first - 2):first + 5]:
        print(f"{us:11.1f} us  {n}")
if len(sys.argv) > 2 and sys.argv[-1] == "--seq":
    print("launch sequence:")
    for n, us in seq:
        print(f"{us:11.1f} us  {n}")

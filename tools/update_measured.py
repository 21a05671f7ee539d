"""Rewrite the MEASURED table of tests/test_infer_parity_gpu.py from a `pytest -m gpu -s` log (lines "PARITY <tag>: ...").
usage: python tools/update_measured.py gpurun_out/r02_parity_gpu.log     (prints the old and new values per tag)"""
import re
import sys

log = open(sys.argv[1]).read()
vals = {}
for m in re.finditer(r"PARITY (\S+): depth ARel ([0-9.e+-]+) max ([0-9.e+-]+); intrinsics rel ([0-9.e+-]+)", log):
    tag, a, d, k = m.group(1), float(m.group(2)), float(m.group(3)), float(m.group(4))
    if log[max(0, m.start() - 2):m.start()].endswith("V1"):
        continue
    o = vals.get(tag)
    vals[tag] = (max(a, o[0]), max(d, o[1]), max(k, o[2])) if o else (a, d, k)
path = "tests/test_infer_parity_gpu.py"
src = open(path).read()
for tag, (a, d, k) in vals.items():
    pat = re.compile(r'(\s+"%s": )\(([0-9.e+-]+), ([0-9.e+-]+), ([0-9.e+-]+)\),' % re.escape(tag))
    m = pat.search(src)
    if not m:
        continue
    print(f"{tag:34s} {m.group(2)} {m.group(3)} {m.group(4)}  ->  {a:.3e} {d:.3e} {k:.3e}")
    src = src[:m.start()] + f'{m.group(1)}({a:.3e}, {d:.3e}, {k:.3e}),' + src[m.end():]
open(path, "w").write(src)

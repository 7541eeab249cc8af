#!/usr/bin/env python
"""Per-source-line view of an ncu capture (development aid): instructions executed and warp-stall samples of one kernel,
aggregated by the CUDA source line each SASS instruction belongs to (nvdisasm -g line info of the SAME library build).

usage: python tools/ncu_lines.py capture.ncu-rep k_decide2 [--top 40] [--lib kuberay_b200/libkrengine.so]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("kernel"); ap.add_argument("--top", type=int, default=40)
ap.add_argument("--lib", default="kuberay_b200/libkrengine.so")
a = ap.parse_args()

raw = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--kernel-name", f"regex:{a.kernel}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
kname = rows[0][1]
hdr = rows[1]
ia, ist, iex = hdr.index("Address"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
inst = []
for r in rows[2:]:
    if len(r) <= iex or not r[ia].startswith("0x") and not r[ia].isdigit():
        if len(inst) and len(r) > 1 and r[0] == "Kernel Name":
            break  # next launch of the same kernel
        continue
    try:
        inst.append((int(r[ia], 0), int(r[ist] or 0), int(r[iex] or 0), r[hdr.index("Source")]))
    except ValueError:
        pass
base = min(x[0] for x in inst)
mangled = None
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.lib)], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin") and "specjson" not in f][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
# find the function whose demangled name matches: use the template args in the ncu name
want = re.sub(r"\W+", "", re.sub(r"\((int|bool)\)", "", kname).split("(")[0].replace("void kr::", ""))
line_of, cur, on = {}, ("?", 0), False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        sym = ln[6:].rstrip(":")
        dem = subprocess.run(["c++filt", sym], capture_output=True, text=True).stdout.strip()
        dem = dem.replace("false", "0").replace("true", "1")  # ncu prints bool template arguments as (bool)0 / (bool)1
        on = re.sub(r"\W+", "", dem.split("(")[0].replace("void kr::", "").replace("kr::", "", 1)) == want
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
agg = collections.defaultdict(lambda: [0, 0])
tot_s = tot_i = 0
for addr, st, ex, _src in inst:
    key = line_of.get(addr - base, ("?", 0))
    agg[key][0] += st; agg[key][1] += ex
    tot_s += st; tot_i += ex
print(f"{kname}\n  {tot_i} warp instructions, {tot_s} stall samples, {len(inst)} SASS instructions")
src_cache = {}
def text(f, n):
    if f not in src_cache:
        p = os.path.join("kuberay_b200/csrc", f)
        src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
    L = src_cache[f]
    return L[n - 1].strip()[:100] if 0 < n <= len(L) else ""
for key, (st, ex) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:a.top]:
    print(f"  {100 * ex / max(tot_i, 1):5.1f}% instr  {100 * st / max(tot_s, 1):5.1f}% stall  {key[0]}:{key[1]:<4d} {text(*key)}")

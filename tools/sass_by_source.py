#!/usr/bin/env python
"""Folds an ncu SASS page onto source lines: which source function / line range the executed warp-instructions belong to.

usage: python tools/sass_by_source.py <report.ncu-rep> <library.so> <mangled kernel name> [--lines]
Needs nvdisasm / cuobjdump (line info from -lineinfo).  The i-th SASS instruction of the kernel's .text section in the
cubin is the i-th row of ncu's listing (ncu appends the out-of-line device functions the kernel calls; they are matched
the same way through their own sections in call order when the counts agree, otherwise reported as 'callee')."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, lib, kern = sys.argv[1:4]
by_line = "--lines" in sys.argv
tmp = tempfile.mkdtemp()
subprocess.run(f"cd {tmp} && cuobjdump -xelf all {os.path.abspath(lib)} > /dev/null", shell=True, check=True)
sass = None
for c in os.listdir(tmp):
    if c.endswith(".cubin"):
        out = subprocess.run(f"nvdisasm -g {tmp}/{c}", shell=True, capture_output=True, text=True).stdout
        if f".text.{kern}:" in out:
            sass = out
            break
assert sass, "kernel not found"
lines = sass.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(f".text.{kern}:"))
cur = ("?", 0)
ins = []   # (file, line, inlined-chain, text)
chain = ""
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((cur[0], cur[1], m.group(2)))
src = subprocess.run(f"ncu -i {rep} --page source --csv --print-source sass", shell=True, capture_output=True, text=True).stdout
rs = list(csv.reader(src.splitlines()))
h = rs[1]; ix = {k: i for i, k in enumerate(h)}
body = [r for r in rs[2:] if len(r) > 10]
n = min(len(ins), len(body))
print(f"# {len(ins)} SASS instructions in the kernel's section, {len(body)} rows in the ncu listing (the rest: out-of-line callees)")
# source functions: line -> function name, from the .cuh/.cu files
funcs = {}
def load(fn):
    path = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", fn)
    if not os.path.exists(path): return
    name = "?"
    tbl = []
    for i, l in enumerate(open(path), 1):
        m = re.match(r"^(?:static )?__(?:device|global)__.*?([A-Za-z_0-9]+)\s*\(", l) or re.match(r"^__device__ __forceinline__ \S+ ([A-Za-z_0-9]+)\(", l)
        if m and not l.startswith(" "):
            name = m.group(1)
        else:   # member functions of the small structs (Col::load / store / scr ...)
            m2 = re.match(r"^\s+__device__ __forceinline__ [^(]*?([A-Za-z_0-9]+)\s*\(", l)
            if m2:
                name = "member:" + m2.group(1)
        tbl.append(name)
    funcs[fn] = tbl
agg = collections.OrderedDict()
tot = sum(int(r[ix["Instructions Executed"]]) for r in body)
for i, r in enumerate(body):
    ex = int(r[ix["Instructions Executed"]]); th = int(r[ix["Thread Instructions Executed"]]); sm = int(r[ix["# Samples"]])
    if i < n:
        f, ln, _ = ins[i]
        if f not in funcs: load(f)
        fn = funcs.get(f, ["?"] * 100000)
        key = (f, ln) if by_line else (f, fn[ln - 1] if 0 < ln <= len(fn) else "?")
    else:
        key = ("callee", "out-of-line")
    a = agg.setdefault(key, [0, 0, 0, 0])
    a[0] += ex; a[1] += th; a[2] += sm; a[3] += 1
print("file,function_or_line,sass_instructions,exec_share_pct,avg_lanes,sample_share_pct")
tsm = sum(a[2] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if a[0] == 0: continue
    print("%s,%s,%d,%.2f,%.1f,%.2f" % (k[0], k[1], a[3], 100.0 * a[0] / tot, a[1] / max(a[0], 1), 100.0 * a[2] / max(tsm, 1)))

if "--layout" in sys.argv:
    # contiguous runs of the kernel's SASS by source function: where the hot code lies
    print("# layout: first_index,last_index,function,exec_share_pct,avg_lanes")
    runs = []
    for i, r in enumerate(body[:n]):
        f, ln, _ = ins[i]
        fn = funcs.get(f, ["?"] * 100000)
        name = fn[ln - 1] if 0 < ln <= len(fn) else "?"
        ex = int(r[ix["Instructions Executed"]]); th = int(r[ix["Thread Instructions Executed"]])
        if runs and runs[-1][2] == name:
            runs[-1][1] = i; runs[-1][3] += ex; runs[-1][4] += th
        else:
            runs.append([i, i, name, ex, th])
    # merge short runs into 64-instruction windows for readability
    for a, b, name, ex, th in runs:
        if b - a + 1 >= 8 or 100.0 * ex / tot >= 0.3:
            print("%d,%d,%s,%.2f,%.1f" % (a, b, name, 100.0 * ex / tot, th / max(ex, 1)))

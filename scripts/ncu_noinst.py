#!/usr/bin/env python
"""Where do instruction-fetch stalls (stall_no_inst) land?  Top SASS instructions by no_inst samples, plus a
histogram of executed instructions and no_inst samples over 4 KB code pages."""
import csv, os, re, subprocess, sys, tempfile, collections
rep, obj, ksub = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern = None; hdr = None; rows = {}
for row in csv.reader(sass.splitlines()):
    if not row: continue
    if row[0] == "Kernel Name": kern = row[1]; rows[kern] = []; hdr = None; continue
    if row[0] == "Address": hdr = row; continue
    if hdr: rows[kern].append(row)
iE = hdr.index("Instructions Executed"); iN = hdr.index("stall_no_inst"); iS = hdr.index("# Samples")
for kname, rs in rows.items():
    if ksub not in kname: continue
    funcs = {}; cur = None; line = ("?", 0)
    for l in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m: cur = m.group(1); funcs[cur] = []; continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*);", l)
        if m and cur: funcs[cur].append((int(m.group(1), 16), line, m.group(2).strip()))
    c = [f for f, ins in funcs.items() if len(ins) == len(rs)]
    if not c: continue
    ins = funcs[c[0]]
    print("==", kname[:80])
    tot_n = sum(int(r[iN] or 0) for r in rs); tot_e = sum(int(r[iE]) for r in rs)
    top = sorted(((int(r[iN] or 0), int(r[iE]), off, ln, txt) for (off, ln, txt), r in zip(ins, rs)), reverse=True)[:40]
    for n, e, off, ln, txt in top:
        print(f"  {n:6d} no_inst  exec {e:9d}  @{off:06x}  {ln[0]}:{ln[1]}  {txt[:60]}")
    pages = collections.defaultdict(lambda: [0, 0])
    for (off, ln, txt), r in zip(ins, rs):
        pg = off >> 12
        pages[pg][0] += int(r[iE]); pages[pg][1] += int(r[iN] or 0)
    print("  4 KB page: exec%  no_inst%")
    for pg in sorted(pages):
        e, n = pages[pg]
        if e / tot_e > 0.003 or n / max(tot_n, 1) > 0.003:
            print(f"   {pg << 12:06x}: {100 * e / tot_e:5.1f}  {100 * n / max(tot_n, 1):5.1f}")

#!/usr/bin/env python
"""List the hottest CALL sites (IEEE div/sqrt slow paths) of a kernel in an ncu report, with source lines."""
import csv, os, re, subprocess, sys, tempfile
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
iE = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed")
for kname, rs in rows.items():
    if ksub not in kname: continue
    funcs = {}; cur = None; line = ("?", 0)
    for l in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m: cur = m.group(1); funcs[cur] = []; continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*);", l)
        if m and cur: funcs[cur].append((line, m.group(2)))
    c = [f for f, ins in funcs.items() if len(ins) == len(rs)]
    if not c: continue
    out = []
    for (ln, txt), r in zip(funcs[c[0]], rs):
        if txt.strip().startswith("CALL") or "CALL" in txt.split()[0:2]:
            out.append((int(r[iE]), int(r[iT]), ln, txt.strip()[:70]))
    out.sort(reverse=True)
    print("==", kname[:80])
    for e, t, ln, txt in out[:25]:
        print(f"  {e:9d} warps  {t/max(e,1):5.1f} thr  {ln[0]}:{ln[1]}  {txt}")

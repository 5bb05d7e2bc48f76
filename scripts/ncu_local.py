#!/usr/bin/env python
"""Which source lines execute local-memory loads / stores (LDL / STL) in a kernel of an ncu report.
    python scripts/ncu_local.py <prof.ncu-rep> <object.o> <kernel-substring>"""
import csv, os, re, subprocess, sys, tempfile, collections
rep, obj, ksub = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern = None; hdr = None; rows = collections.OrderedDict()
for row in csv.reader(sass.splitlines()):
    if not row: continue
    if row[0] == "Kernel Name": kern = row[1]; rows[kern] = []; hdr = None; continue
    if row[0] == "Address": hdr = row; continue
    if hdr: rows[kern].append(dict(zip(hdr, row)))
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
    print("==", kname[:90], "matching:", len(c))
    if not c: continue
    agg = collections.defaultdict(lambda: [0, 0])
    for (ln, txt), r in zip(funcs[c[0]], rs):
        op = txt.split()[0] if not txt.startswith("@") else txt.split()[1]
        e = int(r["Instructions Executed"])
        if op.startswith("STL"): agg[ln][0] += e
        if op.startswith("LDL"): agg[ln][1] += e
    for ln, (st, ld) in sorted(agg.items(), key=lambda kv: -(kv[1][0] + kv[1][1]))[:25]:
        print(f"  STL {st:12d}  LDL {ld:12d}  {ln[0]}:{ln[1]}")

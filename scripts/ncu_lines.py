#!/usr/bin/env python
"""Join an ncu report's per-SASS-instruction counters with the source lines of the cubin (nvdisasm
--print-line-info), and print the hottest source lines per kernel.
    python scripts/ncu_lines.py <prof.ncu-rep> <object.o> <kernel-substring> [top]
Run here (no GPU needed)."""
import csv, os, re, subprocess, sys, tempfile, collections

rep, obj, ksub = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
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
    # mangled function in the disassembly with the same number of instructions
    funcs = {}; cur = None; line = ("?", 0)
    for l in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m: cur = m.group(1); funcs[cur] = []; continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*);", l)
        if m and cur: funcs[cur].append((int(m.group(1), 16), line, m.group(2)))
    cands = [f for f, ins in funcs.items() if len(ins) == len(rs)]
    print("==", kname[:90], len(rs), "instructions; matching functions:", len(cands))
    if not cands: continue
    ins = funcs[cands[0]]
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot_e = tot_s = 0
    for (off, ln, txt), r in zip(ins, rs):
        e = int(r["Instructions Executed"]); s = int(r["# Samples"]); ni = int(r.get("stall_no_inst") or 0)
        a = agg[ln]; a[0] += e; a[1] += s; a[2] += 1; a[3] += ni
        tot_e += e; tot_s += s
    print(f"total executed warp-inst {tot_e:.3e}, samples {tot_s}")
    print("  exec%  samp%  noinst  #sass  file:line")
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {100*a[0]/tot_e:5.1f}  {100*a[1]/max(tot_s,1):5.1f}  {a[3]:6d}  {a[2]:5d}  {ln[0]}:{ln[1]}")
    byfile = collections.defaultdict(lambda: [0, 0])
    for ln, a in agg.items(): byfile[ln[0]][0] += a[0]; byfile[ln[0]][1] += a[1]
    for f, a in byfile.items(): print(f"  file {f}: exec {100*a[0]/tot_e:.1f}% samples {100*a[1]/max(tot_s,1):.1f}%")

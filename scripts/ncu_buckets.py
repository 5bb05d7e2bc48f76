#!/usr/bin/env python
"""Attribute an ncu report's per-instruction counters to FUNCTIONS through the inline chains of the cubin
(nvdisasm -gi): every SASS instruction is charged to the outermost-listed bucket that appears in its chain.
    python scripts/ncu_buckets.py <prof.ncu-rep> <object.o> <kernel-substring>
Run here (no GPU needed)."""
import csv, os, re, subprocess, sys, tempfile, collections

BUCKETS = [  # (name, file, first line, last line) — checked innermost first
    ("fold intersects_x", "d2d_trace.cuh", 139, 199),
    ("hit_exact", "d2d_device.cuh", 307, 315),
    ("path_loss", "d2d_trace.cuh", 111, 125),
    ("image_path_on", "d2d_trace.cuh", 39, 80),
    ("trace_image_tracked", "d2d_image_bwd.cuh", 51, 126),
    ("image_reverse", "d2d_image_bwd.cuh", 128, 320),
    ("path_value", "d2d_trace.cuh", 201, 207),
    ("validity_from_onx", "d2d_trace.cuh", 209, 258),
    ("tile_may_be_valid", "d2d_driver.cuh", 253, 394),
    ("tile_may_be_valid_tx", "d2d_driver.cuh", 411, 528),
    ("warp_may_be_valid", "d2d_driver.cuh", 537, 559),
    ("test_candidate", "d2d_driver.cuh", 561, 642),
    ("macro_prologue", "d2d_driver.cuh", 677, 722),
    ("mask_prologue/bitmap", "d2d_driver.cuh", 653, 675),
    ("mask_prologue/bitmap", "d2d_driver.cuh", 724, 738),
    ("chunk driver", "d2d_driver.cuh", 740, 898),
    ("make_tile", "d2d_driver.cuh", 133, 220),
    ("build_tab", "d2d_device.cuh", 109, 144),
    ("visit (forward.cu)", "d2d_forward.cu", 17, 56),
    ("visit (backward.cu)", "d2d_backward.cu", 316, 416),
    ("kernel body", "d2d_forward.cu", 57, 200),
    ("kernel body", "d2d_backward.cu", 417, 620),
]

rep, obj, ksub = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kern = None; hdr = None; rows = collections.OrderedDict()
for row in csv.reader(sass.splitlines()):
    if not row: continue
    if row[0] == "Kernel Name": kern = row[1]; rows[kern] = []; hdr = None; continue
    if row[0] == "Address": hdr = row; continue
    if hdr: rows[kern].append(dict(zip(hdr, row)))

def bucket(chain, sub):
    for f, ln in chain:  # innermost first
        for name, bf, a, b in BUCKETS:
            if f == bf and a <= ln <= b: return name
    return sub or "other"

for kname, rs in rows.items():
    if ksub not in kname: continue
    funcs = {}; cur = None; chain = []; pending = []; sub = None
    for l in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m: cur = m.group(1); funcs[cur] = []; sub = None; continue
        m = re.match(r"\s*\.weak\s+(\$\S+)|^(\$\S+):", l)
        if m and cur:
            nm = (m.group(1) or m.group(2))
            mm = re.search(r"\$_ZN3d2d(\d+)(\w+)", nm)
            sub = "sub:" + (mm.group(2)[:int(mm.group(1))] if mm else nm[-30:])
            if "__internal" in nm or "$__" in nm: sub = "sub:" + nm.split("$")[-1][:40]
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m: pending.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*);", l)
        if m and cur:
            if pending: chain = pending; pending = []
            funcs[cur].append((int(m.group(1), 16), chain, m.group(2), sub))
    cands = [f for f, ins in funcs.items() if len(ins) == len(rs)]
    print("==", kname[:100], len(rs), "instructions; matching functions:", len(cands))
    if not cands: continue
    ins = funcs[cands[0]]
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
    visits = max(int(r["Instructions Executed"]) for r in rs) ; hot_thr = float(os.environ.get("HOT_FRAC", "0.02")) * visits
    tot_e = tot_s = 0
    for (off, ch, txt, sb), r in zip(ins, rs):
        e = int(r["Instructions Executed"]); s = int(r["# Samples"]); ni = int(r.get("stall_no_inst") or 0)
        te = int(r["Thread Instructions Executed"])
        a = agg[bucket(ch, sb)]; a[0] += e; a[1] += s; a[2] += 1; a[3] += ni; a[4] += te; a[5] += (e >= hot_thr)
        tot_e += e; tot_s += s
    print(f"total executed warp-inst {tot_e:.3e}, samples {tot_s}")
    print("  exec%  samp%  noinst%  lanes  #sass  #hot  bucket")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {100*a[0]/tot_e:5.1f}  {100*a[1]/max(tot_s,1):5.1f}  {100*a[3]/max(a[1],1):6.1f}  {a[4]/max(a[0],1):5.1f}  {a[2]:5d}  {a[5]:5d}  {name}")

#!/usr/bin/env python
"""Attribute an ncu report's per-instruction counters to FUNCTIONS through the inline chains of the cubin
(nvdisasm -gi): every SASS instruction is charged to the outermost-listed bucket that appears in its chain.
    python scripts/ncu_buckets.py <prof.ncu-rep> <object.o> <kernel-substring>
Run here (no GPU needed)."""
import csv, os, re, subprocess, sys, tempfile, collections

# bucket -> (file, function names whose bodies belong to it); line ranges are read from the CURRENT sources
FUNCS = [
    ("fold intersects_x", "d2d_trace.cuh", ["intersects_x"]),
    ("hit_exact", "d2d_device.cuh", ["hit_exact"]),
    ("path_loss", "d2d_trace.cuh", ["path_loss", "path_loss_dirs"]),
    ("normalize2 / residual", "d2d_device.cuh", ["normalize2", "residual_dirs", "residual"]),
    ("back_project / to_parametric / mirror", "d2d_device.cuh", ["back_project", "to_parametric", "mirror"]),
    ("act", "d2d_device.cuh", ["act", "act_is_zero", "act_is_one", "act_dz", "fold_skip_bound", "fold_start", "fold_skip_of"]),
    ("image_path_on", "d2d_trace.cuh", ["image_path_on", "image_path", "image_apex"]),
    ("trace_image_tracked", "d2d_image_bwd.cuh", ["trace_image_tracked"]),
    ("image_reverse", "d2d_image_bwd.cuh", ["image_reverse_general", "image_reverse", "act_and_dz", "dz_from_act", "fdiv"]),
    ("adjoint pieces", "d2d_adjoint.cuh", ["normalize_adj", "residual_adj", "normalize_adj_hat", "residual_adj_dirs", "dot2", "to_vertices"]),
    ("path_value / path_length", "d2d_trace.cuh", ["path_value"]),
    ("path_value / path_length", "d2d_device.cuh", ["path_length"]),
    ("validity_from_onx", "d2d_trace.cuh", ["validity_from_onx", "validity", "on_objects_x"]),
    ("tile_may_be_valid", "d2d_driver.cuh", ["tile_may_be_valid", "pow2_floor"]),
    ("tile_may_be_valid_tx", "d2d_driver.cuh", ["tile_may_be_valid_tx"]),
    ("warp_may_be_valid", "d2d_driver.cuh", ["warp_may_be_valid"]),
    ("test_candidate", "d2d_driver.cuh", ["test_candidate_inline", "test_candidate"]),
    ("macro_prologue", "d2d_driver.cuh", ["macro_prologue"]),
    ("mask_prologue/bitmap", "d2d_driver.cuh", ["bitmap_prefix", "mask_prologue", "mask_bitmap_fits"]),
    ("chunk driver", "d2d_driver.cuh", ["for_each_candidate", "order_count", "hint_slot"]),
    ("make_tile", "d2d_driver.cuh", ["make_tile"]),
    ("build_tab", "d2d_device.cuh", ["build_tab", "carve_tab"]),
    ("visit (forward.cu)", "d2d_forward.cu", ["run_order"]),
    ("visit + reductions (backward.cu)", "d2d_backward.cu", ["run_order_bwd", "warp_sum", "warp_sum_many", "path_vjp"]),
    ("kernel body", "d2d_forward.cu", ["power_fwd_kernel"]),
    ("kernel body", "d2d_backward.cu", ["power_bwd_kernel"]),
]


def _ranges():
    """(name, file, first, last) for every listed function: from its signature line to the line before the next
    top-level definition (a line starting with `template`, `__device__`, `__global__`, `struct`, `static`, `inline`, `//`
    at column 0 after a closing brace at column 0)."""
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "differt2d_b200", "csrc")
    out = []
    cache = {}
    for name, f, fns in FUNCS:
        if f not in cache:
            cache[f] = open(os.path.join(here, f)).read().splitlines()
        lines = cache[f]
        for fn in fns:
            pat = re.compile(r"\b" + re.escape(fn) + r"\s*\(")
            for i, l in enumerate(lines):
                if not pat.search(l):
                    continue
                # a definition: the signature is at nesting depth 0 and a '{' follows before a ';'
                if not re.match(r"^(template|__device__|__global__|static|inline|__host__|\s{0,4}__device__)", l) and not (i > 0 and re.match(r"^(template|__device__|__global__|static)", lines[i - 1])):
                    continue
                j = i
                while j < len(lines) and "{" not in lines[j] and ";" not in lines[j]:
                    j += 1
                if j >= len(lines) or ";" in lines[j].split("{")[0]:
                    continue  # a declaration
                # body ends at the first line that is exactly "}" (column 0) or "    }" for members
                indent = len(l) - len(l.lstrip())
                k = j
                while k < len(lines) and lines[k].rstrip() != " " * indent + "}":
                    k += 1
                out.append((name, f, i + 1, min(k + 1, len(lines))))
    return out


BUCKETS = _ranges()

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

"""Instruction attribution of one ncu --set full --import-source on capture of k_backward: executed warp-instructions by
source region (LDLT fast step / general step / second solve / helpers / KKT assembly ...) and by opcode.
    python tools/ncu_regions.py <report.ncu-rep> <model> <kernel mangled substring>"""
import csv, re, sys, subprocess, os, tempfile
rep, model, kern = sys.argv[1:4]
tmp = tempfile.mkdtemp()
subprocess.run(f"ncu -i {rep} --page source --csv > {tmp}/src.csv 2>/dev/null", shell=True)
subprocess.run(f"cd {tmp} && cuobjdump -xelf {model} /root/repo/interiorpointddp.jl_b200/libipddp_b200.so >/dev/null 2>&1 && nvdisasm -g {model}.sm_100a.cubin > all.sass 2>/dev/null", shell=True)
lines = open(f"{tmp}/all.sass").read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
cur = ("?", 0); seq = []
for l in lines[start + 1:]:
    if l.startswith("//-----") or l.startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((int(m.group(1), 16), cur, m.group(2)))
rows = list(csv.reader(open(f"{tmp}/src.csv")))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]; ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples"); ca = hdr.index("Address")
data = []
for r in rows[hi + 1:]:
    try: data.append((int(r[ca], 16), int(r[ci]), int(r[cs]), r[1]))
    except Exception: pass
base = data[0][0]
off2 = {o: (c, t) for o, c, t in seq}
tot = sum(d[1] for d in data)
# regions: boundaries located in the sources (first line containing each marker), so edits do not shift the attribution
CS = "/root/repo/interiorpointddp.jl_b200/csrc/"
def first_line(path, marker):
    for i, l in enumerate(open(CS + path), 1):
        if marker in l: return i
    raise SystemExit(f"marker {marker!r} not found in {path}")
LD = [(first_line("ldlt_warp.cuh", m), name) for m, name in (
    ("IPDDP_D int coff(", "ldlt helpers (coff, DivBy, max_ties, swaps, compaction)"),
    ("// Access to the K x NR right-hand sides", "right-hand-side row updates (RhsL: first solve loop, interchanges)"),
    ("// Scratch used by the factorisation", "ldlt scratch accessors"),
    ("// General pivot step (column k)", "ldlt_step (general)"),
    ("// Fast pivot step for 0 <= k < 32", "ldlt_step_fast"),
    ("// The same fast step for a pivot column 32 <= k < 64", "ldlt_step_fast2"),
    ("// dsytf2_rook('U') on the packed matrix A of order K", "factor driver"),
    ("// second loop of dsytrs_rook('U')", "second solve"))]
BW = [(first_line("kernel_backward.cuh", m), name) for m, name in (
    ("template <class M> struct BwLayout", "bw setup / terminal knot"),
    ("// ---- stage inputs", "bw assembly 1 (scatter, barrier terms)"),
    ("// ---- xx_tmp = fx' Vxx+", "bw assembly 2 (products, contractions, park)"),
    ("// ---- factorise + inertia", "bw factor call + ineq gains"),
    ("// ---- Vxx = beta' B + omega' cx + C", "bw value update + write back"),
    ("IPDDP_D int bw_sweep(", "bw sweep driver / kernel"))]
def region(f, ln):
    for tab, name in (("ldlt_warp.cuh", LD), ("kernel_backward.cuh", BW)):
        if f == tab:
            r = "preamble"
            for start, nm in name:
                if ln >= start: r = nm
            return r
    return f
agg = {}; ops = {}
for a, n, s, txt in data:
    (f, ln), t = off2.get(a - base, (("?", 0), "?"))
    e = agg.setdefault(region(f, ln), [0, 0]); e[0] += n; e[1] += s
    op = t.split()[0] if not t.startswith("@") else t.split()[1]
    op = op.split(".")[0]
    ops[op] = ops.get(op, 0) + n
print("total", f"{tot:.3e}")
for k, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]): print(f"{100*n/tot:5.1f}%  {k}")
print("--- opcodes")
for k, n in sorted(ops.items(), key=lambda kv: -kv[1])[:22]: print(f"{100*n/tot:5.1f}%  {k}")

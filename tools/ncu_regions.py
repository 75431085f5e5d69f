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
# regions
def region(f, ln):
    if f == "ldlt_warp.cuh":
        if ln < 171: return "ldlt helpers (coff, DivBy, max_ties, swaps)"
        if ln < 385: return "ldlt_step (general)"
        if ln < 472: return "ldlt_step_fast"
        if ln < 572: return "ldlt_step_fast2"
        if ln < 605: return "factor driver"
        return "second solve"
    if f == "kernel_backward.cuh":
        if ln < 250: return "bw assembly 1 (scatter, barrier terms)"
        if ln < 320: return "bw assembly 2 (products, contractions, park)"
        if ln < 350: return "bw factor call + ineq gains"
        return "bw value update + write back"
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

"""Joins an ncu SASS-level source page (csv) with nvdisasm -g line info to attribute executed
instructions / stall samples to CUDA source lines.
usage: sass_by_line.py <ncu_source.csv> <nvdisasm_all.sass> <kernel mangled substring> [top]"""
import csv, re, sys
ncu_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# parse nvdisasm: sequence of instructions with current (file,line)
lines = open(sass).read().split("\n")
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l][0]
cur = ("?", 0); seq = []
for l in lines[start + 1:]:
    if l.startswith("//-----") or (l.startswith(".text.") ): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((int(m.group(1), 16), cur, m.group(2)))
rows = list(csv.reader(open(ncu_csv)))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]; ci = hdr.index("Instructions Executed"); cs = hdr.index("# Samples"); ca = hdr.index("Address")
data = []
for r in rows[hi + 1:]:
    try: data.append((int(r[ca], 16), int(r[ci]), int(r[cs]), r[1]))
    except Exception: pass
base = data[0][0]
agg = {}
tot_i = tot_s = 0
off2line = {o: c for o, c, _ in seq}
for a, n, s, txt in data:
    c = off2line.get(a - base, ("?", 0))
    k = c
    e = agg.setdefault(k, [0, 0]); e[0] += n; e[1] += s; tot_i += n; tot_s += s
print(f"total warp-instructions {tot_i:.3e}, samples {tot_s}")
src_cache = {}
def src(f, ln):
    import glob
    if f not in src_cache:
        g = glob.glob(f"/root/repo/interiorpointddp.jl_b200/csrc/**/{f}", recursive=True)
        src_cache[f] = open(g[0]).read().split("\n") if g else []
    L = src_cache[f]
    return L[ln - 1].strip()[:100] if 0 < ln <= len(L) else ""
for (f, ln), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/tot_i:5.1f}% inst {100*s/max(1,tot_s):5.1f}% stall  {f}:{ln:<4} {src(f, ln)}")

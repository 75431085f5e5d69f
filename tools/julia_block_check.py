"""Crude structural check of Julia sources (no Julia in this image): block openers vs `end`, bracket balance; strings and
comments stripped; comprehension `for`/`if`, `a[end]`, `:end` symbols handled.
    python tools/julia_block_check.py julia/*.jl"""
import re, sys
OPEN = {"function","if","for","while","struct","begin","let","do","try","module","quote","macro","baremodule"}
def strip(src):
    """remove comments, strings, chars; keep structure"""
    out=[]; i=0; n=len(src)
    while i<n:
        c=src[i]
        if src.startswith('#=',i):
            j=src.find('=#',i+2); i = n if j<0 else j+2; continue
        if c=='#':
            j=src.find('\n',i); i = n if j<0 else j; continue
        if src.startswith('"""',i):
            j=i+3
            while j<n and not src.startswith('"""',j):
                j+= 2 if src[j]=='\\' else 1
            i=j+3; out.append('""'); continue
        if c=='"':
            j=i+1
            while j<n and src[j]!='"':
                j+= 2 if src[j]=='\\' else 1
            i=j+1; out.append('""'); continue
        if c=="'" and i+2<n and (src[i+2]=="'" or (src[i+1]=='\\' and i+3<n and src[i+3]=="'")):
            i += 3 if src[i+2]=="'" else 4; out.append("' '"); continue
        out.append(c); i+=1
    return ''.join(out)
def check(path):
    s=strip(open(path).read())
    stack=[]; depth_br=0
    toks=re.finditer(r"[A-Za-z_@][A-Za-z_0-9!]*|[\[\]\(\)\{\}]|\n|:", s)
    prev=None; line=1; errs=[]
    br=[]
    for m in toks:
        t=m.group(0)
        if t=='\n': line+=1; prev=t; continue
        if t in '([{': br.append((t,line)); prev=t; continue
        if t in ')]}':
            if not br: errs.append(f"{path}:{line}: unmatched {t}"); prev=t; continue
            o,l=br.pop()
            if '([{'.index(o)!=')]}'.index(t): errs.append(f"{path}:{line}: {t} closes {o} from line {l}")
            prev=t; continue
        if t==':': prev=t; continue
        inside_index = any(o=='[' for o,_ in br)
        if prev==':' : prev=t; continue          # :end, :function symbols
        if t=='mutable' or t=='abstract' or t=='primitive': prev=t; continue
        if t=='struct' or t=='type':
            if t=='struct': stack.append((t,line))
            elif prev in ('abstract','primitive'): stack.append((t,line))
            prev=t; continue
        if t in OPEN:
            if t=='if' and prev=='else': prev=t; continue   # 'else if' is not Julia, ignore
            if t in ('for','if') and br and br[-1][0] in '([':   # comprehension / generator
                prev=t; continue
            stack.append((t,line))
        elif t=='end':
            if inside_index and br[-1][0]=='[': prev=t; continue   # a[end]
            if not stack: errs.append(f"{path}:{line}: 'end' without opener")
            else: stack.pop()
        prev=t
    for t,l in stack: errs.append(f"{path}:{l}: '{t}' never closed")
    for o,l in br: errs.append(f"{path}:{l}: '{o}' never closed")
    return errs
if __name__=="__main__":
    bad=0
    for p in sys.argv[1:]:
        e=check(p); bad+=len(e)
        print(p, "OK" if not e else "\n".join(e))
    sys.exit(1 if bad else 0)

"""CPU evidence script (no GPU): random model SHAPES -- nx 1..10 states, 1..4 force-like and 0..3 slack-like controls, 0..3
constraints, the four bound patterns -- through the whole solve on the emulator (plugins compiled from the emitted device
headers) against a scratch oracle carrying the same traced closures (tests/emu/user_model_harness.py), speculative and bulk
kernels, bit for bit.  The run that found the single-control bug of the gains' index arithmetic was of this kind
(tests/test_emu_model_shapes.py keeps the shapes that matter in the suite).

    python tests/tools/model_shape_fuzz.py <seed> <number of models>  >> profiles/r2_model_shape_fuzz.txt
"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for _p in (ROOT, os.path.join(ROOT, 'tests'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests', 'emu')):
    sys.path.insert(0, _p)
import numpy as np, sympy as sp
import helpers
import user_model_harness as H
from ipddp_b200.batch import BatchSolver
from ipddp_b200.codegen import generate, workloads
inf=float("inf"); dt=0.05
M=workloads.ModelDef
rng=np.random.default_rng(int(sys.argv[1]) if len(sys.argv)>1 else 0)
def make(name, nx, nf, ns, nc, bound_kind):
    # nf force-like controls, ns slack-like controls (>=0); nc constraints each using a slack if available
    nu=nf+ns
    A=rng.uniform(-0.5,0.5,(nx,nx)).round(2); Bm=rng.uniform(-1,1,(nx,nu)).round(2)*(rng.uniform(size=(nx,nu))<0.6)
    tgt=rng.uniform(-0.5,0.5,nx).round(2)
    def f(x,u,p):
        return [x[i]+dt*(sum(float(A[i,j])*x[j] for j in range(nx))+sum(float(Bm[i,j])*u[j] for j in range(nu) if Bm[i,j]!=0)
                + 0.1*sp.sin(x[(i+1)%nx])*u[i%nu]) for i in range(nx)]
    def stage(x,u,p):
        return dt*(sum(0.1*u[i]*u[i] for i in range(nf))+sum(1.5*u[nf+i] for i in range(ns))+0.01*sum(x[i]*x[i] for i in range(nx)))
    def term(x,p): return 10.0*sum((x[i]-float(tgt[i]))**2 for i in range(nx))
    def c(x,u,p):
        out=[]
        for i in range(nc):
            e = x[i%nx]*u[i%nf] + 0.02 + 0.1*u[(i+1)%nf]
            if ns>0: e = e - u[nf+(i%ns)]
            out.append(e)
        return out
    if bound_kind==0: lo=[-2.0]*nf; up=[2.0]*nf
    elif bound_kind==1: lo=[-inf]*nf; up=[2.0]*nf
    elif bound_kind==2: lo=[-2.0]*nf; up=[inf]*nf
    else: lo=[-inf]*nf; up=[inf]*nf
    return M(name=name,nx=nx,nu=nu,np_=0,f=f,stage_cost=stage,term_cost=term,c=c,lower=lambda p:lo+[0.0]*ns,upper=lambda p:up+[inf]*ns,
             u_init=[0.0]*nf+[0.01]*ns,dt=dt)
shapes=[]
for q in range(int(sys.argv[2]) if len(sys.argv)>2 else 16):
    nx=int(rng.integers(1,11)); nf=int(rng.integers(1,5)); ns=int(rng.integers(0,4)); nc=int(rng.integers(0,min(nf+ns,4)))
    if ns==0 and nc>=nf: nc=max(0,nf-1)
    shapes.append((f"fz{q}",nx,nf,ns,nc,int(rng.integers(0,4))))
mds=[make(*s_) for s_ in shapes]
t=time.time()
models=[(md, generate.trace(md)) for md in mds]
emu=H.emulator_with_models(models); orc=H.scratch_oracle(tempfile.mkdtemp(), models)
print("built", time.time()-t, flush=True)
bad=0
for (md,_),sh in zip(models,shapes):
    name=md.name; nx,nu,nc,np_,_s=emu.model_dims(name)
    for N in (7,):
        B,maxit=2,25
        x1=0.2*rng.standard_normal((B,nx)); ubar=np.tile(np.asarray(md.u_init,dtype=float),(B,N-1))
        P=np.zeros((B,0))
        lower=np.tile(np.asarray(md.lower([]),dtype=float),(B,1)); upper=np.tile(np.asarray(md.upper([]),dtype=float),(B,1))
        res,xo,uo=orc.solve_batch(name,N,P,lower,upper,x1,ubar,options=orc.default_options(optimality_tolerance=1e-7,max_iterations=maxit),want_traj=True)
        try:
            for spec in (-1,0):
                emu.L.ipddp_set_tuning(None,b"fw_spec_max",spec); emu.L.ipddp_set_tuning(None,b"bw_spec_max",spec)
                s=BatchSolver(name,B,N,options=emu.default_options(optimality_tolerance=1e-7,max_iterations=maxit),lib=emu)
                s.set_inputs(x1,ubar,None,lower,upper); r=s.solve(); x,u=s.trajectory(); s.close()
                emu.L.ipddp_set_tuning(None,b"fw_spec_max",-1); emu.L.ipddp_set_tuning(None,b"bw_spec_max",-1)
                for i in range(B):
                    o=res[i]; got=(int(r.status[i]),int(r.k[i]),int(r.j[i]),int(r.l[i]))
                    assert got==(o.status,o.k,o.j,o.l),(name,N,spec,i,got,(o.status,o.k,o.j,o.l))
                    helpers.assert_same_bits(r.objective[i],o.objective,"obj")
                helpers.assert_same_bits(x,xo,"x"); helpers.assert_same_bits(u,uo,"u")
            print(sh,(nx,nu,nc),[(r_.status,r_.k) for r_ in res],"OK",flush=True)
        except AssertionError as e:
            bad+=1; print(sh,"MISMATCH",str(e)[:300],flush=True)
print("mismatches",bad)

#!/usr/bin/env python
"""Structure of the LU patterns (mistra_b200/mech/*.json): loads / stores / multiply-adds of the
tile-blocked factorisation for several tile shapes, and the split of the work between head pivots
and the dense tail.  Used for the ceiling analysis in DESIGN.md 5.1.  Run from the repo root."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mistra_b200.mechgen import mech as mechmod
def count(name, tail, BR, BC):
    m=mechmod.load(name); n=m.nvar
    h=n-tail
    pos=m.pos
    fill=set(nz for nz in range(m.lu_nonzero) if not m.jvs[nz]) - set(int(d) for d in m.diag[:n])
    loads=0; stores=0; fmas=0
    # pass 0 (rows programs): approximate loads = entries touched + U loads per upd
    for k in range(n):
        lo,dg,hi=int(m.crow[k]),int(m.diag[k]),int(m.crow[k+1])
        touched=set()
        for kk in range(lo,dg):
            j=int(m.icol[kk])
            if k>=h and j>=h: break
            touched.add(kk); loads+=1  # pivot recip
            for jj in range(int(m.diag[j])+1,int(m.crow[j+1])):
                c=int(m.icol[jj])
                if k>=h and c>=h: continue
                touched.add(pos[(k,c)]); loads+=1; fmas+=1
        loads+=len([t for t in touched if t not in fill]); stores+=len(touched)
    p0=(loads,stores,fmas)
    ntr=(tail+BR-1)//BR; ntc=(tail+BC-1)//BC
    tl=0;ts=0;tf=0;maxacc=0
    for I in range(ntr):
        R=list(range(h+I*BR,min(n,h+I*BR+BR)))
        for J in range(ntc):
            Cc=list(range(h+J*BC,min(n,h+J*BC+BC)))
            ent={(i,j) for i in R for j in Cc if (i,j) in pos}
            if not ent: continue
            maxacc=max(maxacc,len(ent))
            tl+=len([e for e in ent if pos[e] not in fill]); ts+=len(ent)
            ks=set()
            for i in R:
                for c in range(int(m.crow[i]),int(m.diag[i])): ks.add(int(m.icol[c]))
            for k in sorted(ks):
                if k>=R[-1] or k>Cc[-1]: break
                rows_k=[i for i in R if i>k and (i,k) in pos]
                cols_k=[j for j in Cc if j>k and (k,j) in pos]
                if not rows_k: continue
                inrow = k in R; incol = k in Cc
                if incol: tl+=1  # pivot
                if not cols_k: continue
                if not incol: tl+=len(rows_k)
                if not inrow: tl+=len(cols_k)
                tf+=len(rows_k)*len(cols_k)
    return p0,(tl,ts,tf),maxacc
for name,tail in (("aer",96),("aer",64),("aer",128),("tot",128),("gas",32)):
    for BR,BC in ((8,8),(12,8),(8,12),(16,8),(8,16),(12,12),(16,16)):
        p0,t,ma=count(name,tail,BR,BC)
        print(name,tail,(BR,BC),"p0 loads/stores/fma",p0,"tiles loads/stores/fma",t,"total slots",p0[0]+p0[1]+t[0]+t[1],"maxacc",ma)

def split(name, tail, B=8):
    m=mechmod.load(name); n=m.nvar; h=n-tail; pos=m.pos
    nt=tail//B
    head_loads=head_f=tail_loads=tail_f=0
    for I in range(nt):
        R=list(range(h+I*B,h+I*B+B))
        for J in range(nt):
            Cc=list(range(h+J*B,h+J*B+B))
            ent={(i,j) for i in R for j in Cc if (i,j) in pos}
            if not ent: continue
            ks=set()
            for i in R:
                for c in range(int(m.crow[i]),int(m.diag[i])): ks.add(int(m.icol[c]))
            for k in sorted(ks):
                if k>=R[-1] or k>Cc[-1]: break
                rows_k=[i for i in R if i>k and (i,k) in pos]
                cols_k=[j for j in Cc if j>k and (k,j) in pos]
                if not rows_k or not cols_k: continue
                l=(0 if k in Cc else len(rows_k))+(0 if k in R else len(cols_k))
                f=len(rows_k)*len(cols_k)
                if k<h: head_loads+=l; head_f+=f
                else: tail_loads+=l; tail_f+=f
    print(name,tail,"head-pivot loads",head_loads,"fma",head_f,"| tail-pivot loads",tail_loads,"fma",tail_f)
print()
for t in (64,96,128): split("aer",t)

#!/usr/bin/env python
"""Developer experiment: would Jacobi sweeps restricted to occupied-virtual pairs make the intermediate
SCF solves cheaper?  Runs the caffeine SCF of the oracle with a scalar cyclic Jacobi in the previous iteration's eigenbasis
(the warm start of k_scf) and counts sweeps to max|off| <= 1e-9: full sweeps vs OV-only sweeps (cost 1/2 each).
Result: full 41 sweep-equivalents, OV-only 39.5 -- the restriction turns quadratic into linear convergence.

    python tools/jacobi_ov_sim.py [molecule]
"""
import sys, json
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
import numpy as np
from oracle import gfn1_oracle as O
from scipy.linalg import eigh, cholesky
import bench
mols=json.load(open(__import__('pathlib').Path(__file__).resolve().parent.parent / 'tests/golden/molecules.json'))
name=sys.argv[1] if len(sys.argv)>1 else 'caffeine'
z=np.array(mols[name]['numbers']); p0=np.array(mols[name]['positions'])
p=bench._perturb(p0,[3])[0]
m=O.make_mol(z); par=O.params()
S,_=O.overlap(m,p); cn,_=O.cn_d3(m,p); H0=O.h0(m,p,S,cn)
gam=O.gamma_shell(m,p); g3=O.gam3(m)
n0=O._shell_param(m,"refocc")[m.ao_sh]/(2*m.sh_l[m.ao_sh]+1)
nel=n0.sum(); nocc=int(round(nel/2)); n=m.nao
kt=300*par.kelvin2au

def sweep(A,V,pairs):
    for (i,j) in pairs:
        apq=A[i,j]
        if abs(apq)<1e-300: continue
        d=A[j,j]-A[i,i]
        t=np.sign(d)*2*apq/(abs(d)+np.sqrt(d*d+4*apq*apq)) if d!=0 else 1.0
        c=1/np.sqrt(1+t*t); s=t*c
        # rotate: new_i = c*i - s*j ; new_j = s*i + c*j
        Ai=A[:,i].copy(); Aj=A[:,j].copy()
        A[:,i]=c*Ai-s*Aj; A[:,j]=s*Ai+c*Aj
        Ai=A[i,:].copy(); Aj=A[j,:].copy()
        A[i,:]=c*Ai-s*Aj; A[j,:]=s*Ai+c*Aj
        Vi=V[:,i].copy(); Vj=V[:,j].copy()
        V[:,i]=c*Vi-s*Vj; V[:,j]=s*Vi+c*Vj
all_pairs=[(i,j) for i in range(n) for j in range(i+1,n)]

def run(mode):
    L=cholesky(S,lower=True); C=np.linalg.inv(L).T
    q0=O.guess_orbital_charges(m,O.eeq_charges(m,p,0.0))
    v,_,_=O._potential(m,q0,gam,g3)
    mixer=O.Anderson(n)
    tot=0; log=[]
    occ_idx=None
    for it in range(40):
        F=H0-0.5*S*(v[:,None]+v[None,:])
        A=C.T@F@C; A=0.5*(A+A.T)
        sw=0
        if occ_idx is None or mode=='full':
            while np.abs(A-np.diag(np.diag(A))).max()>1e-9:
                sweep(A,C,all_pairs); sw+=1
            order=np.argsort(np.diag(A)); occ_idx=order[:nocc]; vir_idx=order[nocc:]
            ov=[(min(i,a),max(i,a)) for i in occ_idx for a in vir_idx]
        else:
            while np.abs(A[np.ix_(occ_idx,vir_idx)]).max()>1e-9:
                sweep(A,C,ov); sw+=0.5
                if sw>30: break
        tot+=sw; log.append(sw)
        f=np.zeros(n); f[occ_idx]=2.0
        P=(C*f)@C.T
        q=n0-np.einsum('ik,ki->i',P,S)
        vnew,_,_=O._potential(m,q,gam,g3)
        vold=v
        v=mixer.iter(vnew,v)
        if it>0 and mixer.converged(1e-4,1e-5): break
    return it+1, tot, log
for mode in ('full','ov'):
    print(mode, run(mode))

#!/usr/bin/env python
"""Developer experiment behind the occupied-subspace solve of k_scf: runs the oracle's SCF for a
molecule twice in NumPy -- every iteration diagonalised by cyclic Jacobi in the previous eigenbasis (the round-1 kernel), and
with the policy of xtb_scf_subspace.cuh (sweeps only until the gap between the diagonal blocks is certified, then the Riccati
fixed point + Newton inverse) -- and prints sweeps / fixed-point / Newton iterations, the charge error of every intermediate
solve against scipy eigh, and a rough cost ratio.  caffeine: 42 sweeps -> 10 sweeps + 65 fixed-point iterations.

    python tools/subspace_sim.py caffeine,LYS_xao [-v]
"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from oracle import gfn1_oracle as O
from scipy.linalg import eigh, cholesky
import bench
mols=json.load(open(Path(__file__).resolve().parent.parent / 'tests/golden/molecules.json'))
KT=300*O.params().kelvin2au

def jacobi_sweep(A,V):
    n=A.shape[0]
    for i in range(n):
        for j in range(i+1,n):
            apq=A[i,j]
            if abs(apq)<1e-300: continue
            d=A[j,j]-A[i,i]
            t=np.sign(d)*2*apq/(abs(d)+np.sqrt(d*d+4*apq*apq)) if d!=0 else 1.0
            c=1/np.sqrt(1+t*t); s=t*c
            Ai=A[:,i].copy(); Aj=A[:,j].copy()
            A[:,i]=c*Ai-s*Aj; A[:,j]=s*Ai+c*Aj
            Ai=A[i,:].copy(); Aj=A[j,:].copy()
            A[i,:]=c*Ai-s*Aj; A[j,:]=s*Ai+c*Aj
            Vi=V[:,i].copy(); Vj=V[:,j].copy()
            V[:,i]=c*Vi-s*Vj; V[:,j]=s*Vi+c*Vj

def offmax(A): return np.abs(A-np.diag(np.diag(A))).max()

def run(name, seed, fast, rtol=1e-10, kcert=60.0, maxric=16, verbose=False, pos=None):
    z=np.array(mols[name]['numbers']); p0=np.array(mols[name]['positions'])
    p=bench._perturb(p0,[seed])[0] if seed>=0 else p0
    m=O.make_mol(z)
    S,_=O.overlap(m,p); cn,_=O.cn_d3(m,p); H0=O.h0(m,p,S,cn)
    gam=O.gamma_shell(m,p); g3=O.gam3(m)
    n0=O._shell_param(m,"refocc")[m.ao_sh]/(2*m.sh_l[m.ao_sh]+1)
    nel=n0.sum(); no=int(round(nel/2)); n=m.nao; nv=n-no
    L=cholesky(S,lower=True); C=np.linalg.inv(L).T
    q0=O.guess_orbital_charges(m,O.eeq_charges(m,p,0.0))
    v,_,_=O._potential(m,q0,gam,g3)
    mixer=O.Anderson(n)
    X=None; Z=None; elig=False
    tot=dict(sw=0,ric=0,newt=0,fastit=0)
    maxerr=0
    def full(A,C,tol):
        sw=0
        while offmax(A)>tol:
            jacobi_sweep(A,C); sw+=1
        return sw
    for it in range(80):
        F=H0-0.5*S*(v[:,None]+v[None,:])
        A=C.T@F@C; A=0.5*(A+A.T)
        w,U=eigh(F,S); Pex=2*U[:,:no]@U[:,:no].T
        sw=0; ric=0; newt=0; mode='full'
        done=False
        while not done:
            if fast and elig:
                d=np.diag(A); off=np.abs(A-np.diag(d))
                hi=(d[:no]+off[:no,:no].sum(1)).max(); lo=(d[no:]-off[no:,no:].sum(1)).min()
                gapc=lo-hi
                if gapc>=kcert*KT:
                    # riccati
                    Aoo=A[:no,:no]; Avv=A[no:,no:]; Avo=A[no:,:no]
                    den=np.diag(Avv)[:,None]-np.diag(Aoo)[None,:]
                    ok=False; rprev=None
                    Xw=X.copy()
                    for k in range(maxric):
                        T=A[:,no:]@Xw
                        Lam=Aoo+T[:no]
                        R=Avo+T[no:]-Xw@Lam
                        r=np.abs(R).max()
                        if r<=rtol: ok=True; break
                        Xw=Xw-R/den; ric+=1
                    if ok and np.abs(Xw).max()<1.0:
                        X=Xw
                        Y=C[:,:no]+C[:,no:]@X
                        G=np.eye(no)+X.T@X
                        for kk in range(10):
                            E=np.eye(no)-G@Z
                            if np.abs(E).max()<=1e-12: break
                            Z=Z+Z@E; newt+=1
                        P=2*Y@Z@Y.T
                        mode='fast'; done=True; tot['fastit']+=1
                        continue
            # one sweep (or converged full)
            d_=np.diag(A); o_=np.argsort(d_,kind='stable'); A2=A[np.ix_(o_,o_)]; off_=np.abs(A2-np.diag(np.diag(A2)))
            early = EARLY and (not elig) and sw>0 and ((np.diag(A2)[no:]-off_[no:,no:].sum(1)).min()-(np.diag(A2)[:no]+off_[:no,:no].sum(1)).max())>=kcert*KT
            if early:
                C[:]=C[:,o_]; A[:]=A2; elig=True; X=np.zeros((nv,no)); Z=np.eye(no); continue
            if offmax(A)<=1e-9:
                # converged: classify
                order=np.argsort(np.diag(A),kind='stable')
                C[:]=C[:,order]; A[:]=A[np.ix_(order,order)]
                e=np.diag(A)
                elig = (e[no]-e[no-1])>=kcert*KT
                X=np.zeros((nv,no)); Z=np.eye(no)
                P=2*C[:,:no]@C[:,:no].T
                done=True
            else:
                jacobi_sweep(A,C); sw+=1
                if X is not None: X[:]=0; Z=np.eye(no)
        errq=np.abs(np.einsum('ik,ki->i',P-Pex,S)).max(); maxerr=max(maxerr,errq)
        tot['sw']+=sw; tot['ric']+=ric; tot['newt']+=newt
        if verbose: print(f'  it {it:2d} {mode} sweeps {sw} ricc {ric:2d} newt {newt} errq {errq:.1e}')
        q=n0-np.einsum('ik,ki->i',P,S)
        vnew,_,_=O._potential(m,q,gam,g3)
        v=mixer.iter(vnew,v)
        if it>0 and mixer.converged(1e-4,1e-5): break
    # final solve
    F=H0-0.5*S*(vnew[:,None]+vnew[None,:])
    A=C.T@F@C; A=0.5*(A+A.T)
    fsw=full(A,C,1e-13)
    return dict(name=name,n=n,no=no,iters=it+1,final_sw=fsw,maxerrq=maxerr,**tot)

EARLY=True
if __name__=='__main__':
    names=sys.argv[1].split(',')
    for nm in names:
        for seed in (3,):
            a=run(nm,seed,False); b=run(nm,seed,True,verbose='-v' in sys.argv)
            ca=a['sw']+a['final_sw']; cb=b['sw']+b['final_sw']+0.08*b['ric']+0.04*b['newt']+0.12*b['fastit']
            print(a); print(b); print(f'{nm}: cost full {ca:.1f} fast {cb:.1f} ratio {ca/cb:.2f}')

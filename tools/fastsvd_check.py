"""CPU check of csrc/sfm_fastsvd.cuh (host-callable): the short-latency DLT null vector and the closed-form essential
decomposition against numpy's SVD route on synthetic scenes (noise 0 / 0.5 / 5 px, with and without outliers, pixel and
normalised camera matrices).  usage: python tools/fastsvd_check.py   (builds tools/bin/libfastsvd_check.so with nvcc)"""
import ctypes, numpy as np, sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "tools", "bin"), exist_ok=True)
subprocess.check_call(["nvcc", "-O2", "-shared", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets", "-o",
                       os.path.join(ROOT, "tools", "bin", "libfastsvd_check.so"), os.path.join(ROOT, "tools", "fastsvd_check.cu")])
from structure_from_motion_b200.scenes import make_scene
from oracle import restatement as o
L=ctypes.CDLL(os.path.join(ROOT, 'tools', 'bin', 'libfastsvd_check.so'))
P=lambda a: a.ctypes.data_as(ctypes.c_void_p)
def null4(A):
    A=np.ascontiguousarray(A,dtype=np.float64); x=np.empty(4)
    ok=L.null4(P(A),P(x)); return bool(ok),x
def frames3(A):
    A=np.ascontiguousarray(A,dtype=np.float64); U=np.empty((3,3));V=np.empty((3,3));s=np.empty(3)
    ok=L.frames3(P(A),P(U),P(V),P(s)); return bool(ok),U,V,s
def build_A(xa,ya,xb,yb,P1,P2):
    return np.array([ya*P1[2]-P1[1], P1[0]-xa*P1[2], yb*P2[2]-P2[1], P2[0]-xb*P2[2]])
tot=0; fb=0; worst=0
for noise in (0.0,0.5,5.0):
  for frac in (0.0,0.4):
    K,x1,x2,Rt,tt,_=make_scene(2000,frac,seed=3,noise_px=noise)
    T=np.eye(4);T[:3,:3]=Rt;T[:3,3]=tt
    Kx=np.hstack([K,np.zeros((3,1))]);P1=Kx@np.eye(4);P2=Kx@T
    nxa,nya=o.k_normalise(x1[:,0],x1[:,1],K); nxb,nyb=o.k_normalise(x2[:,0],x2[:,1],K)
    for mode in (0,1):
      nf=0;w=0
      for i in range(0,2000,3):
        A = build_A(x1[i,0],x1[i,1],x2[i,0],x2[i,1],P1,P2) if mode==0 else build_A(nxa[i],nya[i],nxb[i],nyb[i],np.eye(4)[:3],T[:3])
        ok,x=null4(A)
        tot+=1
        if not ok: nf+=1; continue
        vh=np.linalg.svd(A)[2][-1]; Xr=vh[:3]/vh[3]; X=x[:3]/x[3]
        w=max(w,np.linalg.norm(X-Xr)/np.linalg.norm(Xr))
      fb+=nf; worst=max(worst,w)
      print(f"noise {noise} frac {frac} mode {mode}: fallback {nf}/{len(range(0,2000,3))} worst rel {w:.2e}")
print("total",tot,"fallbacks",fb,"worst",worst)
# essential decomposition frames
rng=np.random.default_rng(1)
wr=0;wt=0;nfb=0
for trial in range(3000):
    if trial%3==0:
        # exact essential from random R,t, scaled
        q=rng.normal(size=4);q/=np.linalg.norm(q)
        a,b,c,d=q
        R=np.array([[a*a+b*b-c*c-d*d,2*(b*c-a*d),2*(b*d+a*c)],[2*(b*c+a*d),a*a-b*b+c*c-d*d,2*(c*d-a*b)],[2*(b*d-a*c),2*(c*d+a*b),a*a-b*b-c*c+d*d]])
        t=rng.normal(size=3)
        tx=np.array([[0,-t[2],t[1]],[t[2],0,-t[0]],[-t[1],t[0],0]])
        E=tx@R*rng.uniform(0.1,10)
    else:
        # rank-2 projected 8-point fit of a noisy scene sample
        K,x1,x2,*_=make_scene(200,0.3,seed=trial,noise_px=0.5)
        nxa,nya=o.k_normalise(x1[:,0],x1[:,1],K); nxb,nyb=o.k_normalise(x2[:,0],x2[:,1],K)
        s=rng.choice(200,8,replace=False)
        try: E=o.eight_point(np.stack([nxa[s],nya[s]],1),np.stack([nxb[s],nyb[s]],1))
        except Exception: continue
    ok,U,V,sv=frames3(E)
    if not ok: nfb+=1; continue
    assert abs(np.linalg.det(U)-1)<1e-12 and abs(np.linalg.det(V)-1)<1e-12, (np.linalg.det(U),np.linalg.det(V))
    assert np.abs(U.T@U-np.eye(3)).max()<1e-13 and np.abs(V.T@V-np.eye(3)).max()<1e-13
    W=np.array([[0,-1,0],[1,0,0],[0,0,1.]])
    R1=U@W.T@V.T; R2=U@W@V.T; t1=U[:,2]
    Ro1,Ro2,to=o.recover_all_r_t(E)
    dR=min(max(np.abs(R1-Ro1).max(),np.abs(R2-Ro2).max()),max(np.abs(R1-Ro2).max(),np.abs(R2-Ro1).max()))
    dt=min(np.abs(t1-to).max(),np.abs(t1+to).max())
    wr=max(wr,dR);wt=max(wt,dt)
print("frames: worst dR",wr,"worst dt",wt,"fallbacks",nfb)

import ctypes as C, numpy as np, subprocess, os, sys
ROOT='/root/repo'; sys.path.insert(0,ROOT)
so='/tmp/libvc_harness6.so'
subprocess.run(['/usr/bin/g++','-O2','-std=c++17','-ffp-contract=off','-shared','-fPIC','-o',so,os.path.join(ROOT,'tests','voronoi_harness.cpp')],check=True)
L=C.CDLL(so)
from voronoirt_b200 import synth
b=np.array([0,1,0,1,0,1.0]); bounds=dict(z_min=0,z_max=1,x_min=0,x_max=1,y_min=0,y_max=1)
def run(pos3,label):
    n=pos3.shape[1]
    gold=np.asarray(synth.voronoi_neighbours(np.asfortranarray(pos3),bounds=bounds))
    pos=np.ascontiguousarray(pos3.T)
    g=max(1,int(round((n/4)**(1/3))))
    nbr=np.zeros((n,64),dtype=np.int64); st=np.zeros(n,dtype=np.int32)
    bad=L.vc_harness(C.c_int64(n),pos.ctypes.data_as(C.c_void_p),b.ctypes.data_as(C.c_void_p),g,g,g,nbr.ctypes.data_as(C.c_void_p),C.c_int64(64),st.ctypes.data_as(C.c_void_p))
    mism=[(i,sorted(nbr[i,1:1+nbr[i,0]].tolist()),sorted(gold[i,1:1+gold[i,0]].tolist())) for i in range(n) if sorted(nbr[i,1:1+nbr[i,0]].tolist())!=sorted(gold[i,1:1+gold[i,0]].tolist())]
    print(label,'n',n,'bad',bad,'mismatch',len(mism),'faces mine/voro',nbr[:,0].mean(),gold[:,0].mean(), mism[:1])
m=6
ax=(np.arange(m)+0.5)/m
Z,X,Y=np.meshgrid(ax,ax,ax,indexing='ij')
lat=np.stack([Z.ravel(),X.ravel(),Y.ravel()])
run(lat,'cubic lattice')
rng=np.random.default_rng(0)
run(lat+rng.normal(scale=1e-9,size=lat.shape),'lattice + 1e-9 jitter')
run(lat+rng.normal(scale=1e-3,size=lat.shape),'lattice + 1e-3 jitter')
# clustered
c=rng.random((3,20)); pts=(c[:,rng.integers(0,20,5000)]+rng.normal(scale=0.01,size=(3,5000)))%1.0
pts[0]=np.clip(pts[0],1e-6,1-1e-6)
run(pts,'clustered')

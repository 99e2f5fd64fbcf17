"""Host code (analysis.cpp, rowperm.cpp) and the CPU plan interpreter under AddressSanitizer + UBSan: see scripts/asan_host_check.sh."""
import ctypes
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import synth_case
from test_oracle_and_plan import run_sim, _ip
from test_rowperm import sim_rowperm, pow2
sim = ctypes.CDLL(os.environ['NKP_ASAN_LIB'])
P=ctypes.POINTER
for shp,nr in [((12,10,5),1),((20,24,10),3),((30,34,20),8),((30,34,20),1)]:
    c=synth_case(*shp); n=c["n"]
    A=sp.csr_matrix((c["nzval"],c["colind"],c["rowptr"]),shape=(n,n))
    xs=np.random.default_rng(0).standard_normal((n,2)); B=A@xs
    rc,X,st=sim_rowperm(sim,n,c["rowptr"],c["colind"],c["nzval"],(c["i"],c["j"],c["k"]),B,nranks=nr)
    print(shp,nr,rc,np.abs(X-xs).max())
    # rowperm path
    rowmap=np.zeros(n,np.int32); R=np.zeros(n); C=np.zeros(n)
    rc=sim.nkp_rowperm_largediag(n,_ip(c["rowptr"]),_ip(c["colind"]),c["nzval"].ctypes.data_as(P(ctypes.c_double)),_ip(rowmap),R.ctypes.data_as(P(ctypes.c_double)),C.ctypes.data_as(P(ctypes.c_double)))
    rc2,X,st=sim_rowperm(sim,n,c["rowptr"],c["colind"],c["nzval"],(c["i"],c["j"],c["k"]),B,rowmap,pow2(R),pow2(C),nranks=nr)
    print("  rowperm",rc,rc2,np.abs(X-xs).max())
    cnt=np.zeros(n,np.int64); perm=np.zeros(n,np.int32); out=np.zeros(4*n,np.int32)
    nf=sim.nkp_sim_fronts(n,_ip(c["rowptr"]),_ip(c["colind"]),_ip(c["i"]),_ip(c["j"]),_ip(c["k"]),64,96,_ip(out),n,_ip(perm))
    sim.nkp_true_colcounts(n,_ip(c["rowptr"]),_ip(c["colind"]),_ip(perm),cnt.ctypes.data_as(P(ctypes.c_longlong)))
    rc=sim.nkp_sim_check_plan(n,_ip(c["rowptr"]),_ip(c["colind"]),_ip(c["i"]),_ip(c["j"]),_ip(c["k"]),64,96,nr)
    print("  fronts",nf,"check_plan",rc)
# degenerate
for A in [sp.diags(np.ones(10)*2.0), sp.diags([-np.ones(199),4*np.ones(200),-np.ones(199)],[-1,0,1])]:
    A=sp.csr_matrix(A); m=A.shape[0]; xs=np.ones((m,1))
    X,st,perm=run_sim(sim,m,A.indptr.astype(np.int32),A.indices.astype(np.int32),A.data.astype(float),None,A@xs)
    print("degenerate",m,np.abs(X-xs).max())

#!/usr/bin/env python
"""Lock-step Python model of WalkerSub (NG nonzeros per warp-wide gather, per-group partial sums) (ge-spmm_b200/csrc/gespmm_spmm.cu) inside kernel A's 32-row batching.
32 "lanes" are stepped together; shuffles, ballots and warp reductions are plain Python over the lane lists, and
cp.async is modelled with commit groups whose bytes only become readable at the matching wait (a read of in-flight
data raises).  A development aid for a GPU-less container: index mappings, row-end handling and ring-slot reuse are
checked against a sequential fp32 loop before the CUDA version goes to the GPU.  Integer-valued operands: the sum does not depend on the association, so equality is exact.
    python scripts/models/walker_sub_model.py
Not part of the product and not a test of it; the GPU parity tests are tests/test_spmm_gpu.py.
"""
import numpy as np, sys

def low_bits(n): return 0xffffffff if n >= 32 else (1 << n) - 1
def ffs(x):
    x=int(x); return (x & -x).bit_length()

class Sub:
    def __init__(s, NG, K, colind, val, B, C, init=0.0, QSMAX=8):
        s.NG=NG; s.LPR=32//NG; s.Q=32//NG; s.QS=min(s.Q,QSMAX); s.SN=s.QS*NG; s.SPC=s.Q//s.QS; s.UB=4
        s.stage=s.QS*512
        s.K=K; s.colind=colind; s.val=val; s.B=B; s.C=C
        s.ring={}   # (slot, quad, lane) -> float4 or None(poison)
        s.pending=[]  # list of groups, each a list of (key, value)
        s.g=[l//s.LPR for l in range(32)]; s.col0=[(l%s.LPR)*4 for l in range(32)]
        s.active=[c<K for c in s.col0]
    def commit(s, grp): s.pending.append(grp)
    def wait(s, n):
        while len(s.pending)>n:
            for k,v in s.pending.pop(0): s.ring[k]=v
    def issue(s, cols, pos0, n, slot):
        full = pos0+s.SN<=n
        grp=[]
        for i in range(s.QS):
            for l in range(32):
                idx=pos0+i*s.NG+s.g[l]
                assert 0<=idx<32
                c=cols[idx]
                if s.active[l] and (full or idx<n):
                    grp.append(((slot,i,l), s.B[c, s.col0[l]:s.col0[l]+4].copy()))
                    s.ring[(slot,i,l)]=None  # in flight: poison
        s.commit(grp)
    def combine(s, t):
        off=16
        t=[x.copy() for x in t]
        while off>=s.LPR:
            t=[ (t[l]+t[l^off]).astype(np.float32) for l in range(32)]
            off>>=1
        return t
    def flush(s, acc, st):
        t=s.combine(acc)
        row=st['rb']+ffs(st['rows_left'])-1
        for l in range(32):
            if s.g[l]==0 and s.active[l]:
                s.C[row, s.col0[l]:s.col0[l]+4]=t[l]
        st['written'][row]+=1
        st['rows_left']&=st['rows_left']-1
        for l in range(32): acc[l]=np.zeros(4,np.float32)
    def step(s, acc, l, a, b):
        if s.val is not None: acc[l]=(acc[l]+np.float32(a)*b).astype(np.float32)
        else: acc[l]=(acc[l]+b).astype(np.float32)
    def consume(s, vals, pos0, n, endmask, acc, st, slot):
        full=pos0+s.SN<=n
        for i0 in range(0,s.QS,s.UB):
            b=[[None]*32 for _ in range(s.UB)]; a=[[1.0]*32 for _ in range(s.UB)]
            for i in range(s.UB):
                for l in range(32):
                    idx=pos0+(i0+i)*s.NG+s.g[l]
                    assert 0<=idx<32
                    a[i][l]=vals[idx]
                    b[i][l]=np.zeros(4,np.float32)
                    if s.active[l] and (full or idx<n):
                        v=s.ring[(slot,i0+i,l)]
                        assert v is not None, "read of in-flight data"
                        b[i][l]=v
            ends=(endmask>>(pos0+i0*s.NG))&low_bits(s.UB*s.NG)
            if full and ends==0:
                for i in range(s.UB):
                    for l in range(32): s.step(acc,l,a[i][l],b[i][l])
            else:
                for i in range(s.UB):
                    live=[full or (pos0+(i0+i)*s.NG+s.g[l]<n) for l in range(32)]
                    e4=(ends>>(i*s.NG))&((1<<s.NG)-1); lo=0
                    while e4:
                        hi=ffs(e4)-1
                        for l in range(32):
                            if live[l] and lo<=s.g[l]<=hi: s.step(acc,l,a[i][l],b[i][l])
                        s.flush(acc,st)
                        lo=hi+1; e4&=e4-1
                    for l in range(32):
                        if live[l] and s.g[l]>=lo: s.step(acc,l,a[i][l],b[i][l])
    def stream(s, S, E, acc, my_end, rows, rb, written):
        def ld(arr,p,d): return [ (arr[p+l] if p+l<E else d) for l in range(32)]
        ccol=ld(s.colind,S,0); cval=ld(s.val,S,1.0) if s.val is not None else [1.0]*32
        ncol=ld(s.colind,S+32,0); fcol=[0]*32; nval=[1.0]*32
        st={'rows_left':rows,'rb':rb,'written':written}
        slot=0
        s.issue(ccol,0,min(32,E-S),0)
        p0=S
        while p0<E:
            f2=ld(s.colind,p0+64,None); fcol=[f2[l] if f2[l] is not None else fcol[l] for l in range(32)]
            if s.val is not None:
                n2=ld(s.val,p0+32,None); nval=[n2[l] if n2[l] is not None else nval[l] for l in range(32)]
            endmask=0
            for l in range(32):
                if (rows>>l)&1:
                    rel=my_end[l]-1-p0
                    if 0<=rel<32: endmask|=1<<rel
            n=min(32,E-p0); n_next=E-p0-32
            for j in range(s.SPC):
                if j*s.SN>=n: break
                nxt=j+1>=s.SPC
                s.issue(ncol if nxt else ccol, 0 if nxt else (j+1)*s.SN, n_next if nxt else n, slot^s.stage)
                s.wait(1)
                s.consume(cval,j*s.SN,n,endmask,acc,st,slot)
                slot^=s.stage
            ccol,ncol,cval=ncol,fcol,nval
            p0+=32
        s.wait(0)
        assert st['rows_left']==0, "rows left unflushed"

def kernelA(NG,K,rowptr,colind,val,B,long_row=4096):
    M=len(rowptr)-1
    C=np.full((M,K if K%4==0 else K),np.nan,np.float32)
    Cpad=np.full((M,64),np.nan,np.float32)
    w=Sub(NG,K,colind,val,B,Cpad)
    written=np.zeros(M,int)
    rb=0
    while rb<M:
        nrows=min(32,M-rb)
        my_start=[int(rowptr[rb+l]) if l<nrows else 0 for l in range(32)]
        my_end=[int(rowptr[rb+l+1]) if l<nrows else 0 for l in range(32)]
        ln=[my_end[l]-my_start[l] for l in range(32)]
        long_mask=sum(1<<l for l in range(32) if ln[l]>long_row)
        nonempty=sum(1<<l for l in range(32) if ln[l]>0)&~long_mask
        em=~(nonempty|long_mask)&low_bits(nrows)
        while em:
            r=rb+ffs(em)-1
            for l in range(32):
                if w.g[l]==0 and w.active[l]: Cpad[r,w.col0[l]:w.col0[l]+4]=0
            written[r]+=1; em&=em-1
        run=0
        while True:
            stop=(ffs(long_mask)-1) if long_mask else nrows
            rows=nonempty&low_bits(stop)&~low_bits(run)
            if rows:
                S=my_start[ffs(rows)-1]; E=my_end[rows.bit_length()-1]
                acc=[np.zeros(4,np.float32) for _ in range(32)]
                w.stream(S,E,acc,my_end,rows,rb,written)
            if stop>=nrows: break
            long_mask&=long_mask-1; run=stop+1
        rb+=32
    return Cpad[:,:K],written

rng=np.random.default_rng(0)
for trial in range(40):
    NG=[2,4,8][trial%3]
    K={2:[36,48,64],4:[20,32,24],8:[4,8,12,16]}[NG][trial%3]
    M=int(rng.integers(1,150)); N=50
    deg=rng.integers(0,12,M)
    if trial%4==0: deg[rng.integers(0,M)]=rng.integers(60,200)
    if trial%5==0: deg[:]=rng.integers(0,3,M)
    rowptr=np.concatenate([[0],np.cumsum(deg)]).astype(np.int64)
    nnz=int(rowptr[-1])
    colind=rng.integers(0,N,nnz)
    B=rng.integers(-8,9,(N,64)).astype(np.float32); B[:,K:]=np.nan
    val=None if trial%2 else rng.integers(-2,3,nnz).astype(np.float32)
    C,written=kernelA(NG,K,rowptr,colind,val,B)
    # exact integer arithmetic -> order independent
    want=np.zeros((M,K),np.float32)
    for r in range(M):
        for p in range(rowptr[r],rowptr[r+1]):
            want[r]+= (1.0 if val is None else val[p])*B[colind[p],:K]
    assert (written==1).all(), (trial, written)
    assert np.array_equal(C,want),(trial,NG,K)
    print("ok",trial,NG,K,M,nnz)

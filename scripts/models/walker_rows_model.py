#!/usr/bin/env python
"""Lock-step Python model of WalkerRows (lane groups own disjoint rows, sequential order) (ge-spmm_b200/csrc/gespmm_spmm.cu) inside kernel A's 32-row batching.
32 "lanes" are stepped together; shuffles, ballots and warp reductions are plain Python over the lane lists, and
cp.async is modelled with commit groups whose bytes only become readable at the matching wait (a read of in-flight
data raises).  A development aid for a GPU-less container: index mappings, row-end handling and ring-slot reuse are
checked against a sequential fp32 loop before the CUDA version goes to the GPU.  Real-valued operands: the result must be bit-identical to the sequential loop.
    python scripts/models/walker_rows_model.py
Not part of the product and not a test of it; the GPU parity tests are tests/test_spmm_gpu.py.
"""
import numpy as np

def low_bits(n): return 0xffffffff if n >= 32 else (1 << n) - 1
def ffs(x):
    x=int(x); return (x & -x).bit_length()
def clz(x): return 32-int(x).bit_length()
f32=np.float32

class Rows:
    def __init__(s, NG, K, colind, val, B, C):
        s.NG=NG; s.LPR=32//NG; s.QS=min(s.LPR,8); s.SPC=s.LPR//s.QS; s.UB=4; s.stage=s.QS*512
        s.K=K; s.colind=colind; s.val=val; s.B=B; s.C=C
        s.ring={}; s.pending=[]
        s.g=[l//s.LPR for l in range(32)]; s.sl=[l%s.LPR for l in range(32)]; s.col0=[(l%s.LPR)*4 for l in range(32)]
        s.active=[c<K for c in s.col0]
        s.ones={4:0x11111111,8:0x01010101,16:0x00010001}[s.LPR]
    def wait(s,n):
        while len(s.pending)>n:
            for k,v in s.pending.pop(0): s.ring[k]=v
    def issue(s, cols, k0, n, slot):
        grp=[]
        for i in range(s.QS):
            for l in range(32):
                k=k0+i; assert 0<=k<s.LPR
                c=cols[s.g[l]*s.LPR+k]
                if s.active[l] and k<n[l]:
                    grp.append(((slot,i,l), s.B[c, s.col0[l]:s.col0[l]+4].copy())); s.ring[(slot,i,l)]=None
        s.pending.append(grp)
    def step(s, acc, l, a, b):
        if s.val is not None:
            # fmaf: single rounding -> emulate in float64 then round (products of f32 are exact in f64; sum rounding once)
            acc[l]=(acc[l].astype(np.float64)+np.float64(a)*b.astype(np.float64)).astype(f32)
        else: acc[l]=(acc[l]+b).astype(f32)
    def consume(s, vals, k0, n, nmin, endmask, acc, left, rb, slot, written):
        for i0 in range(0,s.QS,s.UB):
            a=[[1.0]*32 for _ in range(s.UB)]; b=[[None]*32 for _ in range(s.UB)]
            for i in range(s.UB):
                for l in range(32):
                    k=k0+i0+i
                    a[i][l]=vals[s.g[l]*s.LPR+k]
                    b[i][l]=s.ring.get((slot,i0+i,l),'stale')
            ends=endmask & ((s.ones*((1<<s.UB)-1))<<(k0+i0)) & 0xffffffff
            if ends==0 and nmin>=k0+i0+s.UB:
                for i in range(s.UB):
                    for l in range(32):
                        if s.active[l]:
                            assert b[i][l] is not None and not isinstance(b[i][l],str)
                            s.step(acc,l,a[i][l],b[i][l])
            else:
                for i in range(s.UB):
                    k=k0+i0+i
                    for l in range(32):
                        if k<n[l] and s.active[l]:
                            assert b[i][l] is not None and not isinstance(b[i][l],str), "read of in-flight/stale data"
                            s.step(acc,l,a[i][l],b[i][l])
                    for l in range(32):
                        if (endmask>>(s.g[l]*s.LPR+k))&1:
                            row=rb+ffs(left[l])-1
                            if s.active[l]:
                                s.C[row, s.col0[l]:s.col0[l]+4]=acc[l]
                            if s.sl[l]==0: written[row]+=1
                            left[l]&=left[l]-1
                            acc[l]=np.zeros(4,f32)
    def stream(s, S, E, acc, my_end, rows, rb, written):
        NG,LPR=s.NG,s.LPR
        my_row=[(rows>>l)&1 for l in range(32)]
        my_start=[0]*32
        for l in range(32):
            below=rows&low_bits(l)
            prev_end=my_end[31-clz(below)] if below else my_end[0]
            my_start[l]=prev_end if below else S
        grp=[0]*32
        for l in range(32):
            if my_row[l]: grp[l]=min(NG-1, ((my_start[l]+my_end[l]-2*S)*NG)//(2*(E-S)))
        mine=[0]*32
        for q in range(NG):
            m=sum(1<<l for l in range(32) if my_row[l] and grp[l]==q)
            for l in range(32):
                if s.g[l]==q: mine[l]=m
        gs=[0]*32; ge=[0]*32
        for l in range(32):
            if mine[l]:
                gs[l]=my_start[ffs(mine[l])-1]; ge[l]=my_end[31-clz(mine[l])]
        row_gs=[gs[grp[l]*LPR] for l in range(32)]
        maxlen=max(ge[l]-gs[l] for l in range(32))
        left=list(mine); p=list(gs)
        def ld(arr,off,d,prev=None):
            out=[]
            for l in range(32):
                idx=p[l]+off+s.sl[l]
                out.append(arr[idx] if idx<ge[l] else (d if prev is None else prev[l]))
            return out
        ccol=ld(s.colind,0,0); cval=ld(s.val,0,1.0) if s.val is not None else [1.0]*32
        ncol=ld(s.colind,LPR,0); fcol=[0]*32; nval=[1.0]*32
        slot=0
        s.issue(ccol,0,[ge[l]-p[l] for l in range(32)],0)
        c0=0
        while c0<maxlen:
            fcol=ld(s.colind,2*LPR,None,fcol)
            if s.val is not None: nval=ld(s.val,LPR,None,nval)
            endmask=0
            for l in range(32):
                rel=my_end[l]-1-(row_gs[l]+c0)
                if my_row[l] and 0<=rel<LPR: endmask|=1<<(grp[l]*LPR+rel)
            n=[ge[l]-p[l] for l in range(32)]; nmin=min(n)
            for j in range(s.SPC):
                if j*s.QS>=maxlen-c0: break
                nxt=j+1>=s.SPC
                s.issue(ncol if nxt else ccol, 0 if nxt else (j+1)*s.QS, [x-LPR for x in n] if nxt else n, slot^s.stage)
                s.wait(1)
                s.consume(cval,j*s.QS,n,nmin,endmask,acc,left,rb,slot,written)
                slot^=s.stage
            ccol,ncol,cval=ncol,fcol,nval
            c0+=LPR; p=[x+LPR for x in p]
        s.wait(0)
        assert all(x==0 for x in left), "rows left unflushed"

def kernelA(NG,K,rowptr,colind,val,B,long_row=4096):
    M=len(rowptr)-1
    Cpad=np.full((M,64),np.nan,f32)
    w=Rows(NG,K,colind,val,B,Cpad)
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
                acc=[np.zeros(4,f32) for _ in range(32)]
                w.stream(S,E,acc,my_end,rows,rb,written)
            if stop>=nrows: break
            long_mask&=long_mask-1; run=stop+1
        rb+=32
    return Cpad[:,:K],written,long_mask

rng=np.random.default_rng(1)
for trial in range(48):
    NG=[2,4,8][trial%3]
    K={2:[36,48,64],4:[20,32,24],8:[4,8,12,16]}[NG][trial%3]
    M=int(rng.integers(1,130)); N=40
    deg=rng.integers(0,12,M)
    if trial%4==0: deg[rng.integers(0,M)]=rng.integers(60,300)
    if trial%5==0: deg[:]=rng.integers(0,3,M)
    if trial%7==0: deg[rng.integers(0,M)]=150   # "long" row with long_row=100 below
    rowptr=np.concatenate([[0],np.cumsum(deg)]).astype(np.int64)
    nnz=int(rowptr[-1])
    colind=rng.integers(0,N,nnz)
    B=rng.standard_normal((N,64)).astype(f32); B[:,K:]=np.nan
    val=None if trial%2 else rng.standard_normal(nnz).astype(f32)
    LR=100 if trial%7==0 else 4096
    C,written,_=kernelA(NG,K,rowptr,colind,val,B,long_row=LR)
    want=np.zeros((M,K),f32)
    for r in range(M):
        acc=np.zeros(K,f32)
        for pp in range(rowptr[r],rowptr[r+1]):
            if val is None: acc=(acc+B[colind[pp],:K]).astype(f32)
            else: acc=(acc.astype(np.float64)+np.float64(val[pp])*B[colind[pp],:K].astype(np.float64)).astype(f32)
        want[r]=acc
    short=np.diff(rowptr)<=LR
    assert (written[short]==1).all(), (trial, written)
    assert (written[~short]==0).all()
    assert np.array_equal(C[short],want[short]),(trial,NG,K)
    print("ok",trial,NG,K,M,nnz,int((~short).sum()))

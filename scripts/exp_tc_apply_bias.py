"""Experiment (CPU): accumulator-truncation bias of the tensor-core block-rotation apply, emulated
(TF32 operand rounding, exact 8-term products, fp32 accumulator rounded toward zero once per MMA),
for Q T and for the (Q - I) T + T form of the -DTA_DELTA=1 build variant.  python scripts/exp_tc_apply_bias.py"""
import numpy as np
rng=np.random.default_rng(0)
def rz32(x):
    # round float64 -> float32 toward zero
    y=x.astype(np.float32)
    bad=np.abs(y.astype(np.float64))>np.abs(x)
    y[bad]=np.nextafter(y[bad], np.float32(0))
    return y
def tf32(x):
    # round-to-nearest-away (cvt.rna) to 10-bit mantissa
    b=x.astype(np.float32).view(np.uint32).astype(np.uint64)
    b=(b+0x1000)&0xFFFFE000
    return b.astype(np.uint32).view(np.float32)
def split(x):
    hi=tf32(x); lo=(x.astype(np.float32)-hi).astype(np.float32); return hi, tf32_trunc(lo)
def tf32_trunc(x):
    b=x.view(np.uint32)&np.uint32(0xFFFFE000); return b.view(np.float32)
def tc_apply(Q,T,delta):
    P=32; m=T.shape[1]
    Qe=Q-np.eye(P) if delta else Q
    # real embedding: A[m][2k+t], B[2i+u][2k+t]
    A=np.zeros((m,64),np.float32); A[:,0::2]=T.real.T; A[:,1::2]=T.imag.T
    B=np.zeros((64,64),np.float32)
    B[0::2,0::2]=Qe.real; B[0::2,1::2]=-Qe.imag; B[1::2,0::2]=Qe.imag; B[1::2,1::2]=Qe.real
    Ah,Al=split(A); Bh,Bl=split(B)
    acc=np.zeros((m,64),np.float32)
    for ks in range(8):
        sl=slice(8*ks,8*ks+8)
        for (a,b) in ((Ah,Bl),(Al,Bh),(Ah,Bh)):
            d=a[:,sl].astype(np.float64)@b[:,sl].astype(np.float64).T
            acc=rz32(acc.astype(np.float64)+d)
    out=(acc[:,0::2]+1j*acc[:,1::2]).T      # [i][m]
    if delta: out=(T.astype(np.complex64)+out.astype(np.complex64))
    return out.astype(np.complex64)
def rand_unitary_near_I(eps):
    H=rng.standard_normal((32,32))+1j*rng.standard_normal((32,32)); H=(H+H.conj().T)/2
    w,v=np.linalg.eigh(H); return (v*np.exp(1j*eps*w))@v.conj().T
for eps in (0.3,0.03,0.003):
    for delta in (0,1):
        shr=[];err=[]
        for rep in range(6):
            Q=rand_unitary_near_I(eps).astype(np.complex64)
            T=(rng.standard_normal((32,128))+1j*rng.standard_normal((32,128))).astype(np.complex64)
            exact=Q.astype(complex)@T.astype(complex)
            out=tc_apply(Q.astype(complex),T,delta)
            shr.append((np.linalg.norm(out)**2-np.linalg.norm(exact)**2)/np.linalg.norm(exact)**2/2)
            err.append(np.abs(out-exact).max()/np.abs(exact).max())
        print(f"eps={eps} delta={delta}: mean relative norm change {np.mean(shr):+.2e}, max elementwise error {np.max(err):.2e}")

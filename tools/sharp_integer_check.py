"""Sharpness: when does the float32 chain of tfa.image.sharpness truncate to floor(S / 13)?  (profiles/r02_ab_notes.md 6)

S = sum of the 8 neighbours + 5 x centre.  Finding: always when S % 13 != 0; 5.3 % of the S % 13 == 0 cases truncate one lower."""
import numpy as np
f32=np.float32
k1=f32(1.0)/f32(13.0); k5=f32(5.0)/f32(13.0)
rng=np.random.default_rng(0)
def chain(n):  # n: [N,9] uint8 in row-major, index 4 = centre
    acc=np.zeros(n.shape[0],dtype=f32)
    for i in range(9):
        k = k5 if i==4 else k1
        acc = acc + n[:,i].astype(f32)*k
    return acc
tot=0; bad_nonmult=0; mult=0; mult_low=0
for it in range(60):
    kind = it%4
    if kind==0: n=rng.integers(0,256,size=(2_000_000,9)).astype(np.uint8)
    elif kind==1:
        base=rng.integers(0,256,size=(2_000_000,1)); n=np.clip(base+rng.integers(-3,4,size=(2_000_000,9)),0,255).astype(np.uint8)
    elif kind==2:
        n=rng.choice(np.array([0,255,254,1,128,13,26],dtype=np.uint8),size=(2_000_000,9))
    else:
        n=rng.integers(0,256,size=(2_000_000,9)).astype(np.uint8); n[:, rng.integers(0,9)] = 255
    acc=chain(n)
    S=n.astype(np.int64).sum(1)+4*n[:,4].astype(np.int64)
    q=S//13; r=S%13
    t=np.trunc(acc).astype(np.int64)
    nm = r!=0
    bad_nonmult += int((t[nm]!=q[nm]).sum())
    m = ~nm
    mult += int(m.sum()); mult_low += int((t[m]!=q[m]).sum())
    # margin: how close does acc get to an integer boundary for nonmultiples
    tot+=n.shape[0]
print("total",tot,"nonmultiple mismatches",bad_nonmult,"multiples",mult,"of which trunc!=S/13:",mult_low)
# magic multiply check
S=np.arange(0,3316); P=S*5042
assert ((P>>16)==S//13).all()
assert (((P&0xFFFF)<5042)==(S%13==0)).all()
print("magic ok")

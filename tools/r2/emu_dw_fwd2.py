import numpy as np, torch, torch.nn.functional as F
def emu(S,B,H,W,C,x,sc,shf,w9,SEG,relu6=True):
    Ho,Wo=(H-1)//S+1,(W-1)//S+1
    OW=4 if S==1 else 2; NC=(OW-1)*S+3
    xq=x.reshape(-1); y=np.full(B*Ho*Wo*C,np.nan); cv=np.arange(C)
    actf=(lambda z:np.minimum(np.maximum(z,0),6)) if relu6 else (lambda z:np.maximum(z,0))
    ssum=np.zeros(C); ssq=np.zeros(C)
    rows=Ho; ncg=(Wo+OW-1)//OW; nseg=(rows+SEG-1)//SEG
    for b in range(B):
      for seg in range(nseg):
        r_a,r_b=seg*SEG,min(rows,seg*SEG+SEG)
        for wg in range(ncg):
            wo0=wg*OW; wi0=wo0*S-1
            mc=[1.0 if 0<=wi0+j<W else 0.0 for j in range(NC)]
            xco=[min(max(wi0+j,0),W-1)*C for j in range(NC)]
            xrow=W*C; xbase=b*H*xrow
            def loadrow(row):
                mr=1.0 if 0<=row<H else 0.0; o=xbase+min(max(row,0),H-1)*xrow
                return [actf(xq[o+xco[j]+cv]*sc+shf)*(mr*mc[j]) for j in range(NC)]
            if S==1: r0,r1=loadrow(r_a-1),loadrow(r_a)
            else: r0=loadrow(2*r_a-1)
            oy=((b*Ho+r_a)*Wo+wo0)*C; yrow=Wo*C
            for ho in range(r_a,r_b):
                if S==1: r2=loadrow(ho+1)
                else: r1,r2=loadrow(2*ho),loadrow(2*ho+1)
                for o in range(OW):
                    acc=r0[o*S]*w9[0]+r0[o*S+1]*w9[1]+r0[o*S+2]*w9[2]+r1[o*S]*w9[3]+r1[o*S+1]*w9[4]+r1[o*S+2]*w9[5]+r2[o*S]*w9[6]+r2[o*S+1]*w9[7]+r2[o*S+2]*w9[8]
                    if wo0+o<Wo:
                        y[oy+o*C+cv]=acc; ssum+=acc; ssq+=acc*acc
                if S==1: r0,r1=r1,r2
                else: r0=r2
                oy+=yrow
    return y.reshape(B,Ho,Wo,C),ssum,ssq
rng=np.random.default_rng(1)
for S in (1,2):
    for (B,H,W,C,SEG,r6) in [(2,6,6,4,2,True),(1,7,5,8,3,True),(2,8,12,4,8,False),(1,5,7,4,2,True),(1,14,14,4,8,True),(1,7,7,4,7,True)]:
        x=rng.normal(size=(B,H,W,C))*2; sc,shf=rng.uniform(.5,2,C),rng.normal(size=C)+1; w9=rng.normal(size=(9,C))
        y,s1,s2=emu(S,B,H,W,C,x,sc,shf,w9,SEG,r6)
        z=torch.tensor(x*sc+shf); a=(z.clamp(0,6) if r6 else z.clamp(min=0)).permute(0,3,1,2)
        w=torch.tensor(w9.reshape(3,3,C)).permute(2,0,1).reshape(C,1,3,3)
        ref=F.conv2d(a,w,None,S,1,1,C).permute(0,2,3,1).numpy()
        print(S,(B,H,W,C,SEG,r6),'nan' if np.isnan(y).any() else 'ok','%.1e %.1e %.1e'%(np.abs(y-ref).max(),np.abs(s1-ref.sum((0,1,2))).max(),np.abs(s2-(ref**2).sum((0,1,2))).max()))

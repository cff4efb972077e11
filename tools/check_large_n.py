import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from se3_diffusion_b200 import synthetic as fo
from gpu_common import engine
np.random.seed(3)
for (B,N) in [(1,512),(2,384),(3,100)]:
    r7=fo.random_frames(B, N, seed=3); f=fo.init_feats(r7); f["t"]=torch.full((B,),0.4,dtype=torch.float64)
    f["sc_ca_t"]=torch.tensor(np.random.randn(B,N,3)*15)
    e=engine("fp32"); ref={k:v.cpu().numpy() for k,v in e.forward(f).items()}
    e=engine("bf16x3"); out={k:v.cpu().numpy() for k,v in e.forward(f).items()}
    for k in ("rot_score","trans_score","atom37","psi"):
        d=np.abs(out[k]-ref[k]).max()/max(np.abs(ref[k]).max(),1e-9)
        print(B,N,k,f"{d:.2e}")

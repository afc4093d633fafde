import sys; sys.path.insert(0,'.')
import numpy as np, torch
from phase_guided_terrain_traversal_b200.policy import PolicyNet, reference_forward
def run(sizes, n=128):
    ks=[]; bs=[]
    for i,(a,b) in enumerate(zip(sizes[:-1], sizes[1:])):
        k=np.zeros((a,b),np.float32)
        for j in range(min(a,b)): k[j,j]=1.0
        ks.append(k); bs.append(np.zeros(b,np.float32))
    net=PolicyNet(sizes=sizes); net.set_params(ks,bs,None,None)
    obs=(np.arange(n)[:,None]*0.01+np.arange(sizes[0])[None,:]*0.001+0.5).astype(np.float32)
    out=net.act(torch.from_numpy(obs).cuda(), deterministic=True, want_logits=True)
    torch.cuda.synchronize()
    ref=reference_forward(ks,bs,obs,None,None,None,bf16_operands=True)
    d=(out["logits"].cpu()-ref["logits"]).abs().numpy()
    print(sizes, "max err", d.max(), "bad rows", np.where(d.max(1)>1e-2)[0][:20], "bad cols", np.where(d.max(0)>1e-2)[0])
    if d.max()>1e-2:
        print(" got", out["logits"].cpu().numpy()[0,:8], "\n ref", ref["logits"].numpy()[0,:8])
run([171,128,24]); run([171,128,128,24]); run([171,512,24]); run([171,512,256,128,24]); run([32,128,24]); run([171,512,256,128,24], n=4096)

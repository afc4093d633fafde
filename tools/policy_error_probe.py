import sys; sys.path.insert(0,'.')
import numpy as np, torch
from phase_guided_terrain_traversal_b200.policy import PolicyNet, reference_forward
g = np.load("tests/golden/policy177.npz")
ks, bs = [g[f"kernel{i}"] for i in range(4)], [g[f"bias{i}"] for i in range(4)]
net = PolicyNet(); net.set_params(ks, bs, g["mean"], g["std"])
for n in (128, 4096):
    rng = np.random.default_rng(n)
    obs = (g["mean"] + g["std"] * rng.normal(size=(n, 171)).clip(-3, 3)).astype(np.float32)
    eps = rng.normal(size=(n, 12)).astype(np.float32)
    out = net.act(torch.from_numpy(obs).cuda(), eps=torch.from_numpy(eps).cuda(), want_logits=True)
    torch.cuda.synchronize()
    ref_bf = reference_forward(ks, bs, obs, g["mean"], g["std"], eps, bf16_operands=True)
    ref_32 = reference_forward(ks, bs, obs, g["mean"], g["std"], eps)
    d = (out["logits"].cpu() - ref_bf["logits"]).abs()
    d32 = (out["logits"].cpu() - ref_32["logits"]).abs()
    dr = (ref_bf["logits"] - ref_32["logits"]).abs()
    q = torch.tensor([0.5, 0.9, 0.99, 0.999, 0.9999])
    print(n, "vs bf ref quantiles", torch.quantile(d.flatten(), q).tolist(), "max", d.max().item(), "rows>1e-3:", int((d.max(1).values > 1e-3).sum()))
    print(n, "vs f32 ref max", d32.max().item(), " bf-ref vs f32-ref max", dr.max().item(), "logit scale", ref_32["logits"].abs().max().item())

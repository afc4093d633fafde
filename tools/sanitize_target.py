"""What compute-sanitizer runs over (tools/runs/sanitize.sh): __graft_entry__.smoke() - both physics generations, the task kernel, reset, randomise,
the tcgen05 policy, the graphed rollout - plus one small forward + backward of the learner's whole-MLP path (pgtt_bsplit / pgtt_bgemm / pgtt_bsum)
with the fused gather + normalisation, checked against torch."""
import sys; sys.path.insert(0, ".")
import numpy as np, torch
import __graft_entry__ as g
g.smoke()
from phase_guided_terrain_traversal_b200 import ppo
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev); gen.manual_seed(1)
T, S, mb, width, ld = 5, 40, 24, 37, 40
data = torch.randn(T, S, ld, device=dev, generator=gen)
idx = torch.randperm(S, device=dev, generator=gen)[:mb]
mean, inv = torch.randn(width, device=dev, generator=gen), torch.rand(width, device=dev, generator=gen) + 0.5
ks = [(torch.randn(a, b, device=dev, generator=gen) / np.sqrt(a)).requires_grad_() for a, b in ((37, 96), (96, 40), (40, 5))]
bs = [torch.zeros(b, device=dev, requires_grad=True) for b in (96, 40, 5)]
spec = ppo.GatherInput(data, idx, T, width, mean, inv)
y = ppo.mlp(spec, ks, bs, None, True); y.square().sum().backward(); got = [p.grad.clone() for p in ks + bs]
for p in ks + bs: p.grad = None
y2 = ppo.mlp(spec.materialise(), ks, bs, None, False); y2.square().sum().backward()
torch.cuda.synchronize()
assert float((y - y2).abs().max()) < 1e-4 * float(y2.abs().max())
for a, p in zip(got, ks + bs): assert float((a - p.grad).abs().max()) < 1e-4 * float(p.grad.abs().max())
print("mlp ok")

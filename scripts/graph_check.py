"""CUDA-graph replay of the train step must reproduce the eager step bit for bit (same seeds, same inputs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fudanocr_b200.model.tbsrn import TBSRN
from fudanocr_b200.trainer import TBSRNTrainer
B = 32
def run(use_graph):
    torch.manual_seed(1234)
    m = TBSRN().cuda().train()
    tr = TBSRNTrainer(m, use_graph=use_graph)
    g = torch.Generator(device="cuda").manual_seed(5)
    losses = []
    for i in range(5):
        lr = torch.rand(B, 3, 16, 64, device="cuda", generator=g); hr = torch.rand(B, 3, 32, 128, device="cuda", generator=g)
        losses.append(tr.step(lr, hr, seed=1000 + i).clone())
    torch.cuda.synchronize()
    return torch.cat(losses), tr.flat_p.clone(), tr.kernel_launches, tr._graphs is not None
l0, p0, n0, g0 = run(False)
l1, p1, n1, g1 = run(True)
print("eager losses", l0.tolist()); print("graph losses", l1.tolist())
print("graphs used:", g0, g1, "launches", n0, n1)
print("params equal:", torch.equal(p0, p1), "max diff", (p0 - p1).abs().max().item(), "losses equal:", torch.equal(l0, l1))
assert g1 and not g0 and torch.equal(p0, p1) and torch.equal(l0, l1) and n0 == n1

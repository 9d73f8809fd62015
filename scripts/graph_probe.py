"""How much of the step is launch gaps?  Capture one TBSRN train step (fixed dropout seed) in a CUDA graph and replay it."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fudanocr_b200.model.tbsrn import TBSRN
from fudanocr_b200.trainer import TBSRNTrainer
B = 256
torch.manual_seed(1234)
m = TBSRN().cuda().train()
tr = TBSRNTrainer(m)
lr = torch.rand(B, 3, 16, 64, device="cuda"); hr = torch.rand(B, 3, 32, 128, device="cuda")
for i in range(3):
    tr.step(lr, hr, seed=i)
torch.cuda.synchronize()
def timed(fn, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("eager ms/step", timed(lambda i: tr.step(lr, hr, seed=100 + i)))
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    tr.step(lr, hr, seed=7)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    tr.step(lr, hr, seed=7)
print("graph ms/step", timed(lambda i: g.replay()))
print("loss", tr.loss.item())

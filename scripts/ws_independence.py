"""Two trainers on identical weights/inputs, one with a garbage-filled workspace: gradients must be bit-identical
(no kernel may read workspace bytes it did not write in the same step)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fudanocr_b200.model.tbsrn import TBSRN
from fudanocr_b200.trainer import TBSRNTrainer
from oracle import synth, tbsrn_oracle as O
dev = "cuda"
sd = synth.synth_state_dict(synth.load_spec("tbsrn"), 1234, O.tps_buffers())
B = 4
lr, hr = synth.synth_images(B); lr, hr = lr.to(dev), hr.to(dev)
gs = []
for k in range(4):
    m = TBSRN(STN=False).to(dev)
    m.load_state_dict({a: b for a, b in sd.items() if not (a.startswith("stn_head") or a.startswith("tps"))})
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout): mod.p = 0.0
    m.train()
    tr = TBSRNTrainer(m, lr=0.0)
    ws = m._workspace(B, torch.device(dev))
    if k == 1: ws.view(torch.int16).random_(-30000, 30000)
    if k == 2: ws.fill_(0)
    hr_k = torch.nextafter(hr, hr + 1) if k == 3 else hr   # k = 3: targets moved by one ulp -> d_sr moves by ~1 ulp
    tr.step(lr, hr_k); torch.cuda.synchronize()
    gs.append(tr.flat_g.clone()); names = [m._slot_names[i] for i in m._grad_slots]; tensors, _ = m._slots(); slots = m._grad_slots
for a, b in ((0, 1), (0, 2), (0, 3)):
    d = (gs[a] - gs[b]).abs()
    print(a, b, "identical" if torch.equal(gs[a], gs[b]) else f"DIFFER max {d.max().item():.3e} rel {((gs[a]-gs[b]).norm()/gs[a].norm()).item():.3e}")
    if not torch.equal(gs[a], gs[b]):
        off = 0
        for i, k in zip(slots, names):
            n = tensors[i].numel()
            x, y = gs[a][off:off+n], gs[b][off:off+n]; off += (n + 3) // 4 * 4
            r = ((x - y).norm() / (x.norm() + 1e-20)).item()
            if r > 1e-3: print("   ", k, f"{r:.2e}")

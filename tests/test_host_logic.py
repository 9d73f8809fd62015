"""CPU tests of the host-side logic: C-ABI surface, drop-in module surface, data-parallel plumbing (gloo)."""
import os
import re
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from fudanocr_b200 import _lib as L
    hdr = open(os.path.join(ROOT, "include", "focr.h")).read()
    names = set(re.findall(r"\b(focr_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(L.lib, n), f"{n} declared in include/focr.h but not exported"
    assert L.lib.focr_version() >= 100
    assert L.lib.focr_tbsrn_num_slots(5) == 227


def test_no_cpu_fallback_for_the_engine():
    """the product path must fail loudly without a GPU: no eager/PyTorch fallback"""
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.model.tbsrn import TBSRN
    m = TBSRN()
    with pytest.raises(L.FocrError):
        m(torch.rand(2, 3, 16, 64))


def test_dropin_module_surface():
    from fudanocr_b200.model.tbsrn import TBSRN
    from fudanocr_b200.interfaces.parallel import DataParallel
    from oracle import synth, tbsrn_oracle as O
    spec = synth.load_spec("tbsrn")
    m = TBSRN(scale_factor=2, width=128, height=32, STN=True, srb_nums=5, mask=False, hidden_units=32)
    assert list(m.state_dict().keys()) == list(spec.keys())
    assert all(list(v.shape) == spec[k] for k, v in m.state_dict().items())
    sd = synth.synth_state_dict(spec, 7, O.tps_buffers())
    m.load_state_dict(sd)  # reference checkpoints load unchanged
    named = dict(m.named_parameters())
    named.update(dict(m.named_buffers()))
    for k in m._slot_names:  # every tensor handed to the engine by raw pointer must be dense fp32 / int64
        assert named[k].is_contiguous() and named[k].dtype in (torch.float32, torch.int64), k
    w = DataParallel(m)
    assert w.module is m and len(list(w.parameters())) == len(list(m.parameters()))
    pre = {"module." + k for k in spec}
    assert set(w.state_dict().keys()) == pre  # 'module.'-prefixed keys as under nn.DataParallel (base.py:184-187)
    m2 = TBSRN(STN=False)
    assert not any(k.startswith("stn_head") or k.startswith("tps") for k in m2.state_dict())
    with pytest.raises(NotImplementedError):
        TBSRN(mask=True)


def test_shard_bounds_cover_batch():
    from fudanocr_b200.interfaces.parallel import shard_bounds, shard_batch
    for n in (256, 7, 1024):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    x, labels = torch.arange(10), list("abcdefghij")
    a, b = shard_batch([x, labels], 1, 2)
    assert a.tolist() == [5, 6, 7, 8, 9] and b == list("fghij")


WORKER = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from fudanocr_b200.interfaces.parallel import shard_batch, allreduce_mean_
    from oracle import synth, tbsrn_oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% sys.argv[1], rank=int(sys.argv[2]), world_size=2)
    rank = dist.get_rank()
    torch.set_num_threads(2)
    spec = {k: v for k, v in synth.load_spec("tbsrn").items()
            if not (k.startswith("stn_head") or k.startswith("tps") or any(k.startswith("block%%d." %% i) for i in (3, 4, 5, 6)))}
    # 1-SRB network: block2 = SRB, block7/8 of the reference become block3/4 here
    spec = {k.replace("block7.", "block3.").replace("block8.", "block4."): v for k, v in spec.items()}
    sd = synth.synth_state_dict(spec, 1234)
    lr, hr = synth.synth_images(4)
    keys = [k for k in sd if sd[k].is_floating_point() and not O.is_buffer(k) and "gru" not in k
            and not k.startswith("conv.") and not k.startswith("bn.") and "compress" not in k]
    def shard_grads(r):
        l, h = shard_batch([lr, hr], r, 2)
        leaf = {k: (sd[k].clone().requires_grad_(True) if k in keys else sd[k]) for k in sd}
        sr = O.tbsrn_forward(leaf, l, training=True, stn=False, srb_nums=1)
        (torch.nn.functional.mse_loss(sr, h) * 100).backward()
        return torch.cat([leaf[k].grad.reshape(-1) for k in keys])
    flat = shard_grads(rank)
    allreduce_mean_(flat)
    expect = (shard_grads(0) + shard_grads(1)) / 2   # single-process emulation of both replicas
    assert torch.allclose(flat, expect, rtol=1e-5, atol=1e-7), (flat - expect).abs().max()
    # clip AFTER the reduce -> identical step on every rank
    norm = flat.norm()
    gathered = [torch.zeros_like(norm) for _ in range(2)]
    dist.all_gather(gathered, norm)
    assert gathered[0] == gathered[1]
    dist.destroy_process_group()
    print("ok", rank)
""")


def test_two_rank_gloo_gradient_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [subprocess.Popen([sys.executable, str(script), str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_focus_loss_dropin_surface():
    """the frozen recogniser container carries the reference's state_dict; the loss refuses CPU tensors"""
    import types
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.loss.stroke_focus_loss import StrokeFocusLoss
    from fudanocr_b200.loss.transformer_english_decomposition import Transformer
    from oracle import synth, focus_oracle as FO
    spec = synth.load_spec("focus")
    t = Transformer("tg")
    assert list(t.state_dict().keys()) == list(spec.keys())
    assert all(list(v.shape) == spec[k] for k, v in t.state_dict().items())
    stt = Transformer("stt")
    assert "embedding_word.lut.weight" in stt.state_dict() and stt.state_dict()["generator_word.proj.weight"].shape[0] == 37
    sd = {"module." + k: v for k, v in t.state_dict().items()}          # DataParallel-prefixed asset, as shipped
    crit = StrokeFocusLoss(types.SimpleNamespace(text_focus=True, stroke_lambda=50), decomposition=FO.synth_decomposition(),
                           transformer_state_dict=sd)
    assert not any(p.requires_grad for p in crit.transformer.parameters())
    ln, inp, gt = crit.label_stroke_encoder(["ab3", "Hello"])
    o_ln, o_inp, o_gt = FO.label_stroke_encoder(["ab3", "Hello"], FO.synth_decomposition())
    assert torch.equal(ln, o_ln) and torch.equal(inp, o_inp) and torch.equal(gt, o_gt)
    with pytest.raises(L.FocrError):
        crit(torch.rand(2, 3, 32, 128), torch.rand(2, 3, 32, 128), ["ab3", "Hello"])


def test_text_focus_loss_dropin_surface():
    import types
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.loss.text_focus_loss import TextFocusLoss, str_filt
    from fudanocr_b200.loss.transformer_english_decomposition import Transformer
    from oracle import synth, textfocus_oracle as TF
    spec = synth.load_spec("textfocus")
    t = Transformer("stt")
    assert list(t.state_dict().keys()) == list(spec.keys())
    assert all(list(v.shape) == spec[k] for k, v in t.state_dict().items())
    crit = TextFocusLoss(types.SimpleNamespace(text_focus=True), confuse_counts=TF.synth_confuse_counts(),
                         transformer_state_dict={"module." + k: v for k, v in t.state_dict().items()})
    assert torch.equal(crit.weight_table, TF.confuse_weight_table(TF.synth_confuse_counts()))
    labels = ["B200", "text-Focus!", "a"]
    filt = [str_filt(s, "lower") + "-" for s in labels]
    assert filt == [TF.str_filt(s, "lower") + "-" for s in labels] == ["b200-", "textfocus-", "a-"]
    for a, b in zip(crit.label_encoder(filt), TF.label_encoder(filt)):
        assert torch.equal(a, b)
    with pytest.raises(L.FocrError):
        crit(torch.rand(3, 3, 32, 128), torch.rand(3, 3, 32, 128), labels)


SLD_WORKER = textwrap.dedent("""
    import ctypes, os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% sys.argv[1], rank=int(sys.argv[2]), world_size=2)
    rank = dist.get_rank()
    torch.set_num_threads(3)
    import _recog_mock as M
    from fudanocr_b200.model import recog_ops as ops
    for name in dir(M):
        if not name.startswith("_") and callable(getattr(M, name)) and hasattr(ops, name) and name != "install":
            setattr(ops, name, getattr(M, name))
    ops.BF = torch.float32
    ops.require_cuda = lambda t: None
    def view(ptr, n):
        return np.ctypeslib.as_array((ctypes.c_float * n).from_address(ptr))
    def adadelta(table, n_chunks, gscale, lr, rho, eps, wd):            # host twin of adadelta_kernel over the same chunk table
        for p, g, sq, acc, n in table.tolist():
            P, G, S, A = view(p, n), view(g, n) * np.float32(gscale), view(sq, n), view(acc, n)
            if wd:
                G = G + np.float32(wd) * P
            S[:] = rho * S + (1 - rho) * G * G
            d = np.sqrt(A + eps) / np.sqrt(S + eps) * G
            A[:] = rho * A + (1 - rho) * d * d
            P[:] = P - lr * d
    ops.adadelta_step = adadelta
    from oracle import sld_oracle as SO, synth
    from fudanocr_b200.interfaces.parallel import shard_batch
    from fudanocr_b200.model.transformer import Transformer
    from fudanocr_b200.trainer_sld import SLDTrainer
    model = Transformer("stroke")
    sd = synth.synth_state_dict(synth.load_spec("sld"), 1234 + rank)     # ranks start DIFFERENT: the trainer must broadcast rank 0's
    model.load_state_dict(sd, strict=False)
    model.dropout_p = 0.0
    image, strings = SO.synth_batch(4)
    img, strs = shard_batch([image, strings], rank, 2)
    length, text_input, text_gt = SO.converter_stroke(strs)
    tr = SLDTrainer(model)
    assert tr.world == 2
    before = tr.flat_p.clone()
    sums = [torch.zeros(1, dtype=torch.float64) for _ in range(2)]
    dist.all_gather(sums, before.double().sum().reshape(1))
    assert sums[0] == sums[1]                                           # identical start after the broadcast
    loss = tr.step(img, length, text_input, text_gt)
    losses = [torch.zeros(1) for _ in range(2)]
    dist.all_gather(losses, loss.reshape(1).float())
    assert losses[0] != losses[1]                                       # different shards
    for buf in (tr.flat_g, tr.flat_p):                                  # one summed gradient, one identical step on every rank
        got = [torch.zeros(1, dtype=torch.float64) for _ in range(2)]
        dist.all_gather(got, buf.double().abs().sum().reshape(1))
        assert got[0] == got[1], got
    assert not torch.equal(before, tr.flat_p) and torch.isfinite(tr.flat_p).all()
    assert dict(model.named_parameters())["generator_word.proj.weight"].data_ptr() >= tr.flat_p.data_ptr()
    dist.destroy_process_group()
    print("ok", rank)
""")


def test_two_rank_gloo_recogniser_trainer(tmp_path):
    """the data-parallel step of the trainable recognisers (trainer_sld.py) on 2 gloo ranks: SLDTrainer itself, with the kernel
    wrappers swapped for the CPU stand-ins of tests/_recog_mock.py (test infrastructure) - broadcast of rank 0's weights, one
    summed flat gradient, identical Adadelta step on both ranks"""
    script = tmp_path / "sld_worker.py"
    script.write_text(SLD_WORKER % (ROOT, ROOT))
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [subprocess.Popen([sys.executable, str(script), str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o


def test_collate_host_side_packing_and_no_cpu_fallback():
    """host half of the device-side collate (fudanocr_b200/dataset): ragged crops -> one packed uint8 buffer + (offset, h, w)
    records; without a CUDA device the collate refuses instead of falling back to Pillow"""
    import numpy as np
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.dataset import dataset as D
    rs = np.random.RandomState(0)
    crops = [rs.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w in ((5, 7), (1, 1), (32, 128))]
    buf, meta, max_h, max_w = D.pack_crops(crops)
    assert (max_h, max_w) == (32, 128) and meta.shape == (3, 3) and meta.dtype == torch.int64
    assert meta[:, 0].tolist() == [0, 5 * 7 * 3, 5 * 7 * 3 + 3] and buf.numel() == 5 * 7 * 3 + 3 + 32 * 128 * 3
    for (off, h, w), c in zip(meta.tolist(), crops):
        assert np.array_equal(buf.numpy()[off:off + h * w * 3].reshape(h, w, 3), c)
    # non-contiguous views and torch tensors are accepted; wrong dtypes / ranks / empty crops are not
    D.pack_crops([crops[2][::2, ::2], torch.from_numpy(crops[0])])
    for bad in (np.zeros((4, 4), np.uint8), np.zeros((4, 4, 3), np.float32), np.zeros((0, 4, 3), np.uint8)):
        with pytest.raises(ValueError):
            D.pack_crops([bad])
    with pytest.raises(NotImplementedError):
        D.alignCollate_real(mask=True)
    if not torch.cuda.is_available():
        with pytest.raises(L.FocrError):
            D.resize_normalize_batch(crops, (128, 32))
        with pytest.raises(L.FocrError):
            D.alignCollate_real(imgH=32, imgW=128, down_sample_scale=2)([(crops[0], crops[1], "a")])


def test_ctc_loss_host_side_argument_handling():
    """torch.nn.CTCLoss call forms: 1-D concatenated targets are re-padded on the host; CPU tensors are refused"""
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.loss import ctc_loss as C
    padded = C._pad_targets(torch.tensor([3, 4, 5, 9, 1, 1]), torch.tensor([3, 0, 1, 2]))
    assert padded.tolist() == [[3, 4, 5], [0, 0, 0], [9, 0, 0], [1, 1, 0]]
    assert C._pad_targets(torch.tensor([], dtype=torch.long), torch.tensor([0, 0])).shape == (2, 1)
    with pytest.raises(ValueError):
        C.CTCLoss(reduction="median")
    with pytest.raises(L.FocrError):
        C.ctc_loss(torch.zeros(4, 2, 5), torch.ones(2, 2, dtype=torch.long), [4, 4], [2, 2])


def test_header_is_plain_c(tmp_path):
    """include/focr.h is the drop-in boundary: it must compile as C99 (no C++ / torch types in any signature)"""
    import shutil
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text('#include "focr.h"\nint main(void) { return focr_version() > 0 ? 0 : 1; }\n')
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_label_converters_match_the_reference_semantics():
    """fudanocr_b200/util_recog.py vs the oracle restatements of util.converter (both pinned to the unmodified reference
    converters inside oracle/make_golden_sld.py / make_golden_ids.py)"""
    from fudanocr_b200 import util_recog as U
    from oracle import ids_oracle as IO, sld_oracle as SO
    table = {"甲": "25112", "乙": "5", "丙": "12534"}
    labels = [["甲"], ["乙"], ["丙"], ["乙"]]
    length, text_input, text_gt, back = U.converter_sld("stroke", labels, table, device=None)
    o_len, o_in, o_gt = SO.converter_stroke([table[l[0]] + "$" for l in labels])
    assert torch.equal(length, o_len) and torch.equal(text_input, o_in) and torch.equal(text_gt, o_gt) and back is labels
    assert text_input[:, 0].eq(0).all() and text_gt[5] == 6            # start column, '$' closes the first sample
    a2n = {"START": 0, "天": 1, "地": 2, "人": 3, "END": 4}
    strings = ["天地#", "人#", "地人天#"]
    length, text_input, text_gt, _ = U.converter_ids(strings, a2n, device=None)
    import oracle.ids_oracle as _io
    old = _io.N_CLASS
    _io.N_CLASS = 5                                                    # the oracle writes END as N_CLASS - 1
    try:
        o_len, o_in, o_gt = IO.converter([[a2n[c] for c in s[:-1]] + [0] for s in strings])
    finally:
        _io.N_CLASS = old
    assert torch.equal(length, o_len) and torch.equal(text_input, o_in) and torch.equal(text_gt, o_gt)
    assert text_gt.tolist() == [1, 2, 4, 3, 4, 2, 3, 1, 4]
    with pytest.raises(ValueError):
        U.converter_sld("stroke", labels, None, device=None)
    with pytest.raises(KeyError):
        U.converter_sld("stroke", [["丁"]], table, device=None)


def test_cosine_warm_restarts_matches_torch_scheduler():
    from fudanocr_b200.util_recog import cosine_warm_restarts_lr
    for T_0, T_mult in ((10, 1), (3, 2)):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.Adadelta([p], lr=1.0, rho=0.9, weight_decay=1e-4)
        sch = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(opt, T_0=T_0, T_mult=T_mult)
        for epoch in range(0, 45):
            assert abs(opt.param_groups[0]["lr"] - cosine_warm_restarts_lr(epoch, 1.0, T_0, T_mult)) < 1e-9, (T_0, T_mult, epoch)
            opt.step()
            sch.step()


def test_recogniser_ops_have_no_cpu_fallback():
    """every kernel wrapper of the trainable recognisers refuses CPU tensors (the product path never computes in torch)"""
    from fudanocr_b200 import _lib as L
    from fudanocr_b200.model import recog_ops as ops
    from fudanocr_b200.model.transformer import Transformer
    from fudanocr_b200.model.ids_transformer import Transformer as IDSTransformer
    bf = torch.bfloat16
    x = torch.zeros(2, 16, 16, 64, dtype=bf)
    w, b = torch.zeros(64, 64, 3, 3), torch.zeros(64)
    tok = torch.zeros(128, 1024, dtype=bf)
    calls = [
        lambda: ops.conv_fwd(x, w, b), lambda: ops.conv_dgrad(x, w), lambda: ops.conv_wgrad(x, x, (64, 64, 3, 3)),
        lambda: ops.conv_first_fwd(torch.zeros(2, 3, 32, 32), torch.zeros(64, 3, 3, 3), b),
        lambda: ops.bn_train_fwd(x, b, b, b.clone(), b.clone(), torch.zeros((), dtype=torch.long), 2),
        lambda: ops.bn_eval_fwd(x, b, b, b, b, 0), lambda: ops.add_relu(x, x), lambda: ops.relu_bwd(x, x), lambda: ops.maxpool_fwd(x),
        lambda: ops.dropout(x, 0.1, 1, 2), lambda: ops.linear_fwd(tok, torch.zeros(64, 1024), b),
        lambda: ops.linear_dgrad(tok, torch.zeros(1024, 64)), lambda: ops.linear_wgrad(tok, tok),
        lambda: ops.mha_fwd(tok, tok, tok, 2, 4, 256, 8, 8, 1, 0.0, 0, 0),
        lambda: ops.ln_fwd(tok, tok, torch.ones(1024), torch.zeros(1024)),
        lambda: ops.embed_fwd(torch.zeros(2, 4, dtype=torch.long), torch.zeros(7, 512), 128, 0.0, 0, 0),
        lambda: ops.packed_ce(torch.zeros(128, 64), 2, 4, 7, torch.tensor([4, 4]), torch.zeros(8, dtype=torch.long)),
        lambda: ops.l2norm_fwd(torch.zeros(128, 2048)),
        lambda: Transformer("stroke")(torch.zeros(2, 3, 32, 32), torch.tensor([2, 2]), torch.zeros(2, 2, dtype=torch.long)),
        lambda: IDSTransformer(20)(torch.zeros(4, 3, 32, 256), torch.tensor([2] * 4), torch.zeros(4, 2, dtype=torch.long)),
    ]
    for i, c in enumerate(calls):
        with pytest.raises(L.FocrError):
            c()
    with pytest.raises(ValueError):
        Transformer("stroke").encode(torch.zeros(2, 3, 32, 72))       # 16 x 36 maps do not cut into 128-pixel TMA boxes
    with pytest.raises(L.FocrError):
        Transformer("stroke").encode(torch.zeros(2, 3, 32, 320))      # 16 x 160 maps tile (32 x 4 boxes); CPU tensors still refused


CLIP_WORKER = textwrap.dedent("""
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from fudanocr_b200.loss import clip_contrastive as CC
    from oracle import clip_oracle as CO
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% sys.argv[1], rank=int(sys.argv[2]), world_size=2)
    rank = dist.get_rank()
    torch.set_num_threads(2)

    def fused_standin(image, text, logit_scale, gt, want_grad):      # the kernel's contract, in torch (test stand-in: no GPU here)
        with torch.enable_grad():                                     # (called from inside Function.forward)
            i, t, s = (x.detach().double().requires_grad_(True) for x in (image, text, logit_scale))
            loss, _ = CO.contrastive_loss(i, t, s[0], gt)
            loss.backward()
        return loss.detach().float().reshape(1), i.grad.float(), t.grad.float(), s.grad.float()
    CC._fused = fused_standin

    torch.manual_seed(0)
    B, Din, D = 8, 12, 16
    x_img, x_txt = torch.randn(B, Din), torch.randn(B, Din)
    labels = list("abacbdae")
    gt = CC.ground_truth_from_labels(labels)

    def towers():
        torch.manual_seed(1)
        return torch.nn.Linear(Din, D), torch.nn.Linear(Din, D), torch.nn.Parameter(torch.tensor(2.0))
    # single process, whole batch (what nn.DataParallel's main device computes in the reference)
    fi, ft, ls = towers()
    ref, _ = CO.contrastive_loss(fi(x_img), ft(x_txt), ls, gt)
    ref.backward()
    for mode in ("sum", "mean"):
        gi, gtw, gls = towers()
        sl = slice(rank * 4, rank * 4 + 4)
        loss = CC.clip_contrastive_loss(gi(x_img[sl]), gtw(x_txt[sl]), gls, gt, grad_reduce=mode)
        loss.backward()
        assert abs(float(loss) - float(ref)) < 1e-5, (float(loss), float(ref))
        for p, q in ((gi.weight, fi.weight), (gi.bias, fi.bias), (gtw.weight, ft.weight), (gls, ls)):
            g = p.grad.clone()
            dist.all_reduce(g)                     # what the data-parallel trainer does with the flat gradient
            if mode == "mean":
                g /= 2
            assert torch.allclose(g, q.grad, rtol=1e-4, atol=1e-6), (mode, (g - q.grad).abs().max())
    dist.destroy_process_group()
    print("ok", rank)
""")


def test_two_rank_gloo_clip_contrastive_head(tmp_path):
    """the data-parallel form of the CCR-CLIP contrastive head: all-gathered features, the global B x B loss on every rank, local
    gradient rows - after the usual gradient all-reduce (sum, or mean) every parameter gradient equals the single-process one"""
    script = tmp_path / "clip_worker.py"
    script.write_text(CLIP_WORKER % ROOT)
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = [subprocess.Popen([sys.executable, str(script), str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o

"""Building blocks of the trainable ResNet + Transformer recognisers (csrc/recog_ops.cu) through
the C ABI vs torch fp32 ops evaluated on the same bf16-rounded operands on the GPU (torch = checker only).
Tolerances: bf16 outputs 1e-2 of the largest reference entry, fp32 outputs 1e-4 .. 1e-5, integer / mask work exact."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"
BF = torch.bfloat16


def _lib():
    from fudanocr_b200 import _lib as L
    return L


def _ws(nbytes):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=DEV)


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


def _sync(L):
    L.check(L.lib.focr_sync_check(L.cur_stream()))


def _mha(L, q, k, v, B, H, dk, Tq, Tk, causal, p, seed=7, sid=3):
    out = torch.empty(B * Tq, H * dk, dtype=BF, device=DEV)
    amap = torch.empty(B, H, Tq, Tk, dtype=torch.float32, device=DEV)
    L.check(L.lib.focr_mha_small_fwd(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                                     out.data_ptr(), out.stride(0), amap.data_ptr(), B, H, dk, Tq, Tk, causal, p, seed, sid,
                                     L.cur_stream()), "mha_small_fwd")
    return out, amap


def _mha_ref(q, k, v, B, H, dk, Tq, Tk, causal, keep=None, ks=1.0):
    qh = q.float().view(B, Tq, H, dk).transpose(1, 2)
    kh = k.float().view(B, Tk, H, dk).transpose(1, 2)
    vh = v.float().view(B, Tk, H, dk).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dk)
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(Tq, Tk, device=DEV, dtype=torch.bool), 1), float("-inf"))
    p = F.softmax(s, -1)
    if keep is not None:
        p = p * keep * ks
    o = (p @ vh).transpose(1, 2).reshape(B * Tq, H * dk)
    return o, p


@pytest.mark.parametrize("B,H,dk,Tq,Tk,causal", [(3, 4, 256, 17, 17, 1), (3, 4, 256, 17, 256, 0), (2, 16, 64, 9, 40, 0),
                                                 (2, 8, 128, 31, 31, 1), (1, 4, 256, 1, 1, 1),
                                                 (2, 4, 256, 23, 2560, 0), (1, 4, 64, 300, 300, 1), (1, 2, 128, 9, 2561, 0)])
def test_mha_small_fwd_bwd_no_dropout(B, H, dk, Tq, Tk, causal):
    """the last three cases exceed one CTA's shared-memory score tile (2 560 = the image tokens of 32 x 320 crops): row-split
    forward, row pass + key pass backward with a dS workspace"""
    L = _lib()
    g = torch.Generator(device=DEV).manual_seed(B * 100 + Tq)
    D = H * dk
    # q / k / v as column slices of fused projection outputs (leading dimension != width), like the model passes them
    qkv = torch.randn(B * Tq, 3 * D, device=DEV, generator=g).to(BF)
    kv = torch.randn(B * Tk, 2 * D, device=DEV, generator=g).to(BF)
    q = qkv[:, :D]
    k, v = (qkv[:, D:2 * D], qkv[:, 2 * D:]) if causal else (kv[:, :D], kv[:, D:])
    out, amap = _mha(L, q, k, v, B, H, dk, Tq, Tk, causal, 0.0)
    _sync(L)
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    o_ref, p_ref = _mha_ref(qr, kr, vr, B, H, dk, Tq, Tk, causal)
    assert (amap - p_ref).abs().max().item() < 2e-5
    assert _rel(out, o_ref) < 1e-2
    d_out = torch.randn(B * Tq, D, device=DEV, generator=g).to(BF)
    o_ref.backward(d_out.float())
    dq = torch.empty(B * Tq, D, dtype=BF, device=DEV)
    dkv = torch.empty(B * Tk, 2 * D, dtype=BF, device=DEV)
    dk_, dv = dkv[:, :D], dkv[:, D:]
    nws = L.lib.focr_mha_small_bwd_workspace_bytes(B, H, Tq, Tk)
    assert (nws > 0) == (Tq * Tk * 8 > 200 * 1024)
    ws = _ws(nws)
    L.check(L.lib.focr_mha_small_bwd_ws(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                                        d_out.data_ptr(), d_out.stride(0), amap.data_ptr(), dq.data_ptr(), dq.stride(0),
                                        dk_.data_ptr(), dk_.stride(0), dv.data_ptr(), dv.stride(0), B, H, dk, Tq, Tk, causal, 0.0,
                                        ws.data_ptr(), ws.numel(), L.cur_stream()), "mha_small_bwd")
    _sync(L)
    if nws:  # the no-workspace entry refuses what it cannot do
        assert L.lib.focr_mha_small_bwd(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                                        d_out.data_ptr(), d_out.stride(0), amap.data_ptr(), dq.data_ptr(), dq.stride(0),
                                        dk_.data_ptr(), dk_.stride(0), dv.data_ptr(), dv.stride(0), B, H, dk, Tq, Tk, causal, 0.0,
                                        L.cur_stream()) != 0
    assert _rel(dq, qr.grad) < 1e-2
    assert _rel(dk_, kr.grad) < 1e-2
    assert _rel(dv, vr.grad) < 1e-2


@pytest.mark.parametrize("Tk", [256, 2560])
def test_mha_small_dropout_mask_is_the_oracle_rng_and_gradients_follow_it(Tk):
    from oracle import dropout_rng as R
    L = _lib()
    B, H, dk, Tq, p, seed, sid = 2, 4, 256, 13, 0.1, 99, 5
    D = H * dk
    g = torch.Generator(device=DEV).manual_seed(5)
    q = torch.randn(B * Tq, D, device=DEV, generator=g).to(BF)
    k = torch.randn(B * Tk, D, device=DEV, generator=g).to(BF)
    v = torch.randn(B * Tk, D, device=DEV, generator=g).to(BF)
    out, amap = _mha(L, q, k, v, B, H, dk, Tq, Tk, 0, p, seed, sid)
    _sync(L)
    keep = torch.from_numpy(R._keep(R.drop_key(seed, sid), np.arange(B * H * Tq * Tk, dtype=np.uint64), R.thresh16(p))
                            .reshape(B, H, Tq, Tk)).to(DEV)
    assert torch.equal(amap != 0, keep)          # softmax entries are never exactly 0 at these sizes
    assert abs(1.0 - keep.float().mean().item() - p) < 0.01
    qr, kr, vr = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    o_ref, p_ref = _mha_ref(qr, kr, vr, B, H, dk, Tq, Tk, 0, keep.float(), R.keep_scale(p))
    assert (amap - p_ref).abs().max().item() < 2e-5 and _rel(out, o_ref) < 1e-2
    d_out = torch.randn(B * Tq, D, device=DEV, generator=g).to(BF)
    o_ref.backward(d_out.float())
    dq, dk_, dv = (torch.empty_like(t) for t in (q, k, v))
    ws = _ws(L.lib.focr_mha_small_bwd_workspace_bytes(B, H, Tq, Tk))
    L.check(L.lib.focr_mha_small_bwd_ws(q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D, d_out.data_ptr(), D, amap.data_ptr(),
                                        dq.data_ptr(), D, dk_.data_ptr(), D, dv.data_ptr(), D, B, H, dk, Tq, Tk, 0, p,
                                        ws.data_ptr(), ws.numel(), L.cur_stream()), "mha_small_bwd")
    _sync(L)
    assert _rel(dq, qr.grad) < 1e-2 and _rel(dk_, kr.grad) < 1e-2 and _rel(dv, vr.grad) < 1e-2
    # argument errors are codes, not crashes
    assert L.lib.focr_mha_small_fwd(q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D, out.data_ptr(), D, amap.data_ptr(), B, H, 96,
                                    Tq, Tk, 0, 0.0, 0, 0, L.cur_stream()) != 0


def _ln_ref(x, a, b, eps=1e-6):   # stroke-level-decomposition/model/transformer.py:251-254
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return a * (x - mean) / (std + eps) + b


@pytest.mark.parametrize("T,C,with_res", [(300, 1024, True), (8, 1024, False), (1000, 512, True), (5000, 1024, False)])
def test_layernorm_wide_fwd_bwd(T, C, with_res):
    L = _lib()
    g = torch.Generator(device=DEV).manual_seed(T)
    x = (torch.randn(T, C, device=DEV, generator=g) * 2 + 0.5).to(BF)
    res = torch.randn(T, C, device=DEV, generator=g).to(BF) if with_res else None
    a = torch.rand(C, device=DEV, generator=g) + 0.5
    b = torch.randn(C, device=DEV, generator=g)
    y = torch.empty(T, C, dtype=BF, device=DEV)
    xs = torch.empty(T, C, dtype=BF, device=DEV) if with_res else None
    L.check(L.lib.focr_layernorm_wide_fwd(x.data_ptr(), L.ptr(res), a.data_ptr(), b.data_ptr(), L.ptr(xs), y.data_ptr(), T, C, 1e-6,
                                          L.cur_stream()), "ln_wide_fwd")
    _sync(L)
    xin = (x.float() + res.float()).to(BF) if with_res else x
    if with_res:
        assert torch.equal(xs, xin)
    xr = xin.float().clone().requires_grad_(True)
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = _ln_ref(xr, ar, br)
    assert _rel(y, ref) < 1e-2
    dy = torch.randn(T, C, device=DEV, generator=g).to(BF)
    ref.backward(dy.float())
    dx = torch.empty(T, C, dtype=BF, device=DEV)
    da, db = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    ws = _ws(L.lib.focr_layernorm_wide_workspace_bytes(C))
    L.check(L.lib.focr_layernorm_wide_bwd(dy.data_ptr(), xin.data_ptr(), a.data_ptr(), dx.data_ptr(), da.data_ptr(), db.data_ptr(), T,
                                          C, 1e-6, ws.data_ptr(), ws.numel(), L.cur_stream()), "ln_wide_bwd")
    _sync(L)
    assert _rel(dx, xr.grad) < 1e-2
    assert _rel(da, ar.grad) < 1e-4 and _rel(db, br.grad) < 1e-4


def test_text_embed_fwd_bwd():
    from oracle import dropout_rng as R
    L = _lib()
    vocab, E, B, T, rows_pad = 7, 512, 5, 9, 128
    g = torch.Generator(device=DEV).manual_seed(3)
    idx = torch.randint(0, vocab, (B, T), device=DEV, generator=g)
    lut = torch.randn(vocab, E, device=DEV, generator=g)
    out = torch.empty(rows_pad, 2 * E, dtype=BF, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    # reference: Embeddings * sqrt(E) | PositionalEncoding(zeros)  (transformer.py:168-186, :277-286, :346-348)
    pe = torch.zeros(T, E, device=DEV)
    pos = torch.arange(0, T, device=DEV).unsqueeze(1).float()
    div = torch.exp(torch.arange(0, E, 2, device=DEV).float() * -(math.log(10000.0) / E))
    pe[:, 0::2], pe[:, 1::2] = torch.sin(pos * div), torch.cos(pos * div)
    for p, seed, sid in ((0.0, 0, 0), (0.1, 11, 2)):
        L.check(L.lib.focr_text_embed_fwd(idx.data_ptr(), lut.data_ptr(), vocab, E, B, T, rows_pad, out.data_ptr(), p, seed, sid,
                                          status.data_ptr(), L.cur_stream()), "text_embed_fwd")
        _sync(L)
        assert int(status) == 0
        pe_b = pe.unsqueeze(0).expand(B, T, E)
        if p:
            keep = torch.from_numpy(R._keep(R.drop_key(seed, sid), np.arange(B * T * E, dtype=np.uint64), R.thresh16(p))
                                    .reshape(B, T, E)).to(DEV)
            pe_b = pe_b * keep * R.keep_scale(p)
        ref = torch.cat([lut[idx] * math.sqrt(E), pe_b], 2).reshape(B * T, 2 * E)
        assert (out[:B * T].float() - ref).abs().max().item() < 1e-2 * ref.abs().max().item()
        assert (out[B * T:] == 0).all()
    d_out = torch.randn(rows_pad, 2 * E, device=DEV, generator=g).to(BF)
    d_lut = torch.empty(vocab, E, device=DEV)
    L.check(L.lib.focr_text_embed_bwd(idx.data_ptr(), d_out.data_ptr(), vocab, E, B, T, d_lut.data_ptr(), L.cur_stream()))
    _sync(L)
    ref = torch.zeros(vocab, E, device=DEV).index_add_(0, idx.reshape(-1), d_out[:B * T, :E].float()) * math.sqrt(E)
    assert _rel(d_lut, ref) < 1e-5
    # a large table with many rows (image-ids-CTR: 4303 entries; more rows than one compaction chunk), repeated indices
    V2, B2, T2 = 4303, 96, 25
    idx2 = torch.randint(0, V2, (B2, T2), device=DEV, generator=g)
    idx2[:, 0] = 0
    idx2[::3, 5] = 17
    d2 = torch.randn(_pad := ((B2 * T2 + 127) // 128 * 128), 2 * E, device=DEV, generator=g).to(BF)
    d_lut2 = torch.empty(V2, E, device=DEV)
    L.check(L.lib.focr_text_embed_bwd(idx2.data_ptr(), d2.data_ptr(), V2, E, B2, T2, d_lut2.data_ptr(), L.cur_stream()))
    _sync(L)
    ref2 = torch.zeros(V2, E, device=DEV).index_add_(0, idx2.reshape(-1), d2[:B2 * T2, :E].float()) * math.sqrt(E)
    assert _rel(d_lut2, ref2) < 1e-5
    # an index outside the table is flagged, not dereferenced
    bad = idx.clone()
    bad[0, 0] = vocab + 3
    L.check(L.lib.focr_text_embed_fwd(bad.data_ptr(), lut.data_ptr(), vocab, E, B, T, rows_pad, out.data_ptr(), 0.0, 0, 0,
                                      status.data_ptr(), L.cur_stream()))
    _sync(L)
    assert int(status) == 2


def test_packed_ce_matches_the_reference_packing_loop():
    L = _lib()
    B, T, C, ld = 6, 11, 7, 64
    g = torch.Generator(device=DEV).manual_seed(8)
    logits = torch.randn(B * T, ld, device=DEV, generator=g) * 3
    length = torch.tensor([11, 1, 5, 0, 7, 11], device=DEV)
    gt = torch.randint(0, C, (int(length.sum()),), device=DEV, generator=g)
    loss = torch.empty(1, device=DEV)
    d = torch.full((B * T, ld), 7.0, dtype=BF, device=DEV)
    ws = _ws(L.lib.focr_packed_ce_workspace_bytes(B))
    L.check(L.lib.focr_packed_ce(logits.data_ptr(), ld, B, T, C, length.data_ptr(), gt.data_ptr(), 1.0, loss.data_ptr(), d.data_ptr(),
                                 ld, ws.data_ptr(), ws.numel(), L.cur_stream()), "packed_ce")
    _sync(L)
    x = logits.clone().requires_grad_(True)
    x3 = x.view(B, T, ld)[:, :, :C]
    packed = torch.cat([x3[b, :int(length[b])] for b in range(B)], 0)      # transformer.py:361-373
    ref = F.cross_entropy(packed, gt)                                       # train.py:41,71
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-6 * abs(float(ref)) + 1e-6
    assert _rel(d, x.grad) < 1e-2
    assert (d.float()[x.grad == 0] == 0).all()    # t >= length and the padding columns: exact zeros


def test_dropout_add_relu_maxpool():
    from oracle import dropout_rng as R
    L = _lib()
    g = torch.Generator(device=DEV).manual_seed(4)
    n, p, seed, sid = 4096 * 33, 0.1, 5, 9
    x = torch.randn(n, device=DEV, generator=g).to(BF)
    y = torch.empty_like(x)
    L.check(L.lib.focr_dropout(x.data_ptr(), y.data_ptr(), n, p, seed, sid, L.cur_stream()))
    _sync(L)
    keep = torch.from_numpy(R._keep(R.drop_key(seed, sid), np.arange(n, dtype=np.uint64), R.thresh16(p))).to(DEV)
    ref = (x.float() * keep * R.keep_scale(p)).to(BF)
    assert torch.equal(y, ref)
    a, b = torch.randn(8, 16, 16, 64, device=DEV, generator=g).to(BF), torch.randn(8, 16, 16, 64, device=DEV, generator=g).to(BF)
    out = torch.empty_like(a)
    L.check(L.lib.focr_add_relu(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), L.cur_stream()))
    _sync(L)
    assert torch.equal(out, F.relu(a.float() + b.float()).to(BF))
    dy = torch.randn_like(a)
    dx = torch.empty_like(a)
    L.check(L.lib.focr_relu_bwd(dy.data_ptr(), out.data_ptr(), dx.data_ptr(), a.numel(), L.cur_stream()))
    _sync(L)
    assert torch.equal(dx, torch.where(out > 0, dy, torch.zeros_like(dy)))
    # 2x2 max-pool on an NHWC map (distinct values per window so that the arg-max is unambiguous)
    Bm, H, W, C = 3, 32, 32, 64
    xm = torch.randperm(Bm * H * W * C, device=DEV, generator=g).float().view(Bm, H, W, C)
    rank = torch.rand(Bm, H // 2, W // 2, 4, device=DEV, generator=g).argsort(-1)          # which window slot holds the maximum
    off = rank.view(Bm, H // 2, W // 2, 2, 2).permute(0, 1, 3, 2, 4).reshape(Bm, H, W) * 256
    xm = (xm % 251 + off[..., None]).to(BF)
    ym = torch.empty(Bm, H // 2, W // 2, C, dtype=BF, device=DEV)
    L.check(L.lib.focr_maxpool2x2_fwd(xm.data_ptr(), ym.data_ptr(), Bm, H, W, C, L.cur_stream()))
    _sync(L)
    xr = xm.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
    yr = F.max_pool2d(xr, 2, 2)
    assert torch.equal(ym.float().permute(0, 3, 1, 2), yr)
    dym = torch.randn_like(ym)
    yr.backward(dym.float().permute(0, 3, 1, 2))
    dxm = torch.empty_like(xm)
    L.check(L.lib.focr_maxpool2x2_bwd(xm.data_ptr(), ym.data_ptr(), dym.data_ptr(), dxm.data_ptr(), Bm, H, W, C, L.cur_stream()))
    _sync(L)
    assert torch.equal(dxm.float().permute(0, 3, 1, 2), xr.grad)


@pytest.mark.parametrize("wd", [0.0, 1e-4])
def test_adadelta_matches_torch(wd):
    L = _lib()
    g = torch.Generator(device=DEV).manual_seed(2)
    shapes = [(64, 3, 3, 3), (1024,), (7, 512), (1000, 33)]
    params = [torch.randn(s, device=DEV, generator=g) for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.Adadelta(ref, lr=1.0, rho=0.9, eps=1e-6, weight_decay=wd)
    sq = [torch.zeros_like(p) for p in params]
    acc = [torch.zeros_like(p) for p in params]
    grads = [torch.empty_like(p) for p in params]
    table = torch.tensor([[p.data_ptr(), gr.data_ptr(), s.data_ptr(), a.data_ptr(), p.numel()]
                          for p, gr, s, a in zip(params, grads, sq, acc)], dtype=torch.int64, device=DEV)
    for step in range(3):
        for gr, r in zip(grads, ref):
            gr.copy_(torch.randn(gr.shape, device=DEV, generator=g) * 0.1)
            r.grad = gr.clone()
        opt.step()
        L.check(L.lib.focr_adadelta_step(table.data_ptr(), len(params), 1.0, 1.0, 0.9, 1e-6, wd, L.cur_stream()), "adadelta")
        _sync(L)
        for p, r in zip(params, ref):
            assert torch.allclose(p, r.detach(), rtol=1e-5, atol=1e-6), step


@pytest.mark.parametrize("B,H,W,Ci,Co,nchw", [(4, 32, 32, 3, 64, True), (8, 16, 16, 128, 256, False), (2, 16, 16, 512, 1024, False),
                                              (3, 16, 16, 64, 128, False)])
def test_conv3x3_gemm_fwd_and_wgrad(B, H, W, Ci, Co, nchw):
    L = _lib()
    g = torch.Generator(device=DEV).manual_seed(Ci)
    x = torch.randn(B, Ci, H, W, device=DEV, generator=g)
    w = torch.randn(Co, Ci, 3, 3, device=DEV, generator=g) / (9 * Ci) ** 0.5
    bias = torch.randn(Co, device=DEV, generator=g)
    xb = None if nchw else x.permute(0, 2, 3, 1).contiguous().to(BF)
    xq = x.to(BF).float() if nchw else xb.float().permute(0, 3, 1, 2)
    wr = w.to(BF).float().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    ref = F.conv2d(xq, wr, br, padding=1)
    ws = _ws(L.lib.focr_conv3x3_gemm_workspace_bytes(B, H, W, Ci, Co))
    if (B * H * W) % 128 == 0:
        y = torch.empty(B, H, W, Co, dtype=BF, device=DEV)
        L.check(L.lib.focr_conv3x3_gemm_fwd(L.ptr(xb), x.data_ptr() if nchw else 0, w.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H,
                                            W, Ci, Co, ws.data_ptr(), ws.numel(), L.cur_stream()), "conv3x3_gemm_fwd")
        _sync(L)
        assert _rel(y.permute(0, 3, 1, 2), ref) < 1e-2
    dy = torch.randn(B, H, W, Co, device=DEV, generator=g).to(BF)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    dw, db = torch.empty_like(w), torch.empty_like(bias)
    L.check(L.lib.focr_conv3x3_gemm_wgrad(dy.data_ptr(), L.ptr(xb), x.data_ptr() if nchw else 0, dw.data_ptr(), db.data_ptr(), B, H, W,
                                          Ci, Co, ws.data_ptr(), ws.numel(), L.cur_stream()), "conv3x3_gemm_wgrad")
    _sync(L)
    assert _rel(dw, wr.grad) < 1e-2
    assert _rel(db, br.grad) < 1e-3


@pytest.mark.parametrize("B,Ci,Co,flags", [(8, 128, 256, 0), (2, 512, 1024, 1), (4, 256, 256, 0), (3, 64, 128, 0)])
def test_implicit_conv_on_16x16_maps(B, Ci, Co, flags):
    """the encoder's maps after its one pooling step are 16 x 16 (transformer.py:130): W = 16 tiles of 8 image rows"""
    L = _lib()
    H = W = 16
    g = torch.Generator(device=DEV).manual_seed(Co)
    x = torch.randn(B, H, W, Ci, device=DEV, generator=g).to(BF)
    w = torch.randn(Co, Ci, 3, 3, device=DEV, generator=g) / (9 * Ci) ** 0.5
    bias = torch.randn(Co, device=DEV, generator=g)
    y = torch.empty(B, H, W, Co, dtype=BF, device=DEV)
    ws = _ws(L.lib.focr_conv2d_workspace_bytes(Ci, Co, 3))
    L.check(L.lib.focr_conv2d_fwd(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), 0, 0, B, H, W, Ci, Co, 3, flags,
                                  ws.data_ptr(), ws.numel(), L.cur_stream()), "conv2d_fwd W=16")
    _sync(L)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(BF).float(), bias, padding=1)
    if flags & 1:
        ref = F.relu(ref)
    assert _rel(y.permute(0, 3, 1, 2), ref) < 1e-2
    dy = torch.randn(B, H, W, Co, device=DEV, generator=g).to(BF)
    dx = torch.empty(B, H, W, Ci, dtype=BF, device=DEV)
    L.check(L.lib.focr_conv2d_dgrad(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), B, H, W, Ci, Co, 3, 0, ws.data_ptr(), ws.numel(),
                                    L.cur_stream()), "conv2d_dgrad W=16")
    _sync(L)
    ref_dx = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w.to(BF).float(), padding=1)
    assert _rel(dx.permute(0, 3, 1, 2), ref_dx) < 1e-2


@pytest.mark.parametrize("B,H,W,Ci,Co", [(8, 2, 16, 512, 1024), (4, 2, 16, 1024, 1024), (3, 4, 32, 256, 512), (2, 8, 64, 128, 256),
                                         (1, 16, 128, 64, 128), (3, 16, 160, 64, 128), (2, 16, 160, 256, 512), (1, 8, 96, 128, 128),
                                         (2, 4, 320, 64, 64)])
def test_implicit_conv_on_the_ids_encoder_maps(B, H, W, Ci, Co):
    """image-ids-CTR pools four times (model/transformer.py:126-152): 16x128, 8x64, 4x32 and 2x16 maps; a 2x16 map is a quarter of
    a 128-pixel tile, so the TMA box spans four images (zero halo still per image).  16x160 (the maps of 32x320 crops, BASELINE
    configs[3]) and other widths that are not a power of two are cut into 32- / 64-pixel-wide boxes of 4 / 2 rows."""
    L = _lib()
    g = torch.Generator(device=DEV).manual_seed(H * W + Co)
    x = torch.randn(B, H, W, Ci, device=DEV, generator=g).to(BF)
    w = torch.randn(Co, Ci, 3, 3, device=DEV, generator=g) / (9 * Ci) ** 0.5
    bias = torch.randn(Co, device=DEV, generator=g)
    y = torch.empty(B, H, W, Co, dtype=BF, device=DEV)
    ws = _ws(L.lib.focr_conv2d_workspace_bytes(Ci, Co, 3))
    L.check(L.lib.focr_conv2d_fwd(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), 0, 0, B, H, W, Ci, Co, 3, 0,
                                  ws.data_ptr(), ws.numel(), L.cur_stream()), "conv2d_fwd small map")
    _sync(L)
    wr = w.to(BF).float().requires_grad_(True)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wr, bias, padding=1)
    assert _rel(y.permute(0, 3, 1, 2), ref) < 1e-2
    dy = torch.randn(B, H, W, Co, device=DEV, generator=g).to(BF)
    dx = torch.empty(B, H, W, Ci, dtype=BF, device=DEV)
    L.check(L.lib.focr_conv2d_dgrad(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), B, H, W, Ci, Co, 3, 0, ws.data_ptr(), ws.numel(),
                                    L.cur_stream()), "conv2d_dgrad small map")
    _sync(L)
    assert _rel(dx.permute(0, 3, 1, 2), F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w.to(BF).float(), padding=1)) < 1e-2
    ref.backward(dy.float().permute(0, 3, 1, 2))
    dw, db = torch.empty_like(w), torch.empty_like(bias)
    ws2 = _ws(L.lib.focr_conv3x3_gemm_workspace_bytes(B, H, W, Ci, Co))
    L.check(L.lib.focr_conv3x3_gemm_wgrad(dy.data_ptr(), x.data_ptr(), 0, dw.data_ptr(), db.data_ptr(), B, H, W, Ci, Co, ws2.data_ptr(),
                                          ws2.numel(), L.cur_stream()), "conv3x3_gemm_wgrad small map")
    _sync(L)
    assert _rel(dw, wr.grad) < 1e-2
    # a batch that does not fill whole tiles is refused, not mis-tiled
    if H * W < 128:
        assert L.lib.focr_conv2d_fwd(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), 0, 0, B - 1, H, W, Ci, Co, 3, 0,
                                     ws.data_ptr(), ws.numel(), L.cur_stream()) != 0


def test_l2norm_rows_and_packed_feature_mse():
    """image-ids-CTR/train.py:66-80: pred / pred.norm(dim=1), and -MSE(pred_n, text_features[gt]) over the packed valid rows"""
    L = _lib()
    g = torch.Generator(device=DEV).manual_seed(12)
    B, T, C, V = 5, 7, 2048, 300
    rows_pad = 128
    x = torch.randn(rows_pad, C, device=DEV, generator=g) * 3
    x[40] = 0                                                  # an all-zero row must not produce NaN
    y = torch.empty(rows_pad, C, dtype=BF, device=DEV)
    inv = torch.empty(rows_pad, device=DEV)
    L.check(L.lib.focr_l2norm_rows_fwd(x.data_ptr(), C, y.data_ptr(), inv.data_ptr(), rows_pad, C, L.cur_stream()))
    _sync(L)
    xr = x.clone().requires_grad_(True)
    n = xr.norm(dim=1, keepdim=True)
    ref = xr / torch.where(n > 0, n, torch.ones_like(n))
    assert _rel(y, ref) < 1e-2 and torch.isfinite(y.float()).all() and (y[40] == 0).all()
    dy = torch.randn(rows_pad, C, device=DEV, generator=g).to(BF)
    # reference gradient through the bf16-rounded output the kernel stored
    yq = y.float()
    want = inv[:, None] * (dy.float() - yq * (dy.float() * yq).sum(1, keepdim=True))
    dx = torch.empty(rows_pad, C, dtype=BF, device=DEV)
    L.check(L.lib.focr_l2norm_rows_bwd(dy.data_ptr(), y.data_ptr(), inv.data_ptr(), dx.data_ptr(), C, rows_pad, C, L.cur_stream()))
    _sync(L)
    assert _rel(dx, want) < 1e-2
    ref.backward(dy.float())
    live = torch.arange(rows_pad, device=DEV) != 40     # the guarded zero row has no meaningful reference gradient
    assert _rel(dx[live], xr.grad[live]) < 2e-2 and (dx[40] == 0).all()
    # distance term
    length = torch.tensor([7, 2, 0, 5, 1], device=DEV)
    gt = torch.randint(0, V, (int(length.sum()),), device=DEV, generator=g)
    feats = torch.randn(V, C, device=DEV, generator=g) * 0.3
    loss = torch.empty(1, device=DEV)
    d = torch.full((rows_pad, C), 3.0, dtype=BF, device=DEV)
    ws = _ws(L.lib.focr_packed_ce_workspace_bytes(B))
    L.check(L.lib.focr_packed_feat_mse(y.data_ptr(), B, T, C, length.data_ptr(), gt.data_ptr(), feats.data_ptr(), V, 1.0, loss.data_ptr(),
                                       d.data_ptr(), ws.data_ptr(), ws.numel(), L.cur_stream()), "packed_feat_mse")
    _sync(L)
    yr = y.float().clone().requires_grad_(True)
    y3 = yr[:B * T].view(B, T, C)
    packed = torch.cat([y3[b, :int(length[b])] for b in range(B)], 0)
    ref_l = F.mse_loss(packed, feats[gt])
    ref_l.backward()
    assert abs(float(loss) - float(ref_l.detach())) < 1e-5 * float(ref_l.detach())
    assert _rel(d[:B * T], yr.grad[:B * T]) < 1e-2 and (d[:B * T].float()[yr.grad[:B * T] == 0] == 0).all()


def test_implicit_conv_160_wide_with_residual_tile():
    """the auxiliary (residual) tile of a 16x160 map travels through the same 32x4-pixel TMA boxes as the output"""
    L = _lib()
    B, H, W, Ci, Co = 2, 16, 160, 128, 128
    g = torch.Generator(device=DEV).manual_seed(160)
    x = torch.randn(B, H, W, Ci, device=DEV, generator=g).to(BF)
    res = torch.randn(B, H, W, Co, device=DEV, generator=g).to(BF)
    w = torch.randn(Co, Ci, 3, 3, device=DEV, generator=g) / (9 * Ci) ** 0.5
    bias = torch.randn(Co, device=DEV, generator=g)
    y = torch.empty(B, H, W, Co, dtype=BF, device=DEV)
    ws = _ws(L.lib.focr_conv2d_workspace_bytes(Ci, Co, 3))
    L.check(L.lib.focr_conv2d_fwd(x.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), 0, res.data_ptr(), B, H, W, Ci, Co, 3, 0,
                                  ws.data_ptr(), ws.numel(), L.cur_stream()), "conv2d_fwd W=160 + residual")
    _sync(L)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(BF).float(), bias, padding=1) + res.float().permute(0, 3, 1, 2)
    assert _rel(y.permute(0, 3, 1, 2), ref) < 1e-2


@pytest.mark.parametrize("B,H,W,Ci,Co", [(8, 16, 16, 128, 256), (2, 16, 16, 512, 1024), (4, 2, 16, 1024, 1024), (2, 8, 64, 128, 256),
                                         (1, 16, 128, 64, 128), (2, 16, 160, 128, 256)])
def test_conv_wgrad_on_tcgen05(B, H, W, Ci, Co):
    """dW = dY^T col with the pixel dimension as the GEMM's K (both operands transposed to K-major, fp32 accumulation in TMEM)"""
    L = _lib()
    g = torch.Generator(device=DEV).manual_seed(B * 7 + Co)
    x = torch.randn(B, H, W, Ci, device=DEV, generator=g).to(BF)
    dy = torch.randn(B, H, W, Co, device=DEV, generator=g).to(BF)
    w = torch.zeros(Co, Ci, 3, 3, device=DEV, requires_grad=True)
    F.conv2d(x.float().permute(0, 3, 1, 2), w, None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    dw, db = torch.empty(Co, Ci, 3, 3, device=DEV), torch.empty(Co, device=DEV)
    ws = _ws(L.lib.focr_conv3x3_wgrad_tc_workspace_bytes(B, H, W, Ci, Co))
    L.check(L.lib.focr_conv3x3_wgrad_tc(dy.data_ptr(), x.data_ptr(), 0, dw.data_ptr(), db.data_ptr(), B, H, W, Ci, Co, ws.data_ptr(),
                                        ws.numel(), L.cur_stream()), "conv3x3_wgrad_tc")
    _sync(L)
    assert _rel(dw, w.grad) < 1e-2
    assert _rel(db, dy.float().sum((0, 1, 2))) < 1e-3
    assert L.lib.focr_conv3x3_wgrad_tc(dy.data_ptr(), x.data_ptr(), 0, dw.data_ptr(), db.data_ptr(), B, H, W, Ci, 64, ws.data_ptr(),
                                       ws.numel(), L.cur_stream()) != 0        # 64 output channels: refused (streaming kernel's case)

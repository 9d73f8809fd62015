"""Op-level parity of the hand-written kernels (through the C-ABI) against torch fp32 ops evaluated on the
same bf16-rounded operands.  GPU only.  Tolerances are relative to the tensor's max magnitude:
1e-2 for bf16 outputs (north_star: 1e-2 bf16), tighter for fp32 reductions."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _L():
    from fudanocr_b200 import _lib as L
    return L


def _ws(n):
    return torch.empty(int(n), dtype=torch.uint8, device=DEV)


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-20)).item()


def _bf(x):
    return x.to(torch.bfloat16)


def _sync(L):
    L.check(L.lib.focr_sync_check(L.cur_stream()))


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,C,act", [(2048, 64, 1), (4096, 64, 0), (300, 256, 2), (7, 512, 2), (2048, 32, 2)])
def test_bn_train_fwd_bwd(T, C, act):
    L = _L()
    g = torch.Generator(device=DEV).manual_seed(T + C)
    x = _bf(torch.randn(T, C, device=DEV, generator=g) * 1.5 + 0.3)
    gamma = 1 + 0.1 * torch.randn(C, device=DEV, generator=g)
    beta = 0.1 * torch.randn(C, device=DEV, generator=g)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    nbt = torch.zeros((), dtype=torch.long, device=DEV)
    y = torch.empty_like(x)
    stats = torch.empty(4, C, device=DEV)
    ws = _ws(L.lib.focr_bn_workspace_bytes())
    L.check(L.lib.focr_bn_train_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                                    nbt.data_ptr(), y.data_ptr(), stats.data_ptr(), T, C, act, ws.data_ptr(),
                                    ws.numel(), L.cur_stream()))
    _sync(L)
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    yr = F.batch_norm(xr, rm2, rv2, gr, br, True, 0.1, 1e-5)
    if act == 1:
        yr = yr * torch.tanh(F.softplus(yr))
    elif act == 2:
        yr = F.relu(yr)
    assert _rel(y, yr) < 1e-2
    assert torch.allclose(rm, rm2, atol=1e-4, rtol=1e-3) and torch.allclose(rv, rv2, atol=1e-4, rtol=1e-3)
    assert int(nbt.item()) == 1
    dy = _bf(torch.randn(T, C, device=DEV, generator=g))
    yr.backward(dy.float())
    dx = torch.empty_like(x)
    dg, db = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    L.check(L.lib.focr_bn_bwd(dy.data_ptr(), x.data_ptr(), stats.data_ptr(), dx.data_ptr(), dg.data_ptr(),
                              db.data_ptr(), T, C, act, ws.data_ptr(), ws.numel(), L.cur_stream()))
    _sync(L)
    assert _rel(dx, xr.grad) < 2e-2
    assert _rel(dg, gr.grad) < 5e-3 and _rel(db, br.grad) < 5e-3


def _ln_ref(x, a, b, eps=1e-6):
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)
    return a * (x - mean) / (std + eps) + b


@pytest.mark.parametrize("T", [2048, 1000])
def test_layernorm_std(T):
    L = _L()
    x = _bf(torch.randn(T, 128, device=DEV) * 2 + 0.5)
    a = 1 + 0.1 * torch.randn(128, device=DEV)
    b = 0.1 * torch.randn(128, device=DEV)
    y = torch.empty_like(x)
    L.check(L.lib.focr_layernorm_std_fwd(x.data_ptr(), a.data_ptr(), b.data_ptr(), y.data_ptr(), T, 1e-6,
                                         L.cur_stream()))
    xr = x.float().requires_grad_(True)
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yr = _ln_ref(xr, ar, br)
    _sync(L)
    assert _rel(y, yr) < 1e-2
    dy = _bf(torch.randn(T, 128, device=DEV))
    yr.backward(dy.float())
    dx = torch.empty_like(x)
    da, db = torch.empty(128, device=DEV), torch.empty(128, device=DEV)
    ws = _ws(L.lib.focr_bn_workspace_bytes())
    L.check(L.lib.focr_layernorm_std_bwd(dy.data_ptr(), x.data_ptr(), a.data_ptr(), dx.data_ptr(), da.data_ptr(),
                                         db.data_ptr(), T, 1e-6, ws.data_ptr(), ws.numel(), L.cur_stream()))
    _sync(L)
    assert _rel(dx, xr.grad) < 1e-2
    assert _rel(da, ar.grad) < 5e-3 and _rel(db, br.grad) < 5e-3


# ---------------------------------------------------------------------------------------------
def _attn_ref(qkv, B, keep=None, scale_keep=1.0):
    q, k, v = [qkv[:, i * 128:(i + 1) * 128].view(B, 1024, 4, 32).transpose(1, 2) for i in range(3)]
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(32)
    p = F.softmax(s, dim=-1)
    if keep is not None:
        p = p * keep.to(p.dtype) * scale_keep
    return torch.matmul(p, v).transpose(1, 2).reshape(B * 1024, 128)


@pytest.mark.parametrize("B,p_drop,use_bits,exact,gain", [(2, 0.0, False, 0, 1.2), (1, 0.1, False, 0, 1.2),
                                                          (3, 0.0, True, 0, 1.2), (2, 0.1, True, 0, 1.2),
                                                          (2, 0.1, True, 1, 1.2), (1, 0.0, False, 0, 6.0),
                                                          (1, 0.1, True, 0, 2.5), (2, 0.1, True, 2, 1.2),
                                                          (1, 0.1, False, 2, 1.2)])
def test_attention_fwd_bwd(B, p_drop, use_bits, exact, gain):
    """use_bits: the forward stores its keep decisions (1 bit per element) and the backward reads them (the path the
    engine uses); otherwise the backward regenerates the mask from the seed.  Both must match the numpy twin.
    exact = 1 forces the forward's two-pass (row maximum) route, exact = 2 the two-kernel backward (the default is the
    single-pass kernel); gain = 6 makes the Cauchy-Schwarz score bound so loose
    (logits up to +-100) that the kernel must take that route by itself; gain = 2.5 is a peaked softmax on the bound route."""
    L = _L()
    from oracle import dropout_rng as R
    L.check(L.lib.focr_attn_set_force_exact(1 if exact == 1 else 0))
    L.check(L.lib.focr_attn_set_bwd_two_pass(1 if exact == 2 else 0))
    T = B * 1024
    g = torch.Generator(device=DEV).manual_seed(B)
    qkv = _bf(torch.randn(T, 384, device=DEV, generator=g) * gain)
    out = torch.empty(T, 128, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(B * 4 * 1024, device=DEV)
    seed, blk = 1234, 3
    nb = L.lib.focr_mha_drop_bits_bytes(B)
    assert nb == B * 4 * 1024 * 1024 // 8
    bits = torch.zeros(nb // 4, dtype=torch.int32, device=DEV) if use_bits else None
    bp = bits.data_ptr() if use_bits else None
    L.check(L.lib.focr_mha_flash_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, p_drop, seed, 2 * blk, bp,
                                     L.cur_stream()))
    _sync(L)
    L.check(L.lib.focr_attn_set_force_exact(0))
    keep = R.attn_keep_mask(B, seed, blk, p_drop).to(DEV) if p_drop > 0 else None
    if use_bits and p_drop > 0:
        # decode: word [b,h,k/32,q], key k of the group at bit (k%32)//4 + 8*(k&3)
        w = bits.view(B, 4, 32, 1024).long() & 0xFFFFFFFF             # [b,h,kw,q]
        kk = torch.arange(32, device=DEV)
        sh = kk // 4 + 8 * (kk & 3)
        m = (w[..., None] >> sh) & 1                                   # [b,h,kw,q,kk]
        m = m.permute(0, 1, 3, 2, 4).reshape(B, 4, 1024, 1024)
        assert torch.equal(m.bool(), keep.bool())
    qr = qkv.float().requires_grad_(True)
    ref = _attn_ref(qr, B, keep, R.attn_keep_scale(p_drop))
    assert _rel(out, ref) < 1e-2, _rel(out, ref)
    # log-sum-exp (log2 domain)
    q, k = [qkv.float()[:, i * 128:(i + 1) * 128].view(B, 1024, 4, 32).transpose(1, 2) for i in range(2)]
    lse_ref = torch.logsumexp(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(32), -1) / math.log(2.0)
    assert torch.allclose(lse.view(B, 4, 1024), lse_ref, atol=2e-2, rtol=1e-4)
    d_out = _bf(torch.randn(T, 128, device=DEV, generator=g))
    ref.backward(d_out.float())
    dqkv = torch.empty_like(qkv)
    bws = torch.empty(L.lib.focr_mha_bwd_workspace_bytes(B), dtype=torch.uint8, device=DEV)
    L.check(L.lib.focr_mha_flash_bwd(qkv.data_ptr(), out.data_ptr(), d_out.data_ptr(), lse.data_ptr(), bws.data_ptr(), bws.numel(),
                                     dqkv.data_ptr(), B, p_drop, seed, 2 * blk, bp, L.cur_stream()))
    _sync(L)
    L.check(L.lib.focr_attn_set_bwd_two_pass(0))
    for i, nm in enumerate("qkv"):
        e = _rel(dqkv[:, i * 128:(i + 1) * 128], qr.grad[:, i * 128:(i + 1) * 128])
        assert e < 2e-2, (nm, e)


def test_attention_shift_clamp_adversarial_first_keys():
    """forward on the bound route with the softmax shift pinned by the clamp: the first 32 keys of every row score
    -a^2/sqrt(32) and all others +a^2/sqrt(32) with a^2 c = 104 log2 units, so the first-keys estimate is 208 below the
    row maximum and shift = bound - 100; the row sum must neither overflow nor vanish."""
    L = _L()
    B, T = 1, 1024
    a = math.sqrt(104.0 / (1.4426950408889634 / math.sqrt(32)))
    u = torch.ones(32, device=DEV) * (a / math.sqrt(32))
    qkv = torch.zeros(T, 384, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(5)
    for h in range(4):
        qkv[:, h * 32:(h + 1) * 32] = u
        qkv[:, 128 + h * 32:128 + (h + 1) * 32] = u
        qkv[:32, 128 + h * 32:128 + (h + 1) * 32] = -u
    qkv[:, 256:] = torch.randn(T, 128, device=DEV, generator=g)
    qkv = _bf(qkv)
    out = torch.empty(T, 128, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(4 * 1024, device=DEV)
    L.check(L.lib.focr_mha_flash_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, 0.0, 0, 0, None, L.cur_stream()))
    _sync(L)
    ref = _attn_ref(qkv.float(), B)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    assert _rel(out, ref) < 1e-2
    q, k = [qkv.float()[:, i * 128:(i + 1) * 128].view(B, 1024, 4, 32).transpose(1, 2) for i in range(2)]
    lse_ref = torch.logsumexp(torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(32), -1) / math.log(2.0)
    assert torch.allclose(lse.view(B, 4, 1024), lse_ref, atol=2e-2, rtol=1e-4)


def test_attention_dropout_statistics():
    """in-kernel mask: drop rate = 3277/32768 (p = 0.1 to 2^-15) and E[out] unchanged (uniform V makes the dropped
    output = the keep fraction)"""
    from oracle import dropout_rng as R
    pq = R.attn_drop_rate(0.1)
    assert abs(pq - 0.1) < 1e-5
    L = _L()
    B = 2
    qkv = torch.zeros(B * 1024, 384, device=DEV)
    qkv[:, 256:] = 1.0  # V = 1, Q = K = 0 -> P uniform, out = (#kept/1024) / 0.9
    qkv = _bf(qkv)
    out = torch.empty(B * 1024, 128, dtype=torch.bfloat16, device=DEV)
    lse = torch.empty(B * 4 * 1024, device=DEV)
    L.check(L.lib.focr_mha_flash_fwd(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, 0.1, 77, 0, None, L.cur_stream()))
    _sync(L)
    o = out.float()
    assert abs(o.mean().item() - 1.0) < 2e-3          # unbiased
    sd = o[:, ::32].std().item()                       # per (row, head): binomial(1024, 1-pq)/1024/(1-pq)
    assert abs(sd - math.sqrt(pq * (1 - pq) / 1024) / (1 - pq)) < 2e-3


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,K,N", [(4096, 128, 384), (2048, 128, 128), (1000, 128, 64), (300, 512, 512),
                                   (5000, 384, 64)])
def test_linear_wgrad_and_bias(M, K, N):
    L = _L()
    dy = _bf(torch.randn(M, N, device=DEV))
    x = _bf(torch.randn(M, K, device=DEV))
    dw = torch.empty(N, K, device=DEV)
    db = torch.empty(N, device=DEV)
    ws = _ws(L.lib.focr_wgrad_workspace_bytes())
    L.check(L.lib.focr_linear_wgrad(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), M, K, N, ws.data_ptr(), ws.numel(),
                                    L.cur_stream()))
    L.check(L.lib.focr_bias_grad(dy.data_ptr(), db.data_ptr(), M, N, ws.data_ptr(), ws.numel(), L.cur_stream()))
    _sync(L)
    assert _rel(dw, dy.float().t() @ x.float()) < 2e-3
    assert _rel(db, dy.float().sum(0)) < 2e-3


@pytest.mark.parametrize("M,N", [(64, 64), (192, 128), (4096, 384), (64 * 301, 128), (64 * 1000, 384), (64 * 149, 64),
                                 (8192, 256)])
def test_linear_wgrad_bias_fused_tcgen05(M, N):
    """one pass over dY and X: tcgen05 MN-major operands + ones-operand bias gradient (csrc/wgrad_tc.cu); covers fewer tiles
    than CTAs, ragged tile/CTA splits, M = 64 accumulators (N = 64) and 1-3 M blocks"""
    L = _L()
    K = 128
    g = torch.Generator(device=DEV).manual_seed(M + N)
    dy = _bf(torch.randn(M, N, device=DEV, generator=g))
    x = _bf(torch.randn(M, K, device=DEV, generator=g) + 0.25)
    ws = _ws(L.lib.focr_wgrad_workspace_bytes())
    outs = []
    for rep in range(2):
        dw = torch.full((N, K), float("nan"), device=DEV)
        db = torch.full((N,), float("nan"), device=DEV)
        L.check(L.lib.focr_linear_wgrad_bias(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), db.data_ptr(), M, K, N, ws.data_ptr(),
                                             ws.numel(), L.cur_stream()))
        _sync(L)
        outs.append((dw, db))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])   # deterministic
    ref_w, ref_b = dy.double().t() @ x.double(), dy.double().sum(0)
    assert _rel(outs[0][0], ref_w) < 1e-4, _rel(outs[0][0], ref_w)
    assert _rel(outs[0][1], ref_b) < 1e-4, _rel(outs[0][1], ref_b)
    # weight-only call into a gradient tensor that is only 4-byte aligned (autograd path: views of one flat buffer)
    dw = torch.empty(N * K + 1, device=DEV)[1:].view(N, K)
    L.check(L.lib.focr_linear_wgrad_bias(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), 0, M, K, N, ws.data_ptr(), ws.numel(),
                                         L.cur_stream()))
    _sync(L)
    assert torch.equal(dw, outs[0][0])


@pytest.mark.parametrize("B,Co,shuf", [(2, 64, 0), (3, 128, 0), (2, 256, 1), (20, 64, 0), (37, 64, 0), (256, 64, 0), (33, 256, 1)])
def test_conv3x3_wgrad(B, Co, shuf):
    L = _L()
    H, W = 16, 64
    x = torch.randn(B, 64, H, W, device=DEV)
    xb = _bf(x.permute(0, 2, 3, 1).contiguous())
    if shuf:
        dy = torch.randn(B, 64, 2 * H, 2 * W, device=DEV)
        dyb = _bf(dy.permute(0, 2, 3, 1).contiguous())
        dy_conv = F.pixel_unshuffle(dyb.float().permute(0, 3, 1, 2), 2)
    else:
        dy = torch.randn(B, Co, H, W, device=DEV)
        dyb = _bf(dy.permute(0, 2, 3, 1).contiguous())
        dy_conv = dyb.float().permute(0, 3, 1, 2)
    w = torch.zeros(Co, 64, 3, 3, device=DEV, requires_grad=True)
    F.conv2d(xb.float().permute(0, 3, 1, 2), w, padding=1).backward(dy_conv)
    dw = torch.empty(Co, 64, 3, 3, device=DEV)
    ws = _ws(L.lib.focr_wgrad_workspace_bytes())
    L.check(L.lib.focr_conv2d_wgrad(dyb.data_ptr(), xb.data_ptr(), dw.data_ptr(), B, H, Co, 2 * shuf, ws.data_ptr(),
                                    ws.numel(), L.cur_stream()))
    _sync(L)
    assert _rel(dw, w.grad) < 2e-3


def test_mse_and_adam_clip():
    L = _L()
    torch.manual_seed(0)
    sr, hr = torch.rand(2, 3, 32, 128, device=DEV), torch.rand(2, 3, 32, 128, device=DEV)
    d = torch.empty_like(sr)
    loss = torch.empty(1, device=DEV)
    ws = _ws(1 << 20)
    L.check(L.lib.focr_mse_loss_grad(sr.data_ptr(), hr.data_ptr(), d.data_ptr(), loss.data_ptr(), sr.numel(), 100.0,
                                     ws.data_ptr(), ws.numel(), L.cur_stream()))
    _sync(L)
    srr = sr.clone().requires_grad_(True)
    lr_ = F.mse_loss(srr, hr)
    (lr_ * 100).backward()
    assert abs(loss.item() - lr_.item()) < 1e-6 and torch.allclose(d, srr.grad, rtol=1e-4, atol=1e-9)
    # fused clip + Adam vs torch, 3 steps, tensors of awkward sizes
    shapes = [(70000,), (64, 64, 3, 3), (1,), (513,)]
    ps = [torch.randn(s, device=DEV) for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in ps]
    opt = torch.optim.Adam(ref, lr=1e-4, betas=(0.5, 0.999))
    gs = [torch.empty_like(p) for p in ps]
    ms = [torch.zeros_like(p) for p in ps]
    vs = [torch.zeros_like(p) for p in ps]
    recs = []
    for p, g, m, v in zip(ps, gs, ms, vs):
        n, off = p.numel(), 0
        while off < n:
            ln = min(65536, n - off)
            recs.append([p.data_ptr() + 4 * off, g.data_ptr() + 4 * off, m.data_ptr() + 4 * off, v.data_ptr() + 4 * off, ln])
            off += ln
    table = torch.tensor(recs, dtype=torch.int64, device=DEV)
    step = torch.zeros((), dtype=torch.int64, device=DEV)
    state = torch.zeros(4, device=DEV)
    for it in range(3):
        for g, r in zip(gs, ref):
            g.copy_(torch.randn_like(g) * (3.0 if it == 0 else 0.01))
            r.grad = g.clone()
        gn = torch.nn.utils.clip_grad_norm_(ref, 0.25)
        opt.step()
        L.check(L.lib.focr_adam_clip_step(table.data_ptr(), table.shape[0], 1.0, 0.25, 1e-4, 0.5, 0.999, 1e-8,
                                          step.data_ptr(), state.data_ptr(), ws.data_ptr(), ws.numel(), L.cur_stream()))
        _sync(L)
        assert abs(state[0].item() - gn.item()) < 1e-4 * gn.item()
        for p, r in zip(ps, ref):
            assert torch.allclose(p, r.detach(), rtol=1e-5, atol=1e-7)
    assert int(step.item()) == 3

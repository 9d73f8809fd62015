"""tcgen05 operand-descriptor semantics the attention kernels (csrc/attention.cu) rely on, pinned with a probe kernel:
the test builds the exact shared-memory image (TMA-style 64 B / 128 B swizzles) in numpy, hands the probe the two
64-bit operand descriptors, and compares the dumped TMEM accumulator with A @ B^T computed in fp32."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _L():
    import fudanocr_b200._lib as L
    return L


def _bf16_bits(x: torch.Tensor) -> np.ndarray:
    return x.to(torch.bfloat16).view(torch.int16).cpu().numpy().view(np.uint16)


def _swz(off: int, mode: int) -> int:
    if mode == 128:
        return off ^ (((off >> 7) & 7) << 4)
    if mode == 64:
        return off ^ (((off >> 7) & 3) << 4)
    return off


def _place(img: np.ndarray, base: int, bits: np.ndarray, mode: int):
    """row-major [rows][cols] bf16 tile at byte offset `base` (1024-aligned), rows of cols*2 bytes, swizzled the way a
    TMA box load with CU_TENSOR_MAP_SWIZZLE_<mode>B writes it"""
    rows, cols = bits.shape
    rb = cols * 2
    v = img.view(np.uint16)
    for r in range(rows):
        for c in range(rb // 16):
            dst = base + _swz(r * rb + c * 16, mode)
            v[dst // 2: dst // 2 + 8] = bits[r, c * 8: c * 8 + 8]


def _desc(off: int, lbo: int, sbo: int, layout: int) -> int:
    return ((off >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) | (layout << 61)


def _idesc(M, N, a_mn=0, b_mn=0):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def _run(img, da, db, idesc, nk, sa, sb, ncols):
    L = _L()
    import ctypes as C
    fn = L.lib.focr_umma_probe
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong, C.c_ulonglong, C.c_uint, C.c_int, C.c_uint, C.c_uint, C.c_void_p,
                   C.c_int, C.c_void_p]
    d_img = torch.from_numpy(img.copy()).to(DEV)
    out = torch.zeros(128, ncols, device=DEV)
    L.check(fn(d_img.data_ptr(), img.size, da, db, idesc, nk, sa, sb, out.data_ptr(), ncols, L.cur_stream()))
    torch.cuda.synchronize()
    return out.cpu()


def _rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def _mats(m, n, k, seed):
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(m, k, generator=g).to(torch.bfloat16)
    b = torch.randn(n, k, generator=g).to(torch.bfloat16)
    return a, b


def test_k_major_sw64_qk():
    """S = Q K^T: A = Q tile [128][32], B = K tile [64][32], both K-major with the 64-byte swizzle; two K steps"""
    a, b = _mats(128, 64, 32, 1)
    img = np.zeros(16384, np.uint8)
    _place(img, 0, _bf16_bits(a), 64)
    _place(img, 8192, _bf16_bits(b), 64)
    out = _run(img, _desc(0, 16, 512, 4), _desc(8192, 16, 512, 4), _idesc(128, 64), 2, 2, 2, 64)
    ref = a.float() @ b.float().t()
    assert _rel(out[:, :64], ref) < 1e-5


def test_mn_major_b_sw64_pv():
    """O = P V: A = P [128 q][64 keys] K-major / 128 B swizzle, B = V [64 keys][32 d] as stored (MN-major, 64 B swizzle)"""
    g = torch.Generator().manual_seed(2)
    p = torch.rand(128, 64, generator=g).to(torch.bfloat16)
    v = torch.randn(64, 32, generator=g).to(torch.bfloat16)
    img = np.zeros(16384 + 4096, np.uint8)
    _place(img, 0, _bf16_bits(p), 128)
    _place(img, 16384, _bf16_bits(v), 64)
    ref = p.float() @ v.float()
    errs = {}
    for lbo in (16, 512, 64):
        out = _run(img, _desc(0, 16, 1024, 2), _desc(16384, lbo, 512, 4), _idesc(128, 32, 0, 1), 4, 2, 64, 32)
        errs[lbo] = _rel(out[:, :32], ref)
    print("PV errs by LBO", errs)
    assert errs[16] < 1e-5, errs


def test_mn_major_a_m64_dv():
    """dV = P^T dO with a 64-key accumulator: A = the [128 q][64 keys] 128B-swizzled P buffer read MN-major (M = keys),
    B = dO [128 q][32 d] MN-major / 64 B swizzle, 8 K steps of 16 queries; M = 64 rows land on lanes (r%16) + 32 (r/16)"""
    g = torch.Generator().manual_seed(3)
    p = torch.rand(128, 64, generator=g).to(torch.bfloat16)
    do = torch.randn(128, 32, generator=g).to(torch.bfloat16)
    img = np.zeros(16384 + 8192, np.uint8)
    _place(img, 0, _bf16_bits(p), 128)
    _place(img, 16384, _bf16_bits(do), 64)
    ref = p.float().t() @ do.float()     # [64 keys][32]
    rows = torch.tensor([(r % 16) + 32 * (r // 16) for r in range(64)])
    errs = {}
    for lbo in (16, 1024, 128):
        out = _run(img, _desc(0, lbo, 1024, 2), _desc(16384, 16, 512, 4), _idesc(64, 32, 1, 1), 8, 128, 64, 32)
        errs[lbo] = _rel(out[rows, :32], ref)
    print("dV errs by LBO", errs)
    assert errs[16] < 1e-5, errs


def test_mn_major_a_m128_two_atoms():
    """A = [128 q][128 keys] held as two 128B-swizzled [128][64] blocks 16 KB apart, read MN-major with M = 128 keys
    (two swizzle atoms along M: LBO = 16384), B = [128 q][32 d] MN-major"""
    g = torch.Generator().manual_seed(4)
    p = torch.rand(128, 128, generator=g).to(torch.bfloat16)
    do = torch.randn(128, 32, generator=g).to(torch.bfloat16)
    img = np.zeros(32768 + 8192, np.uint8)
    bits = _bf16_bits(p)
    _place(img, 0, bits[:, :64].copy(), 128)
    _place(img, 16384, bits[:, 64:].copy(), 128)
    _place(img, 32768, _bf16_bits(do), 64)
    ref = p.float().t() @ do.float()     # [128 keys][32]
    out = _run(img, _desc(0, 16384, 1024, 2), _desc(32768, 16, 512, 4), _idesc(128, 32, 1, 1), 8, 128, 64, 32)
    err = _rel(out[:, :32], ref)
    print("M=128 two-atom MN-major A err", err)
    assert err < 1e-5, err


def test_mn_major_both_two_atoms_wgrad():
    """the linear weight-gradient configuration (csrc/wgrad_tc.cu): A = dY tile [64 t][128 n] and B = X tile [64 t][128 k], each
    held as two [64][64] 128B-swizzled atoms 8 KB apart and read MN-major (M = N = 128, LBO = atom stride), 4 K steps of 16
    tokens; plus the bias trick: B = a region of bf16 ones read as an MN-major N = 16 operand"""
    g = torch.Generator().manual_seed(5)
    dy = torch.randn(64, 128, generator=g).to(torch.bfloat16)
    x = torch.randn(64, 128, generator=g).to(torch.bfloat16)
    img = np.zeros(16384 * 2 + 8192, np.uint8)
    by, bx = _bf16_bits(dy), _bf16_bits(x)
    _place(img, 0, by[:, :64].copy(), 128)
    _place(img, 8192, by[:, 64:].copy(), 128)
    _place(img, 16384, bx[:, :64].copy(), 128)
    _place(img, 24576, bx[:, 64:].copy(), 128)
    img[32768:].view(np.uint16)[:] = 0x3F80
    ref = dy.float().t() @ x.float()
    out = _run(img, _desc(0, 8192, 1024, 2), _desc(16384, 8192, 1024, 2), _idesc(128, 128, 1, 1), 4, 128, 128, 128)
    err = _rel(out, ref)
    print("wgrad M=N=128 two-atom MN-major err", err)
    assert err < 1e-5, err
    out = _run(img, _desc(0, 8192, 1024, 2), _desc(32768, 8192, 1024, 2), _idesc(128, 16, 1, 1), 4, 128, 0, 32)
    errb = _rel(out[:, 0], dy.float().sum(0))
    assert errb < 1e-5, errb
    # N = 64 output features: a single atom, M = 64 (rows on lanes (r % 16) + 32 (r / 16))
    rows = torch.tensor([(r % 16) + 32 * (r // 16) for r in range(64)])
    out = _run(img, _desc(0, 16, 1024, 2), _desc(16384, 8192, 1024, 2), _idesc(64, 128, 1, 1), 4, 128, 128, 128)
    err64 = _rel(out[rows], ref[:64])
    assert err64 < 1e-5, err64

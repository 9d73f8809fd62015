"""CRNN evaluator + greedy CTC decode parity (GPU).  Logits: bf16 tensor-core compute vs the fp32 oracle /
reference golden, relative L2 <= 2e-2; decode: INT32 index arrays bit-exact against the reference logic."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def env():
    from oracle import synth, crnn_oracle as C
    from fudanocr_b200.model.crnn import CRNN
    from fudanocr_b200.interfaces import recognition as R
    sd = synth.synth_state_dict(synth.load_spec("crnn"), seed=4321)
    golden = torch.load(synth.GOLDEN_DIR / "crnn_b2.pt", weights_only=False)
    m = CRNN(32, 1, 37, 256).to(DEV).eval()
    m.load_state_dict(sd)
    return dict(C=C, R=R, sd=sd, golden=golden, model=m, synth=synth)


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def test_parse_crnn_data_matches_torch_bicubic(env):
    g = env["golden"]
    out = env["R"].parse_crnn_data(g["sr"].to(DEV))
    assert out.shape == (2, 1, 32, 100)
    assert torch.allclose(out.cpu(), g["gray"], atol=2e-5, rtol=1e-5)


def test_crnn_logits_vs_golden(env):
    g = env["golden"]
    m = env["model"]
    with torch.no_grad():
        a = m(g["gray"].to(DEV))          # CRNN.forward proper
        b = m.forward_rgb(g["sr"].to(DEV))  # parse_crnn_data fused in
    assert a.shape == (26, 2, 37)
    assert _rel(a.cpu(), g["logits"]) < 2e-2, _rel(a.cpu(), g["logits"])
    assert _rel(b.cpu(), g["logits"]) < 2e-2


def test_decode_bit_exact(env):
    C, R, g = env["C"], env["R"], env["golden"]
    # (1) on the golden logits: path, collapsed labels and strings equal the reference's
    path, out, ln = R.ctc_greedy_decode(g["logits"].to(DEV))
    assert torch.equal(path.cpu().long(), g["path"])
    assert R.get_crnn_pred(g["logits"].permute(1, 0, 2).to(DEV)) == g["strings"]
    # (2) adversarial paths: long repeats, blank-separated repeats, all blank, exact ties, B not a multiple of 32
    torch.manual_seed(0)
    T, B, K = 26, 77, 37
    idx = torch.randint(0, 6, (T, B))                       # few classes -> many repeats
    idx[:, 0] = 0                                           # all blank
    idx[:, 1] = 5                                           # one long run
    idx[::2, 2], idx[1::2, 2] = 7, 0                        # x _ x _ ... : blanks separate repeats
    logits = torch.randn(T, B, K) * 0.1
    logits.scatter_(2, idx.unsqueeze(2), 5.0)
    logits[:, 3, :] = 1.0                                   # exact ties everywhere -> index 0 wins (torch.max)
    path, out, ln = R.ctc_greedy_decode(logits.to(DEV))
    ref_path = C.greedy_path(logits)
    assert torch.equal(path.cpu().long(), ref_path)
    for b in range(B):
        want = C.ctc_collapse(ref_path[b].tolist())
        assert out[b, :ln[b]].cpu().tolist() == want and int(ln[b]) == len(want)
        assert (out[b, ln[b]:] == -1).all()
    assert R.get_crnn_pred(logits.permute(1, 0, 2).to(DEV)) == C.get_crnn_pred(logits.permute(1, 0, 2))


def test_eval_pipeline_batch_independence(env):
    """full size (B = 128 per GPU, BASELINE configs[4] shard): every crop is recognised independently"""
    m = env["model"]
    _, hr = env["synth"].synth_images(8, seed=5)
    x = (hr * 2 - 1).to(DEV).repeat(16, 1, 1, 1)
    with torch.no_grad():
        full = m.forward_rgb(x)
        part = m.forward_rgb(x[:8])
    assert torch.equal(full[:, :8], part) and torch.equal(full[:, 8:16], part)
    from oracle import crnn_oracle as C
    sd = {k: v.to(DEV) for k, v in env["sd"].items()}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        ref = C.crnn_forward(sd, C.parse_crnn_data(x[:8]))
    assert _rel(part, ref) < 2e-2

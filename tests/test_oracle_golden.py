"""CPU: the oracle restatement vs golden vectors produced by the real reference modules
(oracle/make_golden.py ran the unmodified /root/reference code; see tests/golden/SHA256SUMS)."""
import hashlib

import pytest
import torch

from oracle import synth, tbsrn_oracle as O


@pytest.fixture(scope="module")
def golden():
    path = synth.GOLDEN_DIR / "tbsrn_b4.pt"
    want = (synth.GOLDEN_DIR / "SHA256SUMS").read_text().split()[0]
    assert hashlib.sha256(path.read_bytes()).hexdigest() == want
    return torch.load(path, weights_only=False)


@pytest.fixture(scope="module")
def sd():
    return synth.synth_state_dict(synth.load_spec("tbsrn"), 1234, O.tps_buffers())


def test_spec_matches_reference_layout():
    spec = synth.load_spec("tbsrn")
    assert len(spec) == 346  # 346 state-dict entries
    assert spec["block2.feature_enhancer.multihead.linears.0.weight"] == [128, 128]
    assert spec["block2.gru1.gru.weight_ih_l0"] == [96, 64]  # dead but present
    assert spec["block8.0.conv.weight"] == [256, 64, 3, 3]
    assert spec["tps.inverse_kernel"] == [23, 23]


def test_eval_forward(golden, sd):
    lr, _ = synth.synth_images(4)
    with torch.no_grad():
        sr = O.tbsrn_forward(sd, lr, training=False)
    assert torch.allclose(sr, golden["eval_sr"], atol=2e-5, rtol=1e-4)


def test_train_step(golden, sd):
    lr, hr = synth.synth_images(4)
    new_sd, info = O.train_step(sd, lr, hr, {}, masks=None)
    assert torch.allclose(info["sr"], golden["train_sr"], atol=2e-5, rtol=1e-4)
    assert abs(info["mse"].item() - golden["train_mse"].item()) < 1e-6
    assert abs(info["grad_norm"].item() - golden["grad_norm"].item()) < 1e-4 * golden["grad_norm"].item()
    for k, g in golden["grads"].items():
        assert torch.allclose(info["grads"][k], g, atol=1e-5 + 1e-4 * g.abs().max().item(), rtol=1e-3), k
    for k, n in golden["grad_l2"].items():
        assert abs(info["grads"][k].norm().item() - n) <= 1e-3 * n + 1e-7, k
    params = {k for k in sd if sd[k].is_floating_point() and not O.is_buffer(k)}
    assert sorted(params - set(info["grads"])) == golden["no_grad_params"]
    assert len(golden["no_grad_params"]) == 114  # 114 tensors never receive a gradient
    for k, (s, a) in golden["new_param_sum"].items():
        assert abs(new_sd[k].double().sum().item() - s) <= 1e-5 * a + 1e-6, k
    for k, v in golden["new_running"].items():
        assert torch.allclose(new_sd[k], v, atol=1e-6, rtol=1e-5), k


def test_dropout_mask_changes_output(sd):
    lr, _ = synth.synth_images(4)
    g = torch.Generator().manual_seed(0)
    masks = {}
    for i in range(2, 7):
        masks[f"block{i}.feature_enhancer.attn"] = torch.rand(4, 4, 1024, 1024, generator=g) >= 0.1
        masks[f"block{i}.feature_enhancer.ffn"] = torch.rand(4, 1024, 128, generator=g) >= 0.1
    with torch.no_grad():
        a = O.tbsrn_forward(sd, lr, training=True, masks=masks)
        b = O.tbsrn_forward(sd, lr, training=True, masks=None)
    assert (a - b).abs().max() > 1e-4

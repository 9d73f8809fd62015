"""Assembly of the stroke-level-decomposition recogniser (fudanocr_b200/model/transformer.py) checked on the CPU: the kernel
wrappers are replaced by torch-fp32 stand-ins (tests/_recog_mock.py, test infrastructure) and the composed model - autograd
wiring, NHWC layouts, parameter mapping, padding rows, packing - must reproduce the golden values recorded from the unmodified
reference module (tests/golden/sld_b3.pt) to fp32 accuracy.  The kernels themselves are checked on the B200."""
import torch

from oracle import sld_oracle as SO, synth


def _setup(monkeypatch):
    import _recog_mock
    _recog_mock.install(monkeypatch)
    from fudanocr_b200.model.transformer import Transformer
    g = torch.load(synth.GOLDEN_DIR / "sld_b3.pt", weights_only=False)
    model = Transformer("stroke")
    sd = synth.synth_state_dict(synth.load_spec("sld"), 1234)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert missing == ["pe.pe"] and not unexpected
    image, strings = SO.synth_batch(g["B"])
    assert strings == g["strings"] and abs(float(image.double().sum()) - g["image_checksum"]) < 1e-6
    return model, g, image


def test_state_dict_keys_match_the_reference():
    from fudanocr_b200.model.transformer import Transformer
    spec = synth.load_spec("sld")
    keys = [k for k in Transformer("stroke").state_dict().keys() if k != "pe.pe"]
    assert keys == list(spec.keys())
    sd = Transformer("stroke").state_dict()
    assert all(list(sd[k].shape) == v for k, v in spec.items()) and tuple(sd["pe.pe"].shape) == (1, 7000, 512)


def test_assembled_train_step_matches_reference_golden(monkeypatch):
    model, g, image = _setup(monkeypatch)
    model.train()
    model.dropout_p = 0.0
    out = model(image, g["length"], g["text_input"])
    assert torch.allclose(out["pred"], g["pred"], rtol=1e-3, atol=1e-4)
    assert torch.allclose(out["map"], g["map"], rtol=1e-3, atol=1e-5)
    assert out["conv"].shape == (g["B"], 1024, 16, 16)
    assert torch.allclose(out["conv"][:, ::16, ::2, ::2], g["conv_sample"], rtol=1e-2, atol=1e-3)   # 40 fp32 convs, other order
    loss = torch.nn.CrossEntropyLoss()(out["pred"], g["text_gt"])           # the reference loop, train.py:68-71
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    model.zero_grad()
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters()}
    for k, n in g["grad_norms"].items():
        if n is None:
            assert grads[k] is None, k
        else:
            assert abs(float(grads[k].norm()) - float(n)) < 1e-2 * float(n) + 1e-7, k   # 40 layers deep: fp32 re-ordering shows at 2e-3
    # The 40-layer train-mode-BN encoder is ill-conditioned: the SAME oracle evaluated in fp32 and in fp64 differs by 1.2e-2
    # (relative L2) in its encoder gradients and by 1e-6 in the decoder's (measured), so re-ordered fp32 arithmetic is held
    # to 5e-2 there and to 2e-3 in the decoder, where a wiring mistake would show as O(1).
    def rel(a, b):
        return float((a - b).norm() / (b.norm() + 1e-30))
    for k, v in g["grads_small"].items():
        if float(v.abs().max()) < 1e-6:     # conv biases in front of a train-mode BatchNorm: the true gradient is 0 (noise 1e-9)
            continue
        assert rel(grads[k], v) < (5e-2 if k.startswith("encoder") else 2e-3), (k, rel(grads[k], v))
    for k, v in g["grad_samples"].items():
        s = grads[k].reshape(-1)[::max(grads[k].numel() // 4096, 1)][:4096]
        assert rel(s, v) < (5e-2 if k.startswith("encoder") else 2e-3), (k, rel(s, v))
    for k, v in g["running_after"].items():
        assert torch.allclose(model.state_dict()[k], v, rtol=1e-4, atol=1e-6), k


def test_fused_loss_equals_reference_loop_and_eval_contract(monkeypatch):
    model, g, image = _setup(monkeypatch)
    model.train()
    model.dropout_p = 0.0
    loss = model.loss(image, g["length"], g["text_input"], g["text_gt"])
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    loss.backward()
    gn = {k: p.grad.norm() for k, p in model.named_parameters() if p.grad is not None}
    for k, n in g["grad_norms"].items():
        if n is not None:
            assert abs(float(gn[k]) - float(n)) < 1e-2 * float(n) + 1e-7, k
    # eval: running statistics (now updated twice by the two train forwards? no - one forward in this test), test=True contract
    model.eval()
    with torch.no_grad():
        ev = model(image, g["length"], g["text_input"], test=True)
        assert ev["pred"].shape == (g["B"], g["text_input"].shape[1], 7) and ev["map"].shape[:3] == (g["B"], 4, g["text_input"].shape[1])
        assert torch.allclose(ev["pred"], g["eval_pred"], rtol=1e-3, atol=1e-4)
        assert torch.allclose(ev["map"], g["eval_map"], rtol=1e-3, atol=1e-5)
        # cached encoder features are accepted back, as the test-time decode loop does (train.py:114-121)
        ev2 = model(None, g["length"], g["text_input"], conv_feature=ev["conv"], test=True)
        assert torch.equal(ev2["pred"], ev["pred"])
        assert set(model(image, None, None).keys()) == {"conv"}

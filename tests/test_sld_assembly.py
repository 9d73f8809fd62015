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


def test_assembled_model_with_dropout_on_matches_oracle_with_the_rng_twin_masks(monkeypatch):
    """dropout ON: the four dropout sites of the recogniser (positional half of the embedding, masked self-attention map,
    cross-attention map, FFN hidden) draw their keep-masks from the counter hash of csrc/common.cuh with element indices
    (b*T+t)*512+c / ((b*4+h)*Tq+i)*Tk+j / row*2048+c and streams 0..3; the oracle fed the numpy twin's masks
    (oracle/dropout_rng.py) must agree with the assembled model - the kernels themselves are checked against the same twin on
    the B200 (tests/test_gpu_recog_ops.py)"""
    import numpy as np
    from oracle import dropout_rng as R
    model, g, image = _setup(monkeypatch)
    model.train()
    p = model.dropout_p
    assert p == 0.1
    B, T = g["text_input"].shape
    seed = (model._seed * 1103515245 + 12345) & 0x7FFFFFFF          # the value decode_hidden() will draw next
    rows_pad = (B * T + 127) // 128 * 128

    def keep(n, sid):
        return torch.from_numpy(R._keep(R.drop_key(seed, sid), np.arange(n, dtype=np.uint64), R.thresh16(p))).float()
    drop = {"scale": R.keep_scale(p),
            "pe": keep(B * T * 512, 0).view(B, T, 512),
            "self": keep(B * 4 * T * T, 1).view(B, 4, T, T),
            "cross": keep(B * 4 * T * 256, 2).view(B, 4, T, 256),
            "ffn": keep(rows_pad * 2048, 3).view(rows_pad, 2048)[:B * T].view(B, T, 2048)}
    sd = synth.synth_state_dict(synth.load_spec("sld"), 1234)
    osd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    o_loss, o_logits, o_map, _ = SO.loss_fn(osd, image, g["length"], g["text_input"], g["text_gt"], drop)
    o_loss.backward()
    loss = model.loss(image, g["length"], g["text_input"], g["text_gt"])
    assert model._seed == seed
    assert abs(float(loss) - float(o_loss)) < 1e-4 * float(o_loss)
    assert abs(float(o_loss) - float(g["loss"])) > 1e-4              # the masks did change the result
    loss.backward()
    for k in ("generator_word.proj.weight", "decoder.pff.w_1.weight", "decoder.multihead.linears.1.weight",
              "decoder.mask_multihead.linears.0.weight", "embedding_word.lut.weight", "decoder.mul_layernorm2.a"):
        a, b = dict(model.named_parameters())[k].grad, osd[k].grad
        assert float((a - b).norm() / b.norm()) < 2e-3, k


def test_greedy_decode_follows_the_reference_test_loop(monkeypatch):
    """util_recog.greedy_decode_sld on the (mock-backed) drop-in module vs the loop of train.py:110-137 written out on the oracle"""
    from fudanocr_b200.util_recog import greedy_decode_sld
    model, g, image = _setup(monkeypatch)
    model.eval()
    sd = synth.synth_state_dict(synth.load_spec("sld"), 1234)
    max_length = 5
    pred, prob, seqs, overall = greedy_decode_sld(model, image, max_length)
    # reference loop on the oracle
    B = image.shape[0]
    with torch.no_grad():
        o_pred = torch.zeros(B, 1, dtype=torch.long)
        o_prob = torch.zeros(B, max_length)
        feats = None
        for i in range(max_length):
            prediction, _, feats = SO.forward(sd, image, o_pred, train=False, conv_feature=feats)
            now_pred = torch.max(torch.softmax(prediction, 2), 2)[1]
            o_prob[:, i] = torch.max(torch.softmax(prediction, 2), 2)[0][:, -1]
            o_pred = torch.cat((o_pred, now_pred[:, -1].view(-1, 1)), 1)
    assert torch.equal(pred, o_pred) and torch.allclose(prob, o_prob, rtol=1e-4, atol=1e-6)
    for b in range(B):
        now = []
        for j in range(max_length):
            now.append(int(o_pred[b][j]))
            if int(o_pred[b][j]) == 6:
                break
        assert seqs[b] == now[1:]
        op = 1.0
        for j in range(len(now) - 1):
            op *= float(o_prob[b][j])
        assert abs(overall[b] - op) < 1e-6


def test_character_mode_assembles_with_a_wide_generator(monkeypatch):
    """mode 'character' (config.py:5): the generator / embedding span the caller's alphabet (3757 symbols in the reference); checked
    here with a 300-symbol alphabet against the oracle on the same weights (loss and the generator / embedding gradients)"""
    import _recog_mock
    _recog_mock.install(monkeypatch)
    from fudanocr_b200.model.transformer import Transformer
    alphabet = "<" + "".join(chr(0x4e00 + i) for i in range(298)) + "$"
    model = Transformer("character", alphabet=alphabet)
    assert model.word_n_class == 300 and model.generator_word.proj.weight.shape == (300, 1024)
    assert model.embedding_word.lut.weight.shape == (300, 512)
    with __import__("pytest").raises(ValueError):
        Transformer("character")
    spec = dict(synth.load_spec("sld"))
    spec["embedding_word.lut.weight"], spec["generator_word.proj.weight"], spec["generator_word.proj.bias"] = [300, 512], [300, 1024], [300]
    sd = synth.synth_state_dict(spec, 99)
    model.load_state_dict(sd, strict=False)
    model.train()
    model.dropout_p = 0.0
    image, _ = SO.synth_batch(2)
    length = torch.tensor([2, 2])                                    # character mode: one character + '$' (train.py:93-96)
    text_input = torch.tensor([[0, 17], [0, 250]])
    text_gt = torch.tensor([17, 299, 250, 299])
    loss = model.loss(image, length, text_input, text_gt)
    osd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    o_loss, *_ = SO.loss_fn(osd, image, length, text_input, text_gt)
    assert abs(float(loss) - float(o_loss)) < 1e-4 * float(o_loss)
    loss.backward()
    o_loss.backward()
    for k in ("generator_word.proj.weight", "generator_word.proj.bias", "embedding_word.lut.weight"):
        a, b = dict(model.named_parameters())[k].grad, osd[k].grad
        assert float((a - b).norm() / b.norm()) < 2e-3, k


def test_gradient_sinks_write_in_place_and_leave_autograd_alone_outside(monkeypatch):
    """inside `grad_sinks` (what the fused trainers wrap their backward in) every parameter gradient is written by the backward
    body straight into the published view and equals autograd's own; outside, a second backward ACCUMULATES as torch defines"""
    model, g, image = _setup(monkeypatch)
    from fudanocr_b200.model.transformer import grad_sinks
    model.train()
    model.dropout_p = 0.0

    def loss_fn():
        out = model(image, g["length"], g["text_input"])
        return torch.nn.CrossEntropyLoss()(out["pred"], g["text_gt"])
    model.zero_grad(set_to_none=True)
    loss_fn().backward()
    plain = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
    loss_fn().backward()                                                    # no sinks: autograd accumulates
    for k, p in model.named_parameters():
        if k in plain:
            assert torch.allclose(p.grad, 2 * plain[k], rtol=1e-5, atol=1e-8), k
    # the trainer's arrangement: .grad views of one flat buffer, published by parameter data pointer
    names = [k for k in plain]
    params = dict(model.named_parameters())
    flat = torch.zeros(sum(params[k].numel() for k in names))
    table, off = {}, 0
    for k in names:
        n = params[k].numel()
        params[k].grad = flat[off:off + n].view_as(params[k])
        table[params[k].data_ptr()] = params[k].grad
        off += n
    flat.fill_(7.0)                                                         # stale content must be overwritten, not added to
    with grad_sinks(table):
        loss_fn().backward()
    direct = [k for k in names if "generator_word" not in k]                # the generator's weight enters padded (F.pad): autograd's path
    for k in direct:
        assert torch.allclose(params[k].grad, plain[k], rtol=1e-5, atol=1e-8), k
    for k in names:
        if k not in direct:
            assert torch.allclose(params[k].grad, plain[k] + 7.0, rtol=1e-5, atol=1e-6), k
    from fudanocr_b200.model import transformer as T
    assert T._GRAD_SINKS is None                                            # the table is gone once the step is over


def test_oracle_and_assembled_model_on_32x320_crops_match_the_reference_golden(monkeypatch):
    """BASELINE configs[3]: 32 x 320 crops -> 16 x 160 maps, 2 560 image tokens.  tests/golden/sld_w320_b2.pt was recorded from the
    UNMODIFIED reference module (oracle/make_golden_sld.py --w320); the oracle restatement - the checker of the GPU tests of that
    configuration - and the assembled drop-in model (kernel wrappers swapped for the torch stand-ins) must reproduce its loss,
    predictions, attention-map samples and every gradient norm"""
    import _recog_mock
    _recog_mock.install(monkeypatch)
    from fudanocr_b200.model.transformer import Transformer
    g = torch.load(synth.GOLDEN_DIR / "sld_w320_b2.pt", weights_only=False)
    parts = [SO.synth_batch(g["B"], seed=1234 + 17 * i) for i in range(10)]
    image = torch.cat([p[0] for p in parts], dim=3)
    assert image.shape == (g["B"], 3, 32, 320) and parts[0][1] == g["strings"]
    assert abs(float(image.double().sum()) - g["image_checksum"]) < 1e-6
    sd = synth.synth_state_dict(synth.load_spec("sld"), 1234)
    # (1) the oracle
    osd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in sd.items()}
    o_loss, o_logits, o_map, o_conv = SO.loss_fn(osd, image, g["length"], g["text_input"], g["text_gt"])
    o_loss.backward()
    assert abs(float(o_loss) - float(g["loss"])) < 1e-5 * float(g["loss"])
    assert torch.allclose(SO.pack(o_logits, g["length"]), g["pred"], rtol=1e-4, atol=1e-5)
    assert o_map.shape[-1] == 2560 and torch.allclose(o_map[:, :, :, ::64], g["map_sample"], rtol=1e-4, atol=1e-7)
    nmax = max(float(n) for n in g["grad_norms"].values() if n is not None)
    for k, n in g["grad_norms"].items():
        if n is None:
            assert osd[k].grad is None, k
        elif float(n) > 1e-6 * nmax:
            assert abs(float(osd[k].grad.norm()) - float(n)) < 5e-3 * float(n), k
    # (2) the assembled model
    model = Transformer("stroke")
    model.load_state_dict(sd, strict=False)
    model.train()
    model.dropout_p = 0.0
    out = model(image, g["length"], g["text_input"])
    assert out["conv"].shape == (g["B"], 1024, 16, 160)
    assert torch.allclose(out["pred"], g["pred"], rtol=1e-3, atol=1e-4)
    assert torch.allclose(out["map"][:, :, :, ::64], g["map_sample"], rtol=1e-3, atol=1e-6)
    loss = torch.nn.CrossEntropyLoss()(out["pred"], g["text_gt"])
    assert abs(float(loss) - float(g["loss"])) < 1e-4

"""Golden vectors for the CRNN evaluator + greedy CTC decode, produced by the UNMODIFIED reference modules
(scene-text-telescope/model/crnn/crnn.py, utils/utils_crnn.py) on CPU.  Run in the build container only."""
from __future__ import annotations

import hashlib
import json
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path(os.environ.get("FOCR_REFERENCE", "/root/reference"))
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF / "scene-text-telescope"))

from oracle import synth, crnn_oracle as C  # noqa: E402


def main():
    import types
    ipy = types.ModuleType("IPython")                    # shim: utils/util.py does `from IPython import embed`
    ipy.embed = lambda *a, **k: None
    sys.modules.setdefault("IPython", ipy)
    from model.crnn import crnn as refcrnn              # reference CRNN
    from utils.utils_crnn import strLabelConverter       # reference decoder
    torch.manual_seed(0)
    model = refcrnn.CRNN(32, 1, 37, 256).eval()          # base.py:310
    spec = {k: list(v.shape) for k, v in model.state_dict().items()}
    gd = synth.GOLDEN_DIR
    (gd / "crnn_spec.json").write_text(json.dumps(spec, indent=0))
    sd = synth.synth_state_dict(spec, seed=4321)
    # LSTM gates saturate with unit-variance inputs: keep the synthetic recurrent weights modest but the output
    # layer sharp so that argmax paths contain repeats, blanks and ties-free maxima
    model.load_state_dict(sd)
    B = 2
    _, hr = synth.synth_images(B, seed=99)
    sr = hr * 2 - 1                                      # an SR-like image in [-1,1] (tanh range)
    # parse_crnn_data is a method of TextBase (interfaces/base.py:319-325); its body is torch-only, restated verbatim
    # in the oracle and checked here against the same torch calls
    gray_ref = torch.nn.functional.interpolate(sr, (32, 100), mode="bicubic")
    gray_ref = 0.299 * gray_ref[:, 0:1] + 0.587 * gray_ref[:, 1:2] + 0.114 * gray_ref[:, 2:3]
    assert torch.equal(C.parse_crnn_data(sr), gray_ref)
    with torch.no_grad():
        logits = model(gray_ref)                         # (26, B, 37)
        o_logits = C.crnn_forward(sd, gray_ref)
    assert logits.shape == (26, B, 37)
    assert torch.allclose(o_logits, logits, atol=2e-5, rtol=1e-4), (o_logits - logits).abs().max()
    # decode with the reference converter (utils_crnn.py:54-89): alphabet '0-9a-z', blank = 0
    conv = strLabelConverter("0123456789abcdefghijklmnopqrstuvwxyz")
    path = logits.max(2)[1]                              # (26, B)
    preds = path.transpose(1, 0).contiguous().view(-1)
    strs = conv.decode(preds, torch.IntTensor([26] * B), raw=False)
    o_path = C.greedy_path(logits)
    assert torch.equal(o_path, path.t())
    o_strs = C.get_crnn_pred(logits.permute(1, 0, 2))
    assert o_strs == list(strs), (o_strs, strs)
    assert ["".join(C.ALPHABET[i] for i in C.ctc_collapse(p.tolist())) for p in o_path] == list(strs)
    out = {"sr": sr, "gray": gray_ref, "logits": logits, "path": path.t().contiguous(), "strings": list(strs)}
    torch.save(out, gd / "crnn_b2.pt")
    h = hashlib.sha256((gd / "crnn_b2.pt").read_bytes()).hexdigest()
    sums = [ln for ln in (gd / "SHA256SUMS").read_text().splitlines() if "crnn_b2.pt" not in ln]
    (gd / "SHA256SUMS").write_text("\n".join(sums + [f"{h}  crnn_b2.pt"]) + "\n")
    print("crnn golden:", strs, "path0", path[:, 0].tolist())


if __name__ == "__main__":
    main()

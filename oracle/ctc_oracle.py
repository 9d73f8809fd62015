"""ORACLE (test infrastructure only).  numpy restatement of CTC forward-backward (Graves et al. 2006, eqs. 6-8, 10-11, 16)
with the conventions of ``torch.nn.functional.ctc_loss`` - the third-party call the north star's "CTC forward-backward"
stands for; the reference has NO CTC-loss call site of its own (its CRNN is eval-only,
scene-text-telescope/interfaces/super_resolution.py:143-158, model/crnn/crnn.py:78-80 gives the (T, B, C) layout).
Dependency: torch (pinned ``torch==1.2.0`` in scene-text-telescope/requirement.txt:12; this container runs 2.11 whose
``ctc_loss`` keeps the same definition).  Pinned by tests/golden/ctc.npz = outputs of ``F.ctc_loss`` + autograd recorded by
oracle/make_golden_ctc.py.  float64 throughout; the input is RAW logits (log-softmax is part of the restatement), the
gradient is with respect to the logits."""
from __future__ import annotations

import numpy as np

NEG = -np.inf


def _lse(*xs):
    m = max(xs)
    if m == NEG:
        return NEG
    return m + np.log(sum(np.exp(x - m) for x in xs))


def log_softmax(logits: np.ndarray) -> np.ndarray:
    x = logits.astype(np.float64)
    m = x.max(-1, keepdims=True)
    return x - (m + np.log(np.exp(x - m).sum(-1, keepdims=True)))


def ctc_sample(lp: np.ndarray, target: np.ndarray, blank: int = 0):
    """lp (Tb, C) log-probabilities of ONE sample (already cut to its input length), target (S,) ->
    (nll, d nll / d logits (Tb, C)); nll = inf (gradient zeros) when no alignment exists"""
    Tb, C = lp.shape
    S = len(target)
    ext = np.full(2 * S + 1, blank, np.int64)
    ext[1::2] = target
    L = len(ext)
    if Tb == 0:
        return (0.0 if S == 0 else np.inf), np.zeros((0, C))
    alpha = np.full((Tb, L), NEG)
    alpha[0, 0] = lp[0, ext[0]]
    if L > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, Tb):
        for s in range(L):
            terms = [alpha[t - 1, s]]
            if s >= 1:
                terms.append(alpha[t - 1, s - 1])
            if s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]:
                terms.append(alpha[t - 1, s - 2])
            v = _lse(*terms)
            alpha[t, s] = NEG if v == NEG else v + lp[t, ext[s]]
    ll = _lse(alpha[Tb - 1, L - 1], alpha[Tb - 1, L - 2]) if L > 1 else alpha[Tb - 1, 0]
    if ll == NEG:
        return np.inf, np.zeros((Tb, C))
    beta = np.full((Tb, L), NEG)
    beta[Tb - 1, L - 1] = lp[Tb - 1, ext[L - 1]]
    if L > 1:
        beta[Tb - 1, L - 2] = lp[Tb - 1, ext[L - 2]]
    for t in range(Tb - 2, -1, -1):
        for s in range(L):
            terms = [beta[t + 1, s]]
            if s + 1 < L:
                terms.append(beta[t + 1, s + 1])
            if s + 2 < L and ext[s] != blank and ext[s] != ext[s + 2]:
                terms.append(beta[t + 1, s + 2])
            v = _lse(*terms)
            beta[t, s] = NEG if v == NEG else v + lp[t, ext[s]]
    grad = np.exp(lp)
    for t in range(Tb):
        for k in set(ext.tolist()):
            idx = np.nonzero(ext == k)[0]
            v = _lse(*[alpha[t, s] + beta[t, s] for s in idx])
            if v != NEG:
                grad[t, k] -= np.exp(v - ll - lp[t, k])
    return -ll, grad


def ctc_loss(logits: np.ndarray, targets: np.ndarray, input_lengths, target_lengths, blank: int = 0,
             reduction: str = "mean", zero_infinity: bool = False):
    """logits (T, B, C) raw scores; targets (B, S_max) padded -> (loss, nll (B,), d loss / d logits (T, B, C))"""
    T, B, C = logits.shape
    lp = log_softmax(logits)
    nll = np.zeros(B)
    grad = np.zeros((T, B, C))
    for b in range(B):
        Tb, S = int(input_lengths[b]), int(target_lengths[b])
        n, g = ctc_sample(lp[:Tb, b], np.asarray(targets[b][:S], np.int64), blank)
        if np.isinf(n):
            if zero_infinity:
                n = 0.0
            g = np.zeros_like(g)
        nll[b] = n
        scale = 1.0 / (B * max(S, 1)) if reduction == "mean" else 1.0
        grad[:Tb, b] = g * scale
    if reduction == "mean":
        loss = float(np.mean(nll / np.maximum(np.asarray(target_lengths, np.float64), 1.0)))
    elif reduction == "sum":
        loss = float(nll.sum())
    else:
        loss = nll.copy()
    return loss, nll, grad


def synth_case(T: int, B: int, C: int, S_max: int, seed: int, ragged: bool = True, repeats: bool = True):
    """deterministic logits / padded targets / lengths; `repeats` plants doubled characters (the s-2 skip rule)"""
    rs = np.random.RandomState(seed)
    logits = (rs.randn(T, B, C) * 2.0).astype(np.float32)
    tl = rs.randint(0 if ragged else S_max, S_max + 1, size=B).astype(np.int64)
    il = (rs.randint(max(T // 2, 1), T + 1, size=B) if ragged else np.full(B, T)).astype(np.int64)
    targets = rs.randint(1, C, size=(B, max(S_max, 1))).astype(np.int64)
    if repeats and S_max >= 2:
        targets[::2, 1] = targets[::2, 0]
    for b in range(B):  # keep every sample feasible: needs S + (number of adjacent repeats) <= Tb
        S = int(tl[b])
        rep = int((targets[b, 1:S] == targets[b, :S - 1]).sum()) if S > 1 else 0
        il[b] = max(il[b], min(T, S + rep))
        if S + rep > T:
            tl[b] = max(0, T // 2 - 1)
    return logits, targets[:, :max(S_max, 1)], il, tl

"""Statistical verification -- the Monte-Carlo outer loop of deepbayes/analyzers/verifiers.py
(:517-653): Chernoff / Okamoto sample bound, Massart sequential halting, Clopper-Pearson
interval, and the per-sample "forward on perturbed inputs -> predicate -> count" evaluated
for ALL samples in one device call (dbx_count_predicate) instead of one Python iteration per
posterior sample.  The IBP worst-case evaluation (verify=True, :593-619: `dbx_ibp_predict` / `dbx_ibp_samples`) and the FGSM
inner attack (:622-625: `dbx_count_predicate_fgsm` / `dbx_fgsm_samples`) run on the device too; interval bounds default to
fp32 arithmetic (bf16 is opt-in and not sound).
"""
from __future__ import annotations

import math

import numpy as np


def okamoto_bound(epsilon, delta):
    """verifiers.py:517-518 (also known as the Chernoff bound)."""
    return (-1 * .5) * math.log(float(delta) / 2) * (1.0 / (epsilon ** 2))


def absolute_massart_halting(succ, trials, I, epsilon, delta, alpha):
    """verifiers.py:521-532."""
    if I[0] < 0.5 and I[1] > 0.5:
        return -1
    elif I[1] < 0.5:
        val = I[1]
        h = (9 / 2.0) * (((3 * val + epsilon) * (3 * (1 - val) - epsilon)) ** (-1))
        return math.ceil((h * (epsilon ** 2)) ** (-1) * math.log((delta - alpha) ** (-1)))
    elif I[0] >= 0.5:
        val = I[0]
        h = (9 / 2.0) * (((3 * (1 - val) + epsilon) * ((3 * val) + epsilon)) ** (-1))
        return math.ceil((h * (epsilon ** 2)) ** (-1) * math.log((delta - alpha) ** (-1)))


def proportion_confint_beta(count, nobs, alpha=0.05):
    """statsmodels' proportion_confint(method='beta') (Clopper-Pearson), verifiers.py:636."""
    from scipy.stats import beta
    lb = beta.ppf(alpha / 2, count, nobs - count + 1)
    ub = beta.ppf(1 - alpha / 2, count + 1, nobs - count)
    return float(lb), float(ub)


def chernoff_sample_bound(confidence=0.95, delta=0.3):
    """verifiers.py:542-543 / :579-580."""
    epsilon = 1 - confidence
    return math.ceil((1 / (2 * epsilon ** 2)) * math.log(2 / delta))


def _mode_of(model, override=None):
    """Arithmetic mode of the device FORWARD path: PosteriorModel.mode / Optimizer.compute_mode (an optimizer's own
    `.mode` is the reference's "classification" / "regression" switch, optimizer.py:60).  Interval bounds do not use this:
    they default to fp32 (see chernoff_bound_verification / massart_bound_check)."""
    for m in (override, getattr(model, "compute_mode", None), getattr(model, "mode", None)):
        if m in ("bf16", "fp32"):
            return m
    return "bf16"


def _engine_of(model):
    eng = model.engine() if callable(getattr(model, "engine", None)) else getattr(model, "_engine", None)
    if eng is None:
        raise TypeError("need a deepbayes_b200 PosteriorModel or optimizer")
    return eng


def statistical_robustness(model, adv_inputs, labels, n=None, confidence=0.95, delta=0.3, alpha=0.05, mode=None,
                           return_flags=False):
    """BASELINE config 5: n posterior samples x K fixed (pre-perturbed) inputs, predicate
    ``argmax == label`` (the tutorials' predicate), all on the device.  n defaults to the
    Chernoff bound (verifiers.py:542-543).  Returns dict(p=success fraction per input,
    counts, n, interval=Clopper-Pearson per input[, flags])."""
    eng = _engine_of(model)
    n = chernoff_sample_bound(confidence, delta) if n is None else int(n)
    mode = _mode_of(model, mode)
    s0 = model._take_ids(n)
    res = eng.count_predicate(adv_inputs, labels, n, model.seed, s0, None, 1.0, mode, return_flags)
    counts = (res[0] if return_flags else res).cpu().numpy()
    out = {"counts": counts, "n": n, "p": counts / float(n),
           "interval": [proportion_confint_beta(int(c), n, alpha) for c in counts]}
    if return_flags:
        out["flags"] = res[1].cpu().numpy()
    return out


def _flat(weights):
    return np.concatenate([np.asarray(w, np.float32).ravel() for w in weights])


def IBP(model, inp, weights, eps, predict=False):
    """`IBP(model, inp, weights, eps, predict=False)` (verifiers.py:68-95) on the device: the logit interval
    (h_l, h_u) of the eps-ball around ``inp`` (clipped to [0,1]) under the explicit weight list ``weights``.
    predict=True (:73-74, :92-94; the regression branch): the box inp -+ eps is NOT clipped and the last layer's
    activation is applied to both bounds.  Dense / Conv2D (`propagate_conv2d`, :11-21) / pooling / Flatten layers; fp32
    arithmetic (the reference's Dense part uses float64)."""
    if predict:
        eng = _engine_of(model)
        a = np.asarray(inp, np.float32)
        x2 = a.reshape(-1, eng.K)
        e = np.float32(eps)
        lo, hi = eng.ibp_state(x2 - e, x2 + e, _flat(weights), 0.0)
        name = [l for l in eng.model.layers if l.kind in ("dense", "conv2d")][-1].activation.name
        shape = a.shape[:-1] + (eng.C,)
        from ..engine import device_activation
        return (device_activation(name, lo.cpu().numpy()).reshape(shape), device_activation(name, hi.cpu().numpy()).reshape(shape))
    eng = _engine_of(model)
    a = np.asarray(inp, np.float32)
    lo, hi = eng.ibp_one(a.reshape(-1, eng.K), eps, _flat(weights), mode="fp32")
    shape = a.shape[:-1] + (eng.C,)
    return lo.cpu().numpy().reshape(shape), hi.cpu().numpy().reshape(shape)


def IBP_state(model, s0, s1, weights, weight_margin=0, logits=True):
    """`IBP_state(model, s0, s1, weights, weight_margin, logits=True)` (verifiers.py:97-123; the same function is
    probverification.IBP_prob, :84-110): logit bounds of the input box [s0, s1] under weights widened by
    weight_margin * posterior_var (the posterior's scale, used as a std like everywhere in the reference).  Needs a
    Gaussian posterior on the model; fp32."""
    eng = _engine_of(model)
    a0, a1 = np.asarray(s0, np.float32), np.asarray(s1, np.float32)
    lo, hi = eng.ibp_state(a0.reshape(-1, eng.K), a1.reshape(-1, eng.K), _flat(weights), float(weight_margin))
    shape = a0.shape[:-1] + (eng.C,)
    return lo.cpu().numpy().reshape(shape), hi.cpu().numpy().reshape(shape)


IBP_prob = IBP_state


def chernoff_bound_verification(model, inp, eps, cls, **kwargs):
    """`chernoff_bound_verification` (verifiers.py:537-557): n = ceil(ln(2/delta) / (2 (1-confidence)^2)) posterior
    samples; for each, IBP logits of the eps-ball, worst case w.r.t. ``cls`` (lower bound of the true class, upper
    bound of the others), the last layer's activation; returns the SUM over samples, like the reference (:556-557).
    ``inp`` may hold several rows, with one class per row (the reference takes one input).  All n samples run in one
    device call.  The interval arithmetic runs in fp32 by default (the reference uses float64, verifiers.py:26-33);
    mode="bf16" (tensor cores) is opt-in and NOT a sound bound: its operands are rounded to bf16, to nearest."""
    delta = kwargs.get('delta', 0.3)
    confidence = kwargs.get('confidence', 0.95)
    eng = _engine_of(model)
    n = chernoff_sample_bound(confidence, delta)
    mode = kwargs.get('mode') or "fp32"
    a = np.asarray(inp, np.float32)
    x2 = a.reshape(-1, eng.K)
    labels = np.broadcast_to(np.asarray(cls, np.int32).reshape(-1), (x2.shape[0],))
    s0 = model._take_ids(n)
    # samples are independent: with model.shard the n global sample ids are split over the torch.distributed ranks and ONE
    # all-reduce(SUM) of the [B,C] buffer combines them (same ids for every world size)
    from .. import distributed as _dd
    rank, ws = _dd.world() if getattr(model, "shard", False) else (0, 1)
    off, cnt = _dd.shard_range(n, rank, ws)
    if cnt > 0:
        worst = eng.ibp_predict(x2, eps, labels, cnt, model.seed, s0 + off, None, 1.0, mode, want=("worst",))["worst"]
    else:
        worst = eng.torch.zeros((x2.shape[0], eng.C), dtype=eng.torch.float32, device=eng.device)
    if ws > 1:
        _dd.allreduce_sum_(worst)
    return worst.cpu().numpy().reshape(a.shape[:-1] + (eng.C,))


def _softmax_rows(z):
    z = np.asarray(z, np.float64)
    e = np.exp(z - z.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


def massart_bound_check(model, inp, eps, predicate=None, **kwargs):
    """verifiers.py:563-653, same signature and defaults (``verify=True``, ``classification=False``, ``chernoff=False``):
    every iteration draws one posterior sample, builds its ``worst_case`` and calls ``predicate(worst_case)``; the loop stops
    at the Massart halting bound (or, with chernoff=True, after the Chernoff bound).  Here the device evaluates whole BLOCKS
    of posterior samples per call and hands back every sample's own output (`dbx_ibp_samples`, `dbx_fgsm_samples`,
    `dbx_predict_samples`); the predicate and the halting rule then walk the block in sample order, so the result equals the
    sequential loop on the same sample sequence (samples past the halting point are discarded).

    worst_case per branch, as in the reference:
      verify and classification (:593-608)  softmax of (upper logit bound everywhere, lower bound at ``cls``), one-hot depth 10
                                            (depth 2 when that does not broadcast), shape like the model output
      verify and not classification (:609-619)  IBP(..., predict=True) on the UN-clipped box, last activation applied to the
                                            bounds; per output the bound that lies farther from ``cls``
      not verify (:621-625)                 model._predict(FGSM(model, inp, loss, eps, samples=1)): each posterior sample
                                            attacked with its own gradient (for a PosteriorModel the reference's FGSM
                                            differentiates through 35 fresh samples of predict(); the per-sample form is what
                                            it does for an optimizer object); ``adv=`` (not in the reference) skips the attack
    ``predicate=None`` (not in the reference) selects the device predicate ``argmax == cls`` (the tutorials' predicate)
    without materialising the per-sample outputs.  chernoff=True also returns the mean of model._predict(glob_adv) over the
    iterations, glob_adv = FGSM(..., samples=35) (:585, :647-648).  Returns (successes/iterations, iterations,
    mean/iterations)."""
    delta = kwargs.get('delta', 0.3)
    alpha = kwargs.get('alpha', 0.05)
    confidence = kwargs.get('confidence', 0.95)
    verify = kwargs.get('verify', True)                       # :568
    classification = kwargs.get('classification', False)      # :569
    chernoff_override = kwargs.get('chernoff', False)         # :576
    cls = kwargs.get('cls', None)
    adv = kwargs.get('adv', None)
    loss_fn = kwargs.get('loss_fn', None)
    block = int(kwargs.get('block', 4096))
    if predicate is None and cls is None:
        raise ValueError("pass a predicate (verifiers.py:563) or cls=<true class> for the device predicate argmax == cls")
    if verify and cls is None:
        raise ValueError("verify=True needs cls= (verifiers.py:596-603 / :607-608 build the worst case around it)")
    if predicate is None and verify and not classification:
        raise ValueError("verify=True with classification=False (verifiers.py:609-619) needs a predicate: cls is a regression target "
                         "there, not a class index; pass classification=True or verify=False for the device predicate argmax == cls")
    eng = _engine_of(model)
    # bounds are computed in fp32 unless the caller opts into bf16 (narrower arithmetic than the reference's float64: not sound)
    mode = kwargs.get('mode') or ("fp32" if verify else _mode_of(model))
    a = np.asarray(inp, np.float32)
    x1 = a.reshape(1, -1)
    _, oshape = eng.model.rows_and_outshape(a)
    C = eng.C
    epsilon = 1 - confidence
    chernoff_bound = math.ceil((1 / (2 * epsilon ** 2)) * math.log(2 / delta))
    successes, iterations = 0.0, 0.0
    halting_bound = chernoff_bound
    mean = 0.0 * np.asarray(model.predict(inp))               # :583 (consumes posterior samples like the reference)
    glob_adv = None
    if chernoff_override:                                     # :584-585
        from . import attacks
        glob_adv = attacks.FGSM(model, inp, loss_fn, eps, samples=35)
    direc = None
    fast = predicate is None and (classification or not verify)
    done = False
    while not done:
        nb = int(min(block, chernoff_bound + 1 - iterations))
        s0 = model._take_ids(nb)
        outcomes = None
        if verify and classification:
            if fast:
                outcomes = eng.ibp_predict(x1, eps, [int(cls)], nb, model.seed, s0, None, 1.0, mode, want=(),
                                           return_flags=True)["flags"].cpu().numpy().reshape(-1)
            else:
                lo, hi = eng.ibp_samples(x1, eps, nb, model.seed, s0, mode=mode)
                lo, hi = lo.cpu().numpy().reshape(nb, C), hi.cpu().numpy().reshape(nb, C)
                depth = 10 if C == 10 else 2                   # :596-603 (one_hot depth 10, else the depth-2 retry)
                if depth != C:
                    raise ValueError("verifiers.py:596-603 builds its one-hot with depth 10 or 2; the model has %d outputs" % C)
                v1 = np.zeros(C, np.float32); v1[int(cls)] = 1.0
                worst = _softmax_rows((1.0 - v1) * hi + v1 * lo).astype(np.float32)
        elif verify:
            e = np.float32(eps)
            lo, hi = eng.ibp_samples(x1 - e, 0.0, nb, model.seed, s0, mode=mode, clip=False, box_hi=x1 + e)
            from ..engine import device_activation
            name = [l for l in eng.model.layers if l.kind in ("dense", "conv2d")][-1].activation.name
            lo = device_activation(name, lo.cpu().numpy().reshape(nb, C)); hi = device_activation(name, hi.cpu().numpy().reshape(nb, C))
            target = np.asarray(cls, np.float32)
            above, below = np.abs(hi - target), np.abs(lo - target)
            worst = np.where(above > below, hi, lo)            # :610-618
        elif adv is not None:
            if fast:
                _, fl = eng.count_predicate(np.asarray(adv).reshape(1, -1), [int(cls)], nb, model.seed, s0, None, 1.0, mode, True)
                outcomes = fl.cpu().numpy().reshape(-1)
            else:
                worst = eng.predict_samples(np.asarray(adv).reshape(1, -1), nb, model.seed, s0, mode=mode).cpu().numpy().reshape(nb, C)
        else:
            if direc is None:                                  # FGSM's default direction=-1: the model's own prediction
                direc = int(np.argmax(np.asarray(model.predict(inp)).reshape(-1)))
            lower, upper = getattr(model, "input_lower", -np.inf), getattr(model, "input_upper", np.inf)
            if fast:
                _, fl = eng.count_predicate_fgsm(x1, [direc], [int(cls)], eps, nb, model.seed, s0, lower=lower, upper=upper,
                                                 return_flags=True, mode=mode)
                outcomes = fl.cpu().numpy().reshape(-1)
            else:
                worst = eng.fgsm_samples(x1, [direc], eps, nb, model.seed, s0, lower=lower, upper=upper, mode=mode).cpu().numpy().reshape(nb, C)
        gl = None
        if chernoff_override:                                  # model._predict(glob_adv) under the same posterior samples (:648)
            gl = eng.predict_samples(np.asarray(glob_adv, np.float32).reshape(1, -1), nb, model.seed, s0,
                                     mode=_mode_of(model, kwargs.get('mode'))).cpu().numpy().reshape(nb, C)
        pos = 0
        while pos < nb:
            if not (iterations <= halting_bound):
                done = True
                break
            if outcomes is not None:
                result = float(outcomes[pos])
            else:
                result = 1.0 if predicate(worst[pos].reshape(oshape)) else 0.0    # :628
            successes += result
            iterations += 1
            lb, ub = proportion_confint_beta(successes, iterations)
            lb = 0.0 if math.isnan(lb) else lb
            ub = 1.0 if math.isnan(ub) else ub
            hb = absolute_massart_halting(successes, iterations, [lb, ub], epsilon, delta, alpha)
            halting_bound = chernoff_bound if (hb == -1 or chernoff_override) else min(hb, chernoff_bound)   # :643-646
            if chernoff_override:
                mean = mean + gl[pos].reshape(oshape)          # :647-648
            pos += 1
        if not (iterations <= halting_bound):
            done = True
    return successes / iterations, iterations, mean / iterations

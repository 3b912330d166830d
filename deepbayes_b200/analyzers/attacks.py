"""First-order attacks over the posterior -- deepbayes/analyzers/attacks.py: ``gradient_expectation`` (:49-75),
``FGSM`` (:81-102), ``PGD`` / ``_PGD`` (:133-182), ``CW`` (:179-226) -- with the input gradient evaluated on the device
(`dbx_input_gradient`): all samples of one ``model.predict`` side by side; forward and backward run in the model's arithmetic
mode (PosteriorModel.mode / Optimizer.compute_mode): "bf16" = per-sample tcgen05 GEMMs both ways, "fp32" = FFMA kernels.

Scope: the sparse categorical cross-entropy (the tutorials' attack loss).  order=1 (the backward pass on the device): Dense MLPs
with a softmax output.  order=0 (``zeroth_order_gradient``, attacks.py:17-45): central differences of the loss through
``model.predict`` alone, so ANY model the forward path takes (convolutions included) -- 2 x feature_size perturbed inputs per image
go through predict as two batches.  As in the reference the model may be a PosteriorModel (its predict() is the mean of 35 sampled forwards, so each
iteration differentiates through 35 samples) or an optimizer object (predict() = one forward with the model's weights).
"""
from __future__ import annotations

import numpy as np

from .verifiers import _engine_of, _mode_of

_PM_PREDICT_N = 35          # PosteriorModel.predict's default n (posteriormodel.py:81)


def _check_loss(loss_fn):
    name = getattr(loss_fn, "__name__", type(loss_fn).__name__).lower() if loss_fn is not None else "sparse_categorical_crossentropy"
    if "sparse" not in name or "crossentropy" not in name:
        raise NotImplementedError("the device input gradient builds the sparse categorical cross-entropy only (got %r)" % name)


def _labels_of(model, inp, direction):
    if type(direction) == int and direction == -1:          # attacks.py:86-91 / :139-144: the model's own prediction
        return np.argmax(np.asarray(model.predict(inp)).reshape(len(inp), -1), axis=1)
    d = np.asarray(direction).reshape(-1)
    return np.broadcast_to(d, (len(inp),)) if d.size == 1 else d


def zeroth_order_gradient(model, inp, direction, loss_fn=None, num_models=25, step=0.35):
    """attacks.py:17-45: for every image, the loss at inp[i] +- step * e_k for each of its feature_size coordinates, evaluated by
    two ``model.predict(batch, n=num_models)`` calls (each with its own fresh posterior samples, like the reference), and
    grad[k] = (loss+ - loss-) / (2 step).  Returns an array shaped like ``inp``."""
    _check_loss(loss_fn)
    a = np.asarray(inp, np.float32)
    n_img, fshape = a.shape[0], a.shape[1:]
    F = int(np.prod(fshape))
    step_mask = (np.eye(F, dtype=np.float32) * np.float32(step)).reshape((-1,) + tuple(fshape))      # :23-29
    d = np.asarray(direction).reshape(-1)
    d = np.broadcast_to(d, (n_img,)) if d.size == 1 else d

    def row_loss(p, y):      # loss_fn(direction[i], prediction_j) for every perturbed input j (:36, :39)
        return -np.log(np.clip(np.asarray(p, np.float64).reshape(F, -1)[:, int(y)], 1e-7, 1 - 1e-7))
    grads = []
    for i in range(n_img):
        lp = row_loss(model.predict(a[i] + step_mask, n=num_models), d[i])
        lm = row_loss(model.predict(a[i] - step_mask, n=num_models), d[i])
        grads.append(((lp - lm) / (2.0 * step)).reshape(fshape))
    return np.asarray(grads, np.float32)


def gradient_expectation(model, inp, direction, loss_fn=None, num_models=10, mode=None):
    """attacks.py:49-75.  Returns the gradient SUM over the iterations, shaped like ``inp``.  ``mode`` (not in the reference)
    overrides the model's arithmetic mode for this call."""
    _check_loss(loss_fn)
    eng = _engine_of(model)
    a = np.asarray(inp, np.float32)
    x2 = a.reshape(-1, eng.K)
    labels = np.asarray(direction).reshape(-1)
    labels = np.broadcast_to(labels, (x2.shape[0],)) if labels.size == 1 else labels
    val = num_models
    n_outer = max(int(num_models), 1)
    is_pm = hasattr(model, "sample_based")                   # PosteriorModel
    if getattr(model, "det", False) or (not is_pm and val in (1, -1)):
        # no sampling: the model's current weights (:57-58); a deterministic PosteriorModel predicts with them too
        w = np.concatenate([np.asarray(t, np.float32).ravel() for t in model.model.get_weights()])
        g = eng.input_gradient_one(x2, labels, w)
        n_eff = 1 if (getattr(model, "det", False) or val == -1) else n_outer
        return (g * float(n_eff)).cpu().numpy().reshape(a.shape) if n_eff > 1 else g.cpu().numpy().reshape(a.shape)
    if val == -1:
        n_outer = 1                                          # :73-74
    n_inner = _PM_PREDICT_N if is_pm else 1
    s0 = model._take_ids(n_outer * n_inner)
    if is_pm and model.sample_based:
        # every sampled forward of predict() draws one stored sample with replacement (posteriormodel.py:77, :95-97)
        idx = np.random.choice(model.num_post_samps, size=n_outer * n_inner, p=model.frequency)
        g = eng.input_gradient(x2, labels, n_outer, n_inner, stored=True, idx=idx, mode=_mode_of(model, mode))
        return g.cpu().numpy().reshape(a.shape)
    g = eng.input_gradient(x2, labels, n_outer, n_inner, model.seed, s0, mode=_mode_of(model, mode))
    return g.cpu().numpy().reshape(a.shape)


def FGSM(model, inp, loss_fn=None, eps=0.1, direction=-1, num_models=1, order=1, samples=1):
    """attacks.py:81-102 (``samples`` overrides ``num_models``, :82); order 1 = gradient_expectation, order 0 =
    zeroth_order_gradient (:92-95)."""
    if order not in (0, 1):
        raise ValueError("order must be 0 or 1 (attacks.py:92-95)")
    num_models = samples
    inp = np.asarray(inp)
    maxi, mini = inp + eps, inp - eps
    direc = _labels_of(model, inp, direction)
    grad = np.sign(gradient_expectation(model, inp, direc, loss_fn, num_models) if order == 1
                   else zeroth_order_gradient(model, inp, direc, loss_fn, num_models))
    adv = np.clip(inp + eps * grad, mini, maxi)
    return np.clip(adv, getattr(model, "input_lower", -np.inf), getattr(model, "input_upper", np.inf))


def _PGD(model, inp, loss_fn=None, eps=0.1, direc=-1, step=0.1, num_steps=15, num_models=35, order=1):
    """attacks.py:133-165: random sign start of size eps/10, num_steps+1 sign-gradient steps of eps/num_steps, final clips."""
    if order != 1:      # the reference's loop computes a gradient for order == 1 only (:152-155: the other branch is commented out)
        raise NotImplementedError("_PGD builds its gradient for order=1 only (attacks.py:152-155)")
    inp = np.asarray(inp)
    direc = _labels_of(model, inp, direc)
    adv = np.asarray(inp, np.float32)
    maxi, mini = adv + eps, adv - eps
    adv = adv + ((eps / 10) * np.sign(np.random.normal(0.0, 1.0, size=adv.shape)))
    lo, hi = getattr(model, "input_lower", -np.inf), getattr(model, "input_upper", np.inf)
    adv = np.clip(adv, lo, hi)
    for _ in range(num_steps + 1):
        grad = np.sign(gradient_expectation(model, adv, direc, loss_fn, num_models))
        adv = (adv + grad * (eps / float(num_steps))).astype(np.float32)
    adv = np.clip(adv, mini, maxi)
    return np.clip(adv, lo, hi)


def PGD(model, inp, loss_fn=None, eps=0.1, direction=-1, step=0.1, num_steps=5, num_models=-1, order=1, restarts=0):
    """attacks.py:167-177: restarts + 1 runs of _PGD; the single result, or the list of all of them when restarts > 0."""
    advs = []
    for _ in range(restarts + 1):
        adv = _PGD(model, inp, loss_fn, eps, direc=direction, step=step, num_steps=num_steps, num_models=num_models, order=order)
        advs.append(adv)
    return adv if restarts == 0 else advs


def CW(model, inp, loss_fn=None, eps=0.01, num_steps=5, num_models=25, direction=-1, stepsize=None, decay1=0.99, decay2=0.9999,
       epsilon=1e-04, order=1):
    """attacks.py:179-226: num_steps sign steps of size eps / num_steps along an Adam-style direction
    moment1 / (sqrt(moment2) + epsilon) of the expected input gradient, clipped to the model's input range after every step and
    to the eps-ball at the end.  As written in the reference: the bias-CORRECTED moments are what the next iteration decays
    (:207-211 overwrite moment1 / moment2), ``stepsize`` is ignored (:189), and only order=1 yields a gradient (:197-200)."""
    if order != 1:
        raise NotImplementedError("CW builds its gradient for order=1 only (attacks.py:197-200)")
    inp = np.asarray(inp)
    direc = _labels_of(model, inp, direction)
    moment1, moment2, time = 0.0, 0.0, 0
    adv = np.asarray(inp, np.float32)
    maxi, mini = adv + eps, adv - eps
    lo, hi = getattr(model, "input_lower", -np.inf), getattr(model, "input_upper", np.inf)
    for _ in range(num_steps):
        grad = np.asarray(gradient_expectation(model, adv, direc, loss_fn, num_models), np.float64)
        time += 1
        moment1 = (moment1 * decay1) + ((1 - decay1) * grad)
        moment2 = (moment2 * decay2) + ((1 - decay2) * (grad ** 2))
        moment1 = moment1 / (1 - (decay1 ** time))
        moment2 = moment2 / (1 - (decay2 ** time))
        d = np.sign(moment1 / (np.sqrt(moment2) + epsilon)) * (eps / float(num_steps))
        adv = np.clip(adv + d, lo, hi).astype(np.float32)
    adv = np.clip(adv, mini, maxi)
    return np.clip(adv, lo, hi)


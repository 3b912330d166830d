"""BayesByBackprop -- hot-path part of deepbayes/optimizers/bayesbybackprop.py:
``compile`` (:31-44, rho = log(exp(std)-1)), the sampling + forward half of ``step``
(:46-65, W = mu + softplus(rho)*eps then model(features)), ``sample`` (:176-181) and ``save``
(:184-195), and the full training step (:46-170, robust_train 0 / 1 / 2) as one device call (`dbx_bbb_step`).
"""
from __future__ import annotations

import numpy as np

from .. import formats
from . import optimizer
from .optimizer import softplus


class BayesByBackprop(optimizer.Optimizer):
    def __init__(self):
        super().__init__()

    def compile(self, keras_model, loss_fn=None, batch_size=64, learning_rate=0.15, decay=0.0,
                epochs=10, prior_mean=-1, prior_var=-1, **kwargs):
        super().compile(keras_model, loss_fn, batch_size, learning_rate, decay,
                        epochs, prior_mean, prior_var, **kwargs)
        # posterior_var now holds rho (bayesbybackprop.py:39-40)
        with np.errstate(divide="ignore"):
            self.posterior_var = [np.log(np.exp(v) - np.float32(1)).astype(np.float32) for v in self.posterior_var]
        self._rho_param = True
        self.kl_weight = kwargs.get('kl_weight', 1.0)
        return self

    def forward_sample(self, features, eps=None, return_weights=False):
        """bayesbybackprop.py:50-65 on the device: noise ~ N(0,1) per weight (or ``eps``
        injected, a list [eps_W0, eps_b0, ...] / flat vector), W = mean + softplus(rho)*noise
        formed in fp32 exactly in that order, forward of the mini-batch.  Returns predictions
        (and, if asked, the sampled weights and ``noise_used``, :57)."""
        self.sync_from_device()     # steps taken with sync=False left the live parameters on the device only
        eng = self.engine()
        _, oshape = self.model.rows_and_outshape(np.asarray(features))
        e = None if eps is None else eng._as_flat(eps)
        step = self._take_ids(1)
        if return_weights:
            out, W, noise = eng.bbb_forward(features, self.seed, step, e, self.compute_mode, True)
            Wn, nn = W.cpu().numpy(), noise.cpu().numpy()
            ws, ns, o = [], [], 0
            for s in self.model.weight_shapes():
                k = int(np.prod(s))
                ws.append(Wn[o:o + k].reshape(s)); ns.append(nn[o:o + k].reshape(s)); o += k
            self.model.set_weights(ws)  # :59
            return out.cpu().numpy().reshape(oshape), ws, ns
        out = eng.bbb_forward(features, self.seed, step, e, self.compute_mode, False)
        return out.cpu().numpy().reshape(oshape)

    # ------------------------------------------------------------------ training
    def _flat_dev(self, tensors):
        eng = self.engine()
        return eng._dev_f32(np.concatenate([np.asarray(t, np.float32).ravel() for t in tensors])).contiguous()

    def _unflat(self, flat):
        out, o = [], 0
        for shp in self.model.weight_shapes():
            k = int(np.prod(shp))
            out.append(flat[o:o + k].reshape(shp))
            o += k
        return out

    def step(self, features, labels, lrate, noise=None, sync=True, eps_draws=None):
        """bayesbybackprop.py:46-170 (robust_train 0 - 4) in ONE device call (`dbx_bbb_step`): draw noise, W = mean +
        softplus(rho) * noise, forward, KL loss (losses.py:40-45) with the sparse categorical cross-entropy as the data
        term, the reference's three gradients and its SGD update.  ``noise`` (list / flat vector) injects the N(0,1) draw.
        The parameters live on the device between steps; with sync=True (default) the updated lists are copied back,
        set on the object and returned like the reference does (:169-170), and the sampled weights are left in the model
        (:59).  robust_train: 1 = IBP (:73-90), 2 = FGSM (:91-101), 3 = IBP averaged over ``loss_mc`` draws of epsilon from
        Exponential(rate = 1 / epsilon) (:102-121; ``eps_draws`` injects them, otherwise they come from a NumPy generator
        seeded with the optimizer's seed), 4 = the FGSM pass repeated loss_mc times with the SAME epsilon (:123-136: the
        drawn eps is not used, :131), i.e. mode 2 with lambda = 0.  Other data losses raise NotImplementedError."""
        rt = int(getattr(self, "robust_train", 0))
        if rt not in (0, 1, 2, 3, 4):
            raise NotImplementedError("robust_train = %d (bayesbybackprop.py:68-136 knows 0 - 4)" % rt)
        epsilon, lam, mode, draws = float(getattr(self, "epsilon", 0.0)), float(getattr(self, "robust_lambda", 1.0)), rt, None
        if rt in (3, 4):
            self.epsilon = epsilon = max(0.0001, epsilon)       # :104, :125
            k = int(getattr(self, "loss_monte_carlo", 2))
            if rt == 3:
                if eps_draws is None:
                    if getattr(self, "_eps_rng", None) is None:
                        self._eps_rng = np.random.Generator(np.random.Philox(int(self.seed) + 0x3E95))
                    eps_draws = self._eps_rng.exponential(scale=epsilon, size=k)      # Exponential(rate 1 / epsilon): mean epsilon
                draws = np.asarray(eps_draws, np.float32).reshape(-1)
            else:
                mode, lam = 2, 0.0
        name = getattr(self.loss_func, "__name__", type(self.loss_func).__name__).lower() if self.loss_func is not None else "sparse_categorical_crossentropy"
        if "sparse" not in name or "crossentropy" not in name:
            raise NotImplementedError("the device step builds the sparse categorical cross-entropy data term only (got %r)" % name)
        eng = self.engine()
        key = (id(self.posterior_mean), id(self.posterior_var))
        if getattr(self, "_train_key", None) != key:      # (re)upload after compile / load / external assignment
            self._d_mean = self._flat_dev(self.posterior_mean).clone()
            self._d_rho = self._flat_dev(self.posterior_var).clone()
            self._d_pm = self._flat_dev(self.prior_mean).clone()
            self._d_pv = self._flat_dev(self.prior_var).clone()
        e = None if noise is None else eng._as_flat(noise)
        res = eng.bbb_step(features, labels, self._d_mean, self._d_rho, self._d_pm, self._d_pv, lrate, self.kl_weight, self.seed,
                           self._take_ids(1), e, want_probs=True, want_weights=sync, robust_mode=mode,
                           epsilon=epsilon, robust_lambda=lam, in_lower=self.input_lower, in_upper=self.input_upper, eps_draws=draws,
                           mode=self._train_mode())
        self._last_loss, self._last_probs = res["loss"], res["probs"]
        if sync:
            self.posterior_mean = self._unflat(self._d_mean.cpu().numpy())
            self.posterior_var = self._unflat(self._d_rho.cpu().numpy())
            self.model.set_weights(self._unflat(res["W"].cpu().numpy()))
        else:
            self._dirty = True
        self._train_key = (id(self.posterior_mean), id(self.posterior_var))
        return self.posterior_mean, self.posterior_var

    def _train_mode(self):
        """Arithmetic of the training step's GEMMs: "fp32" (default, what the reference computes) or "bf16" (compile(...,
        train_mode="bf16"): tensor cores for robust_train 0 / 2 / 4; parameters, losses and the update stay fp32)."""
        m = getattr(self, "train_mode", None) or "fp32"
        if m not in ("fp32", "bf16"):
            raise ValueError("train_mode must be 'fp32' or 'bf16' (got %r)" % (m,))
        return m

    def sync_from_device(self):
        """Copy the device-resident parameters back into posterior_mean / posterior_var (after steps with sync=False)."""
        if getattr(self, "_dirty", False):
            self.posterior_mean = self._unflat(self._d_mean.cpu().numpy())
            self.posterior_var = self._unflat(self._d_rho.cpu().numpy())
            self._train_key = (id(self.posterior_mean), id(self.posterior_var))
            self._dirty = False

    def train(self, X_train, y_train, X_test=None, y_test=None, **kwargs):
        """optimizer.py:100-133: epochs x mini-batches of `step`, learning rate lr / (1 + decay * epoch), Gaussian input
        noise (input_noise), per-epoch logging of the running loss / accuracy and of the validation pass with the weights
        the last step left in the model (model_validate, :139-141).  The reference's `shuffle(100)` windowed shuffle is
        replaced by a seeded permutation per epoch.  With linear_schedule (the default, optimizer.py:75) epsilon starts at 0 and
        grows by max_eps / epochs after every epoch (:105-108, :131-132) -- also when robust_train = 0, where it is unused."""
        X_train = np.asarray(X_train, np.float32)
        y_train = np.asarray(y_train).reshape(-1)
        self.N = len(X_train)
        rng = np.random.Generator(np.random.Philox(self.seed))
        bs = int(self.batch_size)
        self.loss_log, self.acc_log, self.val_log = [], [], []
        if self.robust_linear:      # optimizer.py:105-108: the attack / verification radius ramps up linearly from 0
            self.max_eps = self.epsilon
            self.epsilon = 0.0
            self.max_robust_lambda = self.robust_lambda
        # One device call per EPOCH (`dbx_bbb_train_epoch`): the training set stays in HBM, an epoch's permutation is applied
        # there, and nothing comes back before the epoch's last step.  The per-step loop remains for what needs the host
        # between steps: Gaussian input noise (drawn from the NumPy generator) and the epsilon draws of robust_train 3.
        rt = int(getattr(self, "robust_train", 0))
        device_epochs = kwargs.get("device_epochs", True) and not getattr(self, "input_noise", 0.0) and rt in (0, 1, 2, 4)
        if device_epochs:
            eng = self.engine()
            torch = eng.torch
            Xd = torch.from_numpy(np.ascontiguousarray(X_train.reshape(self.N, -1))).to(eng.device)
            yd = torch.from_numpy(y_train.astype(np.int32)).to(eng.device)
        for epoch in range(self.epochs):
            lrate = self.learning_rate * (1 / (1 + self.decay * epoch))
            perm = rng.permutation(self.N)
            if device_epochs:
                loss, acc = self._epoch_on_device(eng, Xd, yd, perm, bs, lrate)
            else:
                losses, correct, seen = [], 0, 0
                starts = list(range(0, self.N, bs))
                for b0 in starts:
                    idx = perm[b0:b0 + bs]
                    feats = X_train[idx]
                    if getattr(self, "input_noise", 0.0):
                        feats = feats + rng.normal(0.0, self.input_noise, size=feats.shape).astype(np.float32)
                    self.step(feats, y_train[idx], lrate, sync=(b0 == starts[-1]))   # the epoch's last step leaves its W in the model
                    losses.append(self._last_loss)
                    correct += int((self._last_probs.argmax(1).cpu().numpy() == y_train[idx]).sum())
                    seen += len(idx)
                loss = float(np.mean([float(l[0]) for l in losses]))
                acc = correct / max(seen, 1)
            val_loss = val_acc = float("nan")
            if X_test is not None and y_test is not None:
                pv = np.asarray(self.model(np.asarray(X_test, np.float32))).reshape(len(X_test), -1)   # model_validate, :139-141
                yt = np.asarray(y_test).reshape(-1)
                val_acc = float((pv.argmax(1) == yt).mean())
                val_loss = float(-np.mean(np.log(np.clip(pv[np.arange(len(yt)), yt], 1e-7, 1 - 1e-7))))
            self.loss_log.append(loss); self.acc_log.append(acc); self.val_log.append((val_loss, val_acc))
            print("Epoch {}, loss: {:.3f}, acc: {:.3f}, val_loss: {:.3f}, val_acc: {:.3f}".format(epoch + 1, loss, acc, val_loss, val_acc))
            if self.robust_linear:  # optimizer.py:131-132 (epochs already holds the extra epoch compile() added, :77-79)
                self.epsilon += self.max_eps / self.epochs
        return self

    def _epoch_on_device(self, eng, Xd, yd, perm, bs, lrate):
        """The mini-batch loop of one epoch (optimizer.py:112-116) as one `dbx_bbb_train_epoch` call; same step ids, same
        kernels and the same order as `step` per mini-batch, so the parameters come out bit-identical."""
        torch = eng.torch
        rt = int(getattr(self, "robust_train", 0))
        epsilon, lam, mode = float(getattr(self, "epsilon", 0.0)), float(getattr(self, "robust_lambda", 1.0)), rt
        if rt == 4:
            self.epsilon = epsilon = max(0.0001, epsilon)       # :125
            mode, lam = 2, 0.0
        name = getattr(self.loss_func, "__name__", type(self.loss_func).__name__).lower() if self.loss_func is not None else "sparse_categorical_crossentropy"
        if "sparse" not in name or "crossentropy" not in name:
            raise NotImplementedError("the device step builds the sparse categorical cross-entropy data term only (got %r)" % name)
        key = (id(self.posterior_mean), id(self.posterior_var))
        if getattr(self, "_train_key", None) != key:
            self._d_mean = self._flat_dev(self.posterior_mean).clone()
            self._d_rho = self._flat_dev(self.posterior_var).clone()
            self._d_pm = self._flat_dev(self.prior_mean).clone()
            self._d_pv = self._flat_dev(self.prior_var).clone()
        pd = torch.from_numpy(np.asarray(perm, np.int64)).to(eng.device)
        xp, yp = Xd.index_select(0, pd).contiguous(), yd.index_select(0, pd).contiguous()
        nb = (len(perm) + bs - 1) // bs
        res = eng.bbb_train_epoch(xp, yp, bs, self._d_mean, self._d_rho, self._d_pm, self._d_pv, lrate, self.kl_weight, self.seed,
                                  self._take_ids(nb), robust_mode=mode, epsilon=epsilon, robust_lambda=lam,
                                  in_lower=self.input_lower, in_upper=self.input_upper, mode=self._train_mode())
        self._last_loss, self._last_probs = res["loss"][-1], res["probs"][(nb - 1) * bs:]
        loss = float(res["loss"][:, 0].mean().item())
        acc = float((res["probs"].argmax(1).to(torch.int32) == yp).float().mean().item())
        self.posterior_mean = self._unflat(self._d_mean.cpu().numpy())
        self.posterior_var = self._unflat(self._d_rho.cpu().numpy())
        self.model.set_weights(self._unflat(res["W"].cpu().numpy()))       # the epoch's last step leaves its W in the model (:59)
        self._train_key = (id(self.posterior_mean), id(self.posterior_var))
        self._dirty = False
        return loss, acc

    def sample(self):
        """bayesbybackprop.py:176-181: N(mean, scale=softplus(rho)) -- the device holds
        sigma = softplus(rho) (dbx_set_gaussian_rho)."""
        self.sync_from_device()
        return super().sample()

    def save(self, path):
        """bayesbybackprop.py:184-195: var.npy = softplus(rho) (a std); no info.pkl."""
        self.sync_from_device()
        var = [softplus(v) for v in self.posterior_var]
        formats.save_gaussian_posterior(path, self.model, self.posterior_mean, var, None)

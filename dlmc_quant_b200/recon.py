"""GPU-resident FSPTQ / RepAPQ block reconstruction - the caller of the per-channel fake-quant and AdaRound
kernels, re-designed from trainer/fsptq_trainer.py:28-112.

Same procedure and hyper-parameters as the reference trainer:
  * targets: every module whose type is in `block_types`, plus FSPTQ layers literally named "conv1" /
    "linear" (fsptq_trainer.py:50-57);
  * for each target, in model order: collect the quantised model's INPUT of the block and the full-precision
    model's OUTPUT of the block over the calibration set, then fit the block's parameters with Adam on
    random 64-sample mini-batches against the l2 loss (fsptq_trainer.py:76-100), per-name learning rates of
    generate_optimizer (fsptq_trainer.py:136-152), cosine schedule over `epochs`.
What changes is the data path:
  * the full-precision outputs of ALL blocks are cached in ONE pass (hooks registered once), instead of
    re-running the fp model for every block (O(L) instead of O(L^2) forward passes);
  * the quantised pass for block k stops at block k (a pre-forward hook raises once the input is captured);
  * caches stay in HBM - no .cpu() / torch.cat / .to(device) round trips (fsptq_trainer.py:39,42,68-72).
"""
import torch

from .scalar.FSPTQuant.base import FSPTQBase

__all__ = ["FSPTQReconstructor", "l2_loss"]


def l2_loss(t1, t2):
    """trainer/loss/loss.py:22-24."""
    return ((t1 - t2) ** 2).sum(axis=1).mean()


class _StopForward(Exception):
    pass


class FSPTQReconstructor:
    def __init__(self, model, fp_model, block_types=(), epochs=0, criterion=l2_loss, minibatch=64, log_every=500,
                 logger=None, scale_lr=None):
        """scale_lr: the reference gives parameters whose name ends in "scales" a learning rate of 1e-3
        (fsptq_trainer.py:143), but its quantizer parameters are called `in_scale` / `wt_scale` - so with the
        reference's own modules that branch never fires and scales train at the default 1e-5.  None (default)
        reproduces that behaviour exactly; a number gives names ending in "scale" or "scales" that learning rate
        (evidently what was meant)."""
        self.scale_lr = scale_lr
        self.model, self.fp_model = model, fp_model
        self.block_types = tuple(block_types)
        self.epochs, self.criterion, self.minibatch = epochs, criterion, minibatch
        self.log_every, self.logger = log_every, logger
        self.history = {}

    # -- which modules are reconstructed (fsptq_trainer.py:44-59) ---------------------------------------
    def targets(self):
        out = []
        for (name, module), fp_module in zip(self.model.named_modules(), self.fp_model.modules()):
            if isinstance(module, FSPTQBase) and name in ("conv1", "linear"):
                out.append((name, module, fp_module))
            elif self.block_types and type(module) in self.block_types:
                out.append((name, module, fp_module))
        return out

    # -- per-name learning rates (fsptq_trainer.py:136-152) -------------------------------------------------
    def generate_optimizer(self, module):
        groups = []
        for name, param in module.named_parameters():
            if name.endswith("weight") or name.endswith("bias"):
                lr = 1e-5
            elif name.endswith("scales"):
                lr = 1e-3
            elif self.scale_lr is not None and name.endswith("scale"):
                lr = self.scale_lr
            elif name.endswith("gamma") or name.endswith("beta"):
                lr = 0.1
            else:
                lr = 1e-5
            groups.append({"params": param, "lr": lr})
        opt = torch.optim.Adam(groups)
        sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=max(self.epochs, 1), eta_min=0.)
        return opt, sched

    # -- caches ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def cache_fp_outputs(self, batches, targets):
        """One pass of the full-precision model: outputs of every target block, kept on the device."""
        cached = {name: [] for name, _, _ in targets}
        handles = [fp.register_forward_hook(lambda m, i, o, n=name: cached[n].append(o.detach()))
                   for name, _, fp in targets]
        self.fp_model.eval()
        for data in batches:
            self.fp_model(data)
        for h in handles:
            h.remove()
        return {n: torch.cat(v) for n, v in cached.items()}

    @torch.no_grad()
    def cache_block_input(self, batches, module):
        """Quantised model up to (not including) `module`: its inputs over the calibration set."""
        got = []

        def pre(m, args):
            got.append(args[0].detach())
            raise _StopForward()

        h = module.register_forward_pre_hook(pre)
        self.model.eval()
        try:
            for data in batches:
                try:
                    self.model(data)
                except _StopForward:
                    pass
        finally:
            h.remove()
        return torch.cat(got)

    # -- the procedure --------------------------------------------------------------------------------------
    def run(self, batches, generator=None):
        """batches: list of device tensors (the calibration set).  Returns {block name: [loss, ...]}."""
        targets = self.targets()
        with torch.no_grad():                 # first pass also triggers every lazy observer (calibration)
            self.model.eval()
            for data in batches:
                self.model(data)
        fp_out = self.cache_fp_outputs(batches, targets)
        for name, module, _ in targets:
            block_input = self.cache_block_input(batches, module)
            block_output = fp_out[name]
            optimizer, scheduler = self.generate_optimizer(module)
            self.model.train()
            losses = []
            for i in range(self.epochs):
                idx = torch.randperm(block_input.size(0), generator=generator)[: self.minibatch].to(block_input.device)
                optimizer.zero_grad()
                loss = self.criterion(block_output[idx], module(block_input[idx]))
                loss.backward()
                optimizer.step()
                scheduler.step()
                if i % self.log_every == 0 or i == self.epochs - 1:
                    losses.append(float(loss.detach()))
                    if self.logger is not None:
                        self.logger.debug("Reconstruction %s iter %d loss %.6f", name, i, losses[-1])
            self.history[name] = losses
        self.model.eval()
        return self.history

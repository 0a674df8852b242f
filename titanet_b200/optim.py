"""The step after the hot path: ``torch.optim.Adam`` as ONE kernel launch.

The reference builds ``torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=0)``
(src/train.py:130-136); on a 426-tensor TitaNet-S that is hundreds of small launches per step.
``FusedAdam`` keeps the same constructor, ``param_groups`` / ``state_dict`` round trip and arithmetic
(bias-corrected, L2 weight decay folded into the gradient, no amsgrad) and updates every parameter in one
``tn_adam_multi`` launch; the step counter and bias corrections live on the device (``tn_adam_tick``), so the
update can be captured in a CUDA graph behind the backward pass.  LR schedulers work unchanged: the learning
rate is read from ``param_groups`` on every ``step()``.
"""
from __future__ import annotations

from typing import List

import torch

from ._lib import TnAdamJob, call, ptr, require_cuda


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("FusedAdam: amsgrad is not supported")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or not 0.0 <= weight_decay:
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))
        self._plans = {}

    def _plan(self, gi: int, group):
        """Device job table + flat moment buffers of one param group (rebuilt when a pointer moves)."""
        ps: List[torch.Tensor] = [p for p in group["params"] if p.grad is not None]
        if not ps:
            return None
        require_cuda(*ps)
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous() or p.grad.dtype != torch.float32:
                raise TypeError("FusedAdam needs contiguous fp32 parameters and gradients")
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in ps)
        plan = self._plans.get(gi)
        if plan is not None and plan["key"] == key:
            return plan
        dev = ps[0].device
        if plan is None or plan["ids"] != [id(p) for p in ps]:
            total = sum(p.numel() for p in ps)
            plan = dict(m=torch.zeros(total, device=dev), v=torch.zeros(total, device=dev), ids=[id(p) for p in ps],
                        hyper=torch.zeros(10, device=dev), hp=None)
            # Resume: moments / step loaded by ``load_state_dict`` (a torch.optim.Adam or FusedAdam checkpoint, e.g. the
            # "optimizer" entry the reference writes, src/learn.py:188-199) move into the flat buffers; the device step
            # counter continues from the loaded step.
            step0, off = 0.0, 0
            for p in ps:
                st = self.state[p]
                n = p.numel()
                if "exp_avg" in st:
                    plan["m"][off:off + n].copy_(st["exp_avg"].reshape(-1))
                if "exp_avg_sq" in st:
                    plan["v"][off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                if "step" in st:
                    step0 = max(step0, float(st["step"]))
                off += n
            plan["step0"] = step0
            off = 0
            for p in ps:          # expose the moments the way torch.optim.Adam does (views of the flat buffers)
                st = self.state[p]
                st["exp_avg"] = plan["m"][off:off + p.numel()].view_as(p)
                st["exp_avg_sq"] = plan["v"][off:off + p.numel()].view_as(p)
                off += p.numel()
        jobs, off = [], 0
        for p in ps:
            n = p.numel()
            jobs.append(TnAdamJob(p.data_ptr(), p.grad.data_ptr(), plan["m"].data_ptr() + 4 * off, plan["v"].data_ptr() + 4 * off, n))
            off += n
        arr = (TnAdamJob * len(jobs))(*jobs)
        plan["jobs"] = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        plan["njobs"], plan["max_n"], plan["key"] = len(jobs), max(p.numel() for p in ps), key
        self._plans[gi] = plan
        return plan

    @staticmethod
    def _hyper_of(group):
        b1, b2 = group["betas"]
        return (float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plan = self._plan(gi, group)
            if plan is None:
                continue
            hp = self._hyper_of(group)
            if plan["hp"] is None:                              # first step of this plan: all ten entries, step included
                lr, b1, b2, eps, wd = hp
                plan["hyper"].copy_(torch.tensor([lr, b1, b2, eps, wd, plan["step0"], 1.0, 1.0, 1.0 - b1, 1.0 - b2]))
                plan["hp"] = hp
            elif plan["hp"] != hp:                              # LR schedulers (and users) write param_groups[i][...]
                lr, b1, b2, eps, wd = hp
                plan["hyper"][0:5].copy_(torch.tensor([lr, b1, b2, eps, wd]), non_blocking=True)
                plan["hyper"][8:10].copy_(torch.tensor([1.0 - b1, 1.0 - b2]), non_blocking=True)
                plan["hp"] = hp
            call("tn_adam_tick", ptr(plan["hyper"]))
            call("tn_adam_multi", ptr(plan["jobs"]), plan["njobs"], plan["max_n"], ptr(plan["hyper"]))
            # the update ran outside autograd's view: bump the version counters so that saved-tensor checks and the
            # per-step weight-split cache (ops.SplitCache) see the parameters as modified
            torch.autograd.graph.increment_version([p for p in group["params"] if p.grad is not None])
        return loss

    def state_dict(self):
        """torch.optim.Adam's layout, ``step`` included (read back from the device counter), so that either optimizer can
        resume from the other's checkpoint."""
        for gi, group in enumerate(self.param_groups):
            plan = self._plans.get(gi)
            if plan is None:
                continue
            step = float(plan["hyper"][5]) if plan["hp"] is not None else plan["step0"]
            for p in group["params"]:
                if p in self.state and "exp_avg" in self.state[p]:
                    self.state[p]["step"] = torch.tensor(step)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._plans = {}                                        # rebuilt (from the loaded moments and step) on the next step

"""Fused flat-buffer momentum SGD with torch.optim.SGD's surface: `param_groups`
(list of dicts whose 'lr' the train loop rewrites every iteration, train.py:136-139), `zero_grad`, `step`,
`state_dict`, `load_state_dict`. One kernel launch updates all parameters (tsb_sgd_flat)."""
import torch

from . import ops
from ._lib import call, ptr, stream
from .flat import flatten, ensure_flat_grads


class SGD(object):
    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0):
        groups = list(params)
        if len(groups) > 0 and not isinstance(groups[0], dict):
            groups = [dict(params=groups)]
        self.defaults = dict(lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.param_groups = []
        for g in groups:
            g = dict(g)
            g["params"] = list(g["params"])
            for k, v in self.defaults.items():
                g.setdefault(k, v)
            self.param_groups.append(g)
        self.momentum = momentum
        self._params = [p for g in self.param_groups for p in g["params"]]
        assert len(set(id(p) for p in self._params)) == len(self._params), "a parameter appears in two groups"
        self.flat_param, self._spans = flatten(self._params, "data")
        self.flat_grad, _ = ensure_flat_grads(self._params)
        self.flat_mom = torch.zeros_like(self.flat_param)
        ends, idx = [], 0
        for g in self.param_groups:
            idx += len(g["params"])
            ends.append(self._spans[idx - 1][1] if idx > 0 else 0)
        dev = self.flat_param.device
        self._seg_end = torch.tensor(ends, dtype=torch.int64, device=dev)
        self._steps = 0
        self.grad_scale = 1.0

    def zero_grad(self, set_to_none=False):
        ensure_flat_grads(self._params)
        self.flat_grad.zero_()
        ops.zero_arena.reset()

    def step(self, closure=None):
        ensure_flat_grads(self._params)
        dev = self.flat_param.device
        lr = torch.tensor([float(g["lr"]) for g in self.param_groups], dtype=torch.float32, device=dev)
        wd = torch.tensor([float(g["weight_decay"]) for g in self.param_groups], dtype=torch.float32, device=dev)
        call("tsb_sgd_flat", ptr(self.flat_param), ptr(self.flat_grad), ptr(self.flat_mom), self.flat_param.numel(),
             ptr(self._seg_end), ptr(lr), ptr(wd), len(self.param_groups), float(self.momentum), float(self.grad_scale),
             1 if self._steps == 0 else 0, stream())
        self._steps += 1
        ops.pack_cache.invalidate()  # bf16 weight packs are stale now

    def state_dict(self):
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        return {"flat_momentum": self.flat_mom.detach().cpu(), "steps": self._steps, "param_groups": groups}

    def load_state_dict(self, sd):
        self.flat_mom.copy_(sd["flat_momentum"])
        self._steps = int(sd["steps"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)

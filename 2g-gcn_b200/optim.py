"""Optional optimiser of the training driver: ``torch.optim.Adam`` (what train.py:40 builds) as ONE kernel launch per step.

``tggcn_backward_ex`` already writes every parameter gradient of a step into one flat buffer (``model.flat_grad``).  ``FlatAdam``
lays the parameters and both Adam moments out the same way — the ``nn.Parameter`` objects keep their identity, names and shapes,
only their storage becomes a view of one buffer — so the whole update is a single pass over four flat arrays
(``tggcn_adam_step``, csrc/optim.cu) instead of torch's ~20 multi-tensor launches (0.84 -> ~0.3 ms at MPHOI hidden 512).
Same arithmetic as ``torch.optim.Adam(params, lr, betas, eps, weight_decay)`` with ``amsgrad=False`` (tests/test_gpu_optim.py),
same ``state_dict()`` layout (``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter), so checkpoints interchange with the
reference's optimiser (train_utils.py:99-112 stores ``optimizer.state_dict()``).  Parameters that never receive a gradient (the
reference's dead ``*_att_mlp`` tensors) are skipped, as torch skips ``grad is None``.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import abi


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0) or weight_decay < 0.0:
            raise ValueError('FlatAdam: invalid hyper-parameter')
        super().__init__(list(model.parameters()), dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self.model = model
        self._layout = None          # (params, offsets, sizes, total) of the flat buffers
        self._p = self._m = self._v = None
        self._steps = 0

    # ------------------------------------------------------------------------------------------------------------------
    def _build(self, layout):
        params, offs, sizes, total = layout
        dev = self.model.flat_grad.device
        old_state = {p: self.state[p] for p in params if p in self.state and 'exp_avg' in self.state[p]}
        flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
        flat_m = torch.zeros(total, dtype=torch.float32, device=dev)
        flat_v = torch.zeros(total, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o, n in zip(params, offs, sizes):
                view = flat_p[o:o + n].view(p.shape)
                view.copy_(p.data)
                p.data = view                                   # same Parameter object, storage inside the flat buffer
                st = old_state.get(p)
                if st is not None:                              # moments loaded from a checkpoint before the first step
                    flat_m[o:o + n].view(p.shape).copy_(st['exp_avg'])
                    flat_v[o:o + n].view(p.shape).copy_(st['exp_avg_sq'])
                    self._steps = max(self._steps, int(st['step']))
        self._p, self._m, self._v, self._layout = flat_p, flat_m, flat_v, layout
        self.model._ptr_cache = None                            # the weight-pointer table of the C ABI must be rebuilt: storages moved
        self._step_t = torch.tensor(float(self._steps))          # one tensor shared by every parameter's state
        for p, o, n in zip(params, offs, sizes):
            self.state[p] = dict(step=self._step_t, exp_avg=flat_m[o:o + n].view(p.shape), exp_avg_sq=flat_v[o:o + n].view(p.shape))

    def _layout_current(self, layout) -> bool:
        if self._layout is None:
            return False
        params, offs, sizes, total = layout
        mine = self._layout
        if total != mine[3] or offs != mine[1] or len(params) != len(mine[0]) or any(a is not b for a, b in zip(params, mine[0])):
            return False
        # a parameter whose storage was moved since (model.to(...), load of a non-flat copy) must be re-adopted
        base = self._p.data_ptr()
        return all(p.data_ptr() == base + 4 * o for p, o in zip(params, offs))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        model = self.model
        layout = getattr(model, '_flat_layout', None)
        if layout is None or model.flat_grad is None:
            raise RuntimeError('FlatAdam.step(): run a forward + backward of the model first (no flat gradient buffer yet)')
        if not self._layout_current(layout):
            self._build(layout)
        params, grads = model._flat_views
        for p, g in zip(params, grads):
            if p.grad is None:
                raise RuntimeError('FlatAdam.step(): a trainable parameter has no gradient (zero_grad() after the backward?)')
            if p.grad.data_ptr() != g.data_ptr():                # gradient accumulation replaced .grad by another tensor
                g.copy_(p.grad)
        group = self.param_groups[0]
        self._steps += 1
        dev = self._p.device
        with torch.cuda.device(dev):
            rc = abi.lib().tggcn_adam_step(self._p.data_ptr(), model.flat_grad.data_ptr(), self._m.data_ptr(), self._v.data_ptr(),
                                           self._p.numel(), float(group['lr']), float(group['betas'][0]), float(group['betas'][1]),
                                           float(group['eps']), float(group['weight_decay']), self._steps,
                                           C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        abi.check(rc, 'tggcn_adam_step')
        self._step_t.fill_(float(self._steps))
        return loss

    def state_dict(self):
        """torch.optim.Adam's layout.  Every parameter gets its OWN step tensor: torch's multi-tensor Adam increments the step
        tensors of all parameters in place, so one shared tensor would be incremented once per parameter after a load."""
        sd = super().state_dict()
        for st in sd['state'].values():
            if 'step' in st:
                st['step'] = torch.tensor(float(self._steps))
        return sd

    def load_state_dict(self, state_dict):
        """torch's loader assigns fresh tensors to the state; the next step() copies them back into the flat buffers."""
        super().load_state_dict(state_dict)
        self._layout = None
        self._steps = 0
        for st in self.state.values():
            if 'step' in st:
                self._steps = max(self._steps, int(st['step']))

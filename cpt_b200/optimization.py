"""Optimizers and schedules of the CPT few-shot scripts, with the update running as ONE native launch.

Reference (the un-vendored pytorch-transformers 1.x at the commit pinned by /root/reference/install.sh:29-32):
`from transformers.pytorch_transformers import AdamW, WarmupLinearSchedule, WarmupConstantSchedule`
(Oscar/oscar/fewshot/gqa_cpt.py:24,342-348, vcr_nsp_cpt.py:28) — same constructor arguments and state keys
('step', 'exp_avg', 'exp_avg_sq').  `AdamW(..., torch_semantics=True)` gives torch.optim.AdamW's update instead
(Oscar/oscar/fewshot/refcoco_cpt.py:342).  Parameters and gradients must be float32 CUDA tensors; there is no CPU path.
"""
import ctypes as C

import numpy as np
import torch
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LambdaLR

from . import _lib

CHUNK = 16384
_TENSOR = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("n", "<i8"), ("lr", "<f4"), ("wd", "<f4"),
                    ("bc1", "<f4"), ("bc2", "<f4")])
_CHUNK = np.dtype([("tensor", "<i4"), ("count", "<i4"), ("offset", "<i8")])


class AdamW(Optimizer):
    """pytorch-transformers 1.x `AdamW(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0,
    correct_bias=True)`: Adam with the weight decay applied to the parameter after the Adam update."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True,
                 torch_semantics=False):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameters: {} - should be in [0.0, 1.0[".format(betas))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        super().__init__(params, defaults)
        self.torch_semantics = bool(torch_semantics)
        self._lib = _lib.load()
        self._chunks = {}  # (numel tuple) -> device chunk table
        self._clip_state = {}  # device -> (scratch, [norm, scale]) of the fused gradient clipping
        self.last_grad_norm = None

    @torch.no_grad()
    def step(self, closure=None, grad_scale=None, max_grad_norm=None):
        """grad_scale: optional fp32 device scalar every gradient is multiplied by inside the update.
        max_grad_norm: fuses `torch.nn.utils.clip_grad_norm_(parameters, max_grad_norm)` (gqa_cpt.py:454) into the
        step: one extra launch reduces the global gradient norm on the device and the update multiplies every gradient
        by min(1, max_norm / (norm + 1e-6)) as it reads it — the gradients are left untouched and nothing
        synchronises.  The norm of the last such step is kept in `self.last_grad_norm` (a device scalar)."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # betas / eps are per launch: one launch per distinct (betas, eps) among the groups (normally one)
        launches = {}
        for group in self.param_groups:
            bkey = (tuple(group["betas"]), float(group["eps"]))
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_cuda:
                    raise RuntimeError("cpt_b200.optimization.AdamW: parameters and gradients must be float32 CUDA "
                                       "tensors (no CPU path)")
                if not p.is_contiguous():
                    raise RuntimeError("cpt_b200.optimization.AdamW: non-contiguous parameter")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                elif not isinstance(state["step"], int):  # a checkpoint of torch.optim.AdamW stores a tensor
                    state["step"] = int(state["step"])
                # one launch per (device, betas, eps): a launch runs on one device with that device's pointers
                launches.setdefault((p.device,) + bkey, []).append((p, group, state))
        if max_grad_norm is not None and launches:
            grad_scale = self._clip_scale([it for items in launches.values() for it in items], float(max_grad_norm),
                                          grad_scale)
        for (_dev, betas, eps), items in launches.items():
            self._launch(items, betas, eps, grad_scale)
            for _p, _g, state in items:  # counted only once the launch has been enqueued
                state["step"] += 1
        return loss

    def _clip_scale(self, items, max_norm, grad_scale):
        devs = {p.device for p, _, _ in items}
        if len(devs) != 1:
            raise RuntimeError("cpt_b200.optimization.AdamW: max_grad_norm needs all parameters on one device")
        dev = items[0][0].device
        tdev, ck, keep = self._tables(items, (0.0, 0.0))
        st = self._clip_state.get(dev)
        if st is None:
            st = (torch.zeros(2, dtype=torch.float64, device=dev), torch.zeros(2, dtype=torch.float32, device=dev))
            self._clip_state[dev] = st
        scratch, out = st
        with torch.cuda.device(dev):
            _lib.check(self._lib.cpt_grad_clip_scale(
                dev.index, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(tdev.data_ptr()),
                C.c_void_p(ck[0].data_ptr()), ck[1], max_norm,
                C.c_void_p(grad_scale.data_ptr()) if grad_scale is not None else C.c_void_p(0),
                C.c_void_p(scratch.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(out.data_ptr() + 4)))
        del keep
        self.last_grad_norm = out[0]
        return out[1:2]

    def _tables(self, items, betas):
        dev = items[0][0].device
        tab = np.zeros(len(items), dtype=_TENSOR)
        keep = []
        for i, (p, group, state) in enumerate(items):
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            keep.append(g)
            t = state["step"] + 1  # the step this launch performs
            correct = group["correct_bias"] or self.torch_semantics
            tab[i] = (p.data_ptr(), g.data_ptr(), state["exp_avg"].data_ptr(), state["exp_avg_sq"].data_ptr(),
                      p.numel(), group["lr"], group["weight_decay"],
                      1.0 - betas[0] ** t if correct else 1.0, 1.0 - betas[1] ** t if correct else 1.0)
        sizes = tuple(int(n) for n in tab["n"])
        ck = self._chunks.get((dev, sizes))
        if ck is None:
            rows = []
            for i, n in enumerate(sizes):
                for off in range(0, n, CHUNK):
                    rows.append((i, min(CHUNK, n - off), off))
            arr = np.array(rows, dtype=_CHUNK) if rows else np.zeros(0, dtype=_CHUNK)
            ck = (torch.from_numpy(arr.view(np.uint8).copy()).to(dev), len(rows))
            self._chunks[(dev, sizes)] = ck
        tdev = torch.from_numpy(tab.view(np.uint8)).pin_memory().to(dev, non_blocking=True)
        return tdev, ck, keep

    def _launch(self, items, betas, eps, grad_scale):
        dev = items[0][0].device
        tdev, ck, keep = self._tables(items, betas)
        with torch.cuda.device(dev):
            _lib.check(self._lib.cpt_adamw_step(
                dev.index, C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(tdev.data_ptr()),
                C.c_void_p(ck[0].data_ptr()), ck[1], betas[0], betas[1], eps, 0 if self.torch_semantics else 1,
                C.c_void_p(grad_scale.data_ptr()) if grad_scale is not None else C.c_void_p(0)))
        # tdev / keep stay referenced until the launch is enqueued; the caching allocator keeps the stream order
        del keep
        # the kernel wrote the parameters behind torch's back: bump their version counters so that whoever caches on
        # them (the engines' "did a weight change?" check that refreshes the 16-bit copies, autograd's saved-tensor
        # checks) sees the update
        torch.autograd.graph.increment_version([p for p, _, _ in items])


class WarmupLinearSchedule(LambdaLR):
    """pytorch-transformers 1.x: linear warm-up over `warmup_steps`, then linear decay to 0 at `t_total`."""

    def __init__(self, optimizer, warmup_steps, t_total, last_epoch=-1):
        self.warmup_steps, self.t_total = warmup_steps, t_total
        super().__init__(optimizer, self.lr_lambda, last_epoch=last_epoch)

    def lr_lambda(self, step):
        if step < self.warmup_steps:
            return float(step) / float(max(1, self.warmup_steps))
        return max(0.0, float(self.t_total - step) / float(max(1.0, self.t_total - self.warmup_steps)))


class WarmupConstantSchedule(LambdaLR):
    """pytorch-transformers 1.x: linear warm-up over `warmup_steps`, then constant."""

    def __init__(self, optimizer, warmup_steps, last_epoch=-1):
        self.warmup_steps = warmup_steps
        super().__init__(optimizer, self.lr_lambda, last_epoch=last_epoch)

    def lr_lambda(self, step):
        if step < self.warmup_steps:
            return float(step) / float(max(1.0, self.warmup_steps))
        return 1.0


def warmup_linear(step, warmup_step, tot_step):
    """Oscar/oscar/utils/optim_sched.py:16-20 (the RefCOCO few-shot schedule)."""
    if step < warmup_step:
        return step / warmup_step
    return max(0, (tot_step - step) / (tot_step - warmup_step))


def get_lr_sched(global_step, opts):
    """Oscar/oscar/utils/optim_sched.py:39-45."""
    lr_this_step = opts.learning_rate * warmup_linear(global_step, opts.warmup_steps, opts.num_train_steps)
    if lr_this_step <= 0:
        lr_this_step = 1e-8
    return lr_this_step

"""Training step for the MSENet regression models: loss, fused AdaBelief, data-parallel wiring.

Mirrors ``BaseModel.optimize_parameters`` of the reference (``torch_points3d/models/base_model.py:230-256``):
forward -> ``0.5 * smooth_l1`` on z-scored targets (``models/instance/base.py:154-179``) -> backward ->
``clip_grad_value_(100)`` -> AdaBelief step (``core/optimizer/adabelief.py:90-201``) -> LR schedule.
The reference's single-process ``nn.DataParallel`` (``trainer.py:149-150``) is replaced by one process per
GPU with a NCCL all-reduce of ONE flat gradient buffer: every plot is an independent
sample, so the only exchange step of the path is that gradient sum.
"""
from __future__ import annotations

import math

import os

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import lib as L


def reg_loss(pred, target, center, scale, weight=0.5):
    """``compute_reg_loss`` (base.py:154-179): smooth-L1 on z-scored targets, NaN targets masked,
    weighted by ``reg_weights.mean()`` = 0.5 (conf/data/instance/NFI/reg.yaml:21-24)."""
    labels = (target - center) / scale
    mask = ~torch.isnan(labels)
    labels = torch.where(mask, labels, pred.detach())          # masked entries contribute zero loss/grad
    denom = mask.sum().clamp(min=1).to(pred.dtype)
    return weight * F.smooth_l1_loss(pred, labels, reduction="sum") / denom


class FlatAdaBelief:
    """AdaBelief over ONE flat fp32 buffer: parameters and gradients of the model are re-pointed to views
    of two contiguous tensors so that the whole update (plus unscale / clip / inf-skip) is a single
    ``b2s_adabelief_step`` launch and the data-parallel exchange is a single all-reduce."""

    def __init__(self, params, lr=5e-3, betas=(0.9, 0.999), eps=1e-16, weight_decay=1e-2, grad_clip=100.0):
        """``params``: an iterable of parameters, or of torch-style group dicts ``{"params": [...], "lr": ...,
        "weight_decay": ..., "betas": ..., "eps": ...}`` -- what ``MinkowskiBaselineModel.get_parameter_list``
        (models/instance/minkowski.py:54-65: head / backbone settings) hands the reference's optimiser.  A group is a
        contiguous segment of the flat buffer with its own hyper-parameters: one kernel launch per group.  ``self.lr``
        is the scheduled learning rate of a group whose own lr equals the constructor's ``lr``; a group with another lr
        keeps its ratio to it, which is how torch's schedulers treat per-group base rates."""
        params = list(params)
        raw_groups = params if (params and isinstance(params[0], dict)) else [{"params": params}]
        self.groups, flat = [], []
        for gdict in raw_groups:
            ps = [p for p in gdict["params"] if p.requires_grad]
            if not ps:
                continue
            self.groups.append({"first": len(flat), "count": len(ps),
                                "lr_ratio": float(gdict.get("lr", lr)) / float(lr),
                                "betas": tuple(gdict.get("betas", betas)), "eps": float(gdict.get("eps", eps)),
                                "weight_decay": float(gdict.get("weight_decay", weight_decay))})
            flat += ps
        self.params = flat
        assert self.params and all(p.is_cuda and p.dtype == torch.float32 for p in self.params)
        assert len({id(p) for p in self.params}) == len(self.params), "a parameter appears in two groups"
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.numel = n
        self.flat_param = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_var = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_param[off:off + k].view_as(p)
                p.grad = self.flat_grad[off:off + k].view_as(p)
                off += k
        self.lr, self.betas, self.eps, self.weight_decay, self.grad_clip = lr, betas, eps, weight_decay, grad_clip
        off = 0
        for g in self.groups:                       # flat-buffer segment of every group
            k = sum(p.numel() for p in self.params[g["first"]:g["first"] + g["count"]])
            g["offset"], g["numel"] = off, k
            off += k
        self.step_count = 0
        # factor applied to every gradient inside the update kernel (before the clip): 1 / world_size turns the SUM the
        # data-parallel exchange leaves in ``flat_grad`` into the mean without a separate pass over the buffer
        self.grad_scale = 1.0
        self.found_inf = torch.zeros(1, dtype=torch.float32, device=dev)
        # device copy of the hyper-parameter block (captured-graph replays read it; see graph_step.py), fed through a
        # ring of pinned staging buffers: the host may run several replays ahead of the device, and a buffer is only
        # rewritten after the copy that read it has completed (event per slot)
        ng = len(self.groups)
        self.hyper_dev = torch.zeros((ng, 16), dtype=torch.float32, device=dev)
        self.hyper_ring = PinnedRing((ng, 16), torch.float32) if torch.cuda.is_available() else None

    def zero_grad(self):
        self.flat_grad.zero_()
        for p in self.params:                 # in-place gradient writes of the next backward may claim them again
            p._b2s_grad_written = False

    @staticmethod
    def rectified_step(step, beta1, beta2):
        """adabelief.py:169-187 -> (num_sma, step_size); degenerated_to_sgd=True."""
        beta2_t = beta2 ** step
        sma_max = 2.0 / (1.0 - beta2) - 1.0
        sma = sma_max - 2.0 * step * beta2_t / (1.0 - beta2_t)
        if sma >= 5:
            step_size = math.sqrt((1 - beta2_t) * (sma - 4) / (sma_max - 4) * (sma - 2) / sma
                                  * sma_max / (sma_max - 2)) / (1 - beta1 ** step)
        else:
            step_size = 1.0 / (1 - beta1 ** step)
        return sma, step_size

    def hyper_values(self, inv_scale=1.0, group=None):
        """The 16-float hyper-parameter block of ``b2s_adabelief_step`` for the CURRENT step count (``group``: one of
        ``self.groups``; default = the constructor's settings)."""
        g = group or {"lr_ratio": 1.0, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}
        b1, b2 = g["betas"]
        sma, step_size = self.rectified_step(self.step_count, b1, b2)
        return [self.lr * g["lr_ratio"], b1, b2, g["eps"], g["weight_decay"], step_size, 1.0 if sma >= 5 else 0.0,
                inv_scale, self.grad_clip, 0, 0, 0, 0, 0, 0, 0]

    def _segments(self):
        for i, g in enumerate(self.groups):
            o, k = g["offset"], g["numel"]
            yield i, g, (self.flat_param[o:o + k], self.flat_grad[o:o + k], self.exp_avg[o:o + k],
                         self.exp_avg_var[o:o + k], k)

    def step(self, inv_scale=1.0, check_inf=False):
        self.step_count += 1
        inv_scale = float(inv_scale) * self.grad_scale
        found = None
        if check_inf:
            L.call("b2s_grad_check", self.flat_grad, self.numel, float(inv_scale), self.found_inf)
            found = self.found_inf
        for _, g, (p, gr, m, v, k) in self._segments():
            L.call("b2s_adabelief_step", p, gr, m, v, k, L.host_f32(*self.hyper_values(inv_scale, g)), None, found)

    def upload_hyper(self, inv_scale=1.0):
        """Advance the step count and copy this step's hyper-parameters to the device (stream-ordered); the
        captured graph's ``step_from_device`` launch reads them."""
        self.step_count += 1
        sc = float(inv_scale) * self.grad_scale
        self.hyper_ring.upload(torch.tensor([self.hyper_values(sc, g) for g in self.groups], dtype=torch.float32),
                               self.hyper_dev)

    def state_dict(self):
        """Optimiser state for checkpoint / resume (the reference saves ``optimizer.state_dict()`` with the model:
        metrics/model_checkpoint.py:182-193): the two moment buffers per parameter, in parameter order, plus the step
        count and the hyper-parameters."""
        state, off = {}, 0
        for i, p in enumerate(self.params):
            k = p.numel()
            state[i] = {"step": self.step_count, "exp_avg": self.exp_avg[off:off + k].view_as(p).clone(),
                        "exp_avg_var": self.exp_avg_var[off:off + k].view_as(p).clone()}
            off += k
        groups = [{"lr": self.lr * g["lr_ratio"], "betas": g["betas"], "eps": g["eps"],
                   "weight_decay": g["weight_decay"], "grad_clip": self.grad_clip,
                   "params": list(range(g["first"], g["first"] + g["count"]))} for g in self.groups]
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        off = 0
        for i, p in enumerate(self.params):
            k = p.numel()
            st = sd["state"].get(i)
            if st is not None:
                self.exp_avg[off:off + k].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_var[off:off + k].copy_(st["exp_avg_var"].reshape(-1))
                self.step_count = int(st["step"])
            off += k
        saved = sd["param_groups"]
        assert len(saved) == len(self.groups), "checkpoint and optimiser disagree on the parameter groups"
        ref = next((i for i, g in enumerate(self.groups) if g["lr_ratio"] == 1.0), 0)
        self.lr = saved[ref]["lr"] / self.groups[ref]["lr_ratio"]
        for g, sg in zip(self.groups, saved):
            g["betas"], g["eps"], g["weight_decay"] = tuple(sg["betas"]), sg["eps"], sg["weight_decay"]
            g["lr_ratio"] = sg["lr"] / self.lr if self.lr else g["lr_ratio"]
        self.betas, self.eps, self.weight_decay = self.groups[ref]["betas"], self.groups[ref]["eps"], self.groups[ref]["weight_decay"]
        self.grad_clip = saved[0].get("grad_clip", self.grad_clip)

    def step_from_device(self, skip_flag=True):
        """AdaBelief update with the hyper-parameters read from ``hyper_dev`` (capturable).  ``found_inf`` (device
        float, nonzero = skip the update) is honoured: the captured step sets it when a capacity check failed."""
        for i, _, (p, gr, m, v, k) in self._segments():
            L.call("b2s_adabelief_step", p, gr, m, v, k, None, self.hyper_dev[i], self.found_inf if skip_flag else None)


class PinnedRing:
    """Small ring of pinned host staging buffers for per-step uploads that are issued with ``non_blocking=True``: a
    slot is rewritten only after the device copy that last read it has completed (one CUDA event per slot), so the
    host can run ahead of the device without a later step's values overtaking an earlier step's copy."""

    def __init__(self, shape, dtype, slots=4):
        self.bufs = [torch.zeros(shape, dtype=dtype).pin_memory() for _ in range(slots)]
        self.events = [None] * slots
        self.i = 0

    def upload(self, host_values, dst):
        i = self.i
        self.i = (i + 1) % len(self.bufs)
        if self.events[i] is not None:
            self.events[i].synchronize()
        self.bufs[i].copy_(host_values.reshape(self.bufs[i].shape))
        dst.copy_(self.bufs[i], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[i] = ev


class CosineAnnealingWarmRestarts:
    """torch's CosineAnnealingWarmRestarts(T_0, T_mult) evaluated at a fractional epoch, as the reference
    steps it per batch (``base_model.py:219-226``; conf/lr_scheduler/cosineawr.yaml:2-5)."""

    def __init__(self, base_lr, T_0=10, T_mult=2, eta_min=0.0):
        self.base_lr, self.T_0, self.T_mult, self.eta_min = base_lr, T_0, T_mult, eta_min

    def lr_at(self, epoch: float) -> float:
        if epoch >= self.T_0 and self.T_mult > 1:
            n = int(math.log(epoch / self.T_0 * (self.T_mult - 1) + 1, self.T_mult))
            t_cur = epoch - self.T_0 * (self.T_mult ** n - 1) / (self.T_mult - 1)
            t_i = self.T_0 * self.T_mult ** n
        elif epoch >= self.T_0:
            t_cur, t_i = epoch % self.T_0, self.T_0
        else:
            t_cur, t_i = epoch, self.T_0
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * t_cur / t_i)) / 2


def shard_plots(weights, rank, world):
    """Data-parallel partition of a global batch of plots: plots are independent samples, so
    each rank takes a disjoint subset.  ``weights`` are per-plot work estimates (point or voxel counts);
    plots are dealt heaviest-first to the currently lightest rank so per-rank work stays within a few %.
    Returns the sorted list of plot indices owned by ``rank``."""
    order = sorted(range(len(weights)), key=lambda i: (-float(weights[i]), i))
    load = [0.0] * world
    owner = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (len(owner[j]) >= -(-len(weights) // world), load[j], j))
        owner[r].append(i)
        load[r] += float(weights[i])
    return sorted(owner[rank])


def allreduce_mean_(flat: torch.Tensor, world: int):
    """The one exchange step of the path: sum a flat gradient buffer over ranks and average in place
    (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    if world > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.mul_(1.0 / world)
    return flat


class Trainer:
    """One optimisation step of MSENet on one GPU (rank); gradients are averaged across ranks."""

    def __init__(self, model, ME, lr=5e-3, weight_decay=1e-2, grad_clip=100.0, batches_per_epoch=133,
                 target_center=(107.0, 200.0), target_scale=(103.0, 194.0)):
        self.model, self.ME = model, ME
        self.opt = FlatAdaBelief(model.parameters(), lr=lr, weight_decay=weight_decay, grad_clip=grad_clip)
        self.sched = CosineAnnealingWarmRestarts(lr)
        self.batches_per_epoch = batches_per_epoch
        dev = self.opt.flat_param.device
        self.center = torch.tensor(target_center, dtype=torch.float32, device=dev)
        self.scale = torch.tensor(target_scale, dtype=torch.float32, device=dev)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.opt.grad_scale = 1.0 / self.world     # the exchange sums; the update kernel applies the mean
        self.num_batches = 0
        self._wgrad_stream = None
        self._img_stream, self._img_specs = None, None
        if os.environ.get("B2S_BRANCH_STREAM", "1") == "1" and self.opt.flat_param.is_cuda:
            from . import msenet as _msenet          # residual downsample branches on their own stream (-0.17 ms / step)
            if _msenet.BRANCH_STREAM is None:
                _msenet.BRANCH_STREAM = torch.cuda.Stream()
        # Gradient exchange overlapped with the backward pass: the parameters of the last stage and the head (85 % of
        # MSENet's weights) sit at the END of the flat buffer and their gradients are complete EARLY in the backward
        # pass -- as soon as the gradient of the last stage's input exists.  A tensor hook there starts the all-reduce
        # of that slice on a second stream; the rest of the buffer follows after the backward pass.
        self.overlap_comm = os.environ.get("B2S_OVERLAP_COMM", "0") == "1"   # opt-in: measured +0.4 % at 2 GPUs
        self.late_offset = self._late_offset() if (self.world > 1 and dev.type == "cuda") else None
        self.comm_stream = torch.cuda.Stream() if self.late_offset is not None else None
        self._late_started = False
        if self.late_offset is not None:
            model.blocks[-1].register_forward_pre_hook(self._watch_late_input)

    def _late_offset(self):
        """Flat-buffer offset of the first parameter of the last stage (everything behind it belongs to that stage or
        to the head), or None when the model does not have that shape."""
        stages = getattr(self.model, "blocks", None)
        if stages is None or len(stages) < 2:
            return None
        late = {id(p) for p in stages[-1].parameters()}
        early = {id(p) for st in stages[:-1] for p in st.parameters()}
        off, first = 0, None
        for p in self.opt.params:
            if first is None and id(p) in late:
                first = off
            elif first is not None and id(p) in early:      # an early-stage parameter behind the split point
                return None
            off += p.numel()
        return first if first else None

    def _watch_late_input(self, module, args):
        x = args[0]
        feats = getattr(x, "F", None)
        if self.overlap_comm and torch.is_tensor(feats) and feats.requires_grad and torch.is_grad_enabled():
            feats.register_hook(self._late_grads_ready)

    def _late_grads_ready(self, grad):
        """Autograd hook on the input of the last stage: every gradient behind ``late_offset`` is final (in-place
        writes of the backward kernels / AccumulateGrad of the head run before this node)."""
        if not self._late_started:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(self.opt.flat_grad[self.late_offset:], op=dist.ReduceOp.SUM)
            self._late_started = True
        return None

    def exchange_gradients(self):
        """SUM the flat gradient buffer over the ranks (finishing what the backward hook started); the division by
        the world size rides in the optimiser kernel (``FlatAdaBelief.grad_scale``), so the exchange is exactly one
        collective and no extra pass over the 57.8 MB buffer."""
        if self.world <= 1:
            return
        g = self.opt.flat_grad
        if self._late_started:
            dist.all_reduce(g[:self.late_offset], op=dist.ReduceOp.SUM)
            torch.cuda.current_stream().wait_stream(self.comm_stream)
            self._late_started = False
        else:
            dist.all_reduce(g, op=dist.ReduceOp.SUM)

    def deferred_counters(self):
        mods = getattr(self.ME, "modules", None)
        if mods is not None and hasattr(mods, "deferred_bn_counters"):
            return mods.deferred_bn_counters()
        import contextlib
        return contextlib.nullcontext()

    def direct_grads(self):
        """Context of the backward pass: parameter gradients written in place into the flat buffer, and (on a GPU, unless
        B2S_SIDE_WGRAD=0) the weight-gradient kernels on a second stream beside the dgrad chain."""
        import contextlib
        fn = getattr(self.ME, "MinkowskiFunctional", None)
        if fn is None or not hasattr(fn, "direct_param_grads"):
            return contextlib.nullcontext()
        stack = contextlib.ExitStack()
        stack.enter_context(fn.direct_param_grads())
        if hasattr(fn, "side_wgrad") and self.opt.flat_param.is_cuda and os.environ.get("B2S_SIDE_WGRAD", "1") == "1":
            if self._wgrad_stream is None:
                self._wgrad_stream = torch.cuda.Stream()
            stack.enter_context(fn.side_wgrad(self._wgrad_stream))
        return stack

    def prepare_weight_images(self):
        """Start building the weight images of every tensor-core convolution for THIS step on a side stream (call after
        the previous optimiser step, before the forward pass); ``join_weight_images`` after the backward pass."""
        fn = getattr(self.ME, "MinkowskiFunctional", None)
        if (fn is None or not hasattr(fn, "prepare_weight_images") or not self.opt.flat_param.is_cuda
                or os.environ.get("B2S_PREBUILT_IMAGES", "1") != "1" or fn.CONV_IMPL == 1 or fn.WORK_STATS is not None):
            return
        if self._img_stream is None:
            self._img_stream = torch.cuda.Stream()
            self._img_specs = fn.conv_image_specs(self.model)
        fn.prepare_weight_images(self._img_specs, self._img_stream)

    def join_weight_images(self):
        if self._img_stream is not None:
            torch.cuda.current_stream().wait_stream(self._img_stream)

    def broadcast_parameters(self):
        if self.world > 1:
            dist.broadcast(self.opt.flat_param, src=0)

    def step(self, coords, feats, target, dense_index=None):
        """coords int32 [N,4] (plot,x,y,z), feats fp32 [N,3], target fp32 [B,2] -- all on this rank's GPU;
        ``dense_index``: the quantiser's ``"index"`` entry when ``coords`` are its unmodified output rows.
        Returns the (detached, on-device) loss."""
        self.model.train()
        self.opt.zero_grad()
        self.prepare_weight_images()
        kw = {"dense_index": dense_index} if dense_index is not None else {}
        x = self.ME.SparseTensor(features=feats, coordinates=coords, **kw)
        with self.deferred_counters():
            pred = self.model(x)
        loss = reg_loss(pred, target, self.center, self.scale)
        with self.direct_grads():            # .grad = views of the flat buffer zeroed above: kernels write in place
            loss.backward()
        self.join_weight_images()
        self.exchange_gradients()            # == DDP's all-reduce: two buckets, the big one overlapped with backward
        # the reference updates with the current lr, THEN steps the scheduler with the un-incremented batch counter
        # (base_model.py:219-226, 246-256): batch 0 and 1 run at base_lr, batch k at lr_at((k - 1) / batches_per_epoch)
        self.opt.step()
        self.opt.lr = self.sched.lr_at(self.num_batches / self.batches_per_epoch)
        self.num_batches += 1
        return loss.detach()

"""One MSENet training step as ONE CUDA graph.

The eager path (``train.Trainer.step``) issues ~350 C-ABI calls per step from Python and synchronises with the
host once per coordinate map (its row count decides the next allocation), so on a B200 the host, not the GPU,
sets the step time.  Here every coordinate map lives at a fixed row CAPACITY, its live row count stays in device
memory (``include/b200sparse.h`` "row counts"), and the whole step -- voxel quantisation -> coordinate hash ->
strided / kernel maps -> forward -> loss -> backward -> [gradient all-reduce] -> AdaBelief -- is captured once
with ``torch.cuda.graph`` and replayed per batch.  Per replay the host only (1) copies the raw points into the
static input buffers, (2) uploads 16 floats of optimiser hyper-parameters and the per-plot drop-path draws,
(3) launches the graph.  Nothing is cached across steps: every replay rebuilds every coordinate structure from
the new points.

Reference call sites: the step is ``BaseModel.optimize_parameters`` (torch_points3d/models/base_model.py:230-256)
fed by ``MinkowskiBaselineModel.set_input`` (models/instance/minkowski.py:67-80) and the ``GridSampling3D``
transform (core/data_transform/grid_transform.py:112-128).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import lib as L
from . import train as T
from .msenet import DropPath
from .quantize import GridSampling3D


def plan_capacities(gs: GridSampling3D, ME, model, batches, num_plots, bounds, margin=0.12, align=128):
    """Row capacities per tensor stride for :class:`GraphStep`: the largest row count seen at every level over
    the sample ``batches`` (dicts with ``pos``/``batch``/``feats``/``perm`` CUDA tensors), plus ``margin``, rounded
    up to ``align``.  Runs the dynamic (exact-size) path without gradients; a few host syncs, done once."""
    seen = {}
    was_training = model.training
    model.eval()
    with torch.no_grad():
        for d in batches:
            vox = gs(d["pos"], d["batch"], tensors=(d["feats"],), order=d.get("perm"), num_plots=num_plots,
                     bounds=bounds)
            x = ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"], dense_index=vox["index"])
            model(x)
            for key, m in x.coordinate_manager.maps.items():
                ts = key.tensor_stride[0]
                if ts > 0:
                    seen[ts] = max(seen.get(ts, 0), m.n)
    model.train(was_training)
    return {ts: max(align, -(-int(n * (1.0 + margin)) // align) * align) for ts, n in seen.items()}


class _LossHandle:
    """Result of ``GraphStep.step_async``: the loss of one step, readable once its device -> host copy has landed.
    At most ``len(ring) - 1`` newer steps may be launched before ``result()`` is called (the slot is then reused)."""

    __slots__ = ("buf", "event")

    def __init__(self, buf, event):
        self.buf, self.event = buf, event

    def result(self) -> float:
        self.event.synchronize()
        return float(self.buf)


class GraphStep:
    """Captured training step.  ``load(host_batch)`` + ``step()`` per batch; ``verify()`` (one host read) checks
    that no capacity was exceeded by the steps since the last call."""

    def __init__(self, trainer: T.Trainer, gs: GridSampling3D, num_plots, n_points, bounds, capacities, feat_dim=3,
                 target_dim=2, capture_collective=True):
        self.tr, self.gs = trainer, gs
        self.B, self.n_points, self.bounds = int(num_plots), int(n_points), bounds
        self.capacities = dict(capacities)
        dev = trainer.opt.flat_param.device
        self.dev = dev
        self.inp = {
            "pos": torch.zeros((n_points, 3), dtype=torch.float32, device=dev),
            "feats": torch.zeros((n_points, feat_dim), dtype=torch.float32, device=dev),
            "batch": torch.zeros(n_points, dtype=torch.int32, device=dev),
            "perm": torch.arange(n_points, dtype=torch.int32, device=dev),
            "target": torch.zeros((num_plots, target_dim), dtype=torch.float32, device=dev),
        }
        self.n_points_dev = torch.full((1,), n_points, dtype=torch.int32, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.drops = [m for m in trainer.model.modules() if isinstance(m, DropPath)]
        nd = max(len(self.drops), 1)
        self.drop_dev = torch.ones((nd, self.B, 1), dtype=torch.float32, device=dev)
        self.drop_ring = T.PinnedRing((nd, self.B, 1), torch.float32)
        for i, m in enumerate(self.drops):
            m.static_mask = self.drop_dev[i]
        self.capture_collective = capture_collective or trainer.world == 1
        if not self.capture_collective:
            trainer.overlap_comm = False      # the exchange stays outside the graph: no collective inside backward
        self.graph = None
        self.graph_tail = None
        self.overlap_maps = True          # replay the map journal of the first warm-up step on a side stream
        self.map_journal = None
        self.side_stream = torch.cuda.Stream()
        self.status = None          # int32 device tensor: live row counts / flags recorded during capture
        self.status_meta = []
        self.launches_per_step = 0  # C-ABI calls recorded into the graph (== kernels-of-ours launches, lower bound)
        self._stage = None          # staging copies of the inputs for the overlapped host feed (prefetch)
        self._loss_ring = None      # pinned slots for the asynchronous loss read-back (step_async)

    # ------------------------------------------------------------------ the step body (captured)
    def _forward_backward(self):
        tr, ME = self.tr, self.tr.ME
        vox = self.gs(self.inp["pos"], self.inp["batch"], tensors=(self.inp["feats"],), order=self.inp["perm"],
                      num_plots=self.B, bounds=self.bounds, capacity=self.capacities[1],
                      n_points_dev=self.n_points_dev)
        tr.opt.zero_grad()
        tr.prepare_weight_images()           # all weight images of this step, on a side stream beside quantiser + stem
        x = ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"], num_rows=vox["num_rows"],
                            capacities=self.capacities, num_batches=self.B, dense_index=vox["index"])
        cm = x.coordinate_manager
        if self.overlap_maps and self.map_journal is not None:
            # build every coordinate structure of the step on a side stream while the stem convolution runs
            cm.prebuild(self.map_journal, self.side_stream)
        with tr.deferred_counters():
            pred = tr.model(x)
        loss = T.reg_loss(pred, self.inp["target"], tr.center, tr.scale)
        with tr.direct_grads():
            loss.backward()
        tr.join_weight_images()
        cm.join_side()
        if self.map_journal is None:
            self.map_journal = list(cm.journal)
        self.loss.copy_(loss.detach())
        cm = x.coordinate_manager
        self.status_meta = [(what, cap) for what, cap, _ in cm.checks]
        status = torch.cat([t.reshape(-1)[:1] for _, _, t in cm.checks])
        if self.status is None:
            self.status = torch.zeros_like(status)
            self.status_min = torch.zeros_like(status)
            self.caps_dev = torch.tensor([cap for _, cap in self.status_meta], dtype=torch.int32, device=status.device)
        # running maximum AND minimum over replays since the last verify(): a negative row count is the quantiser's
        # "point outside the voxel box" flag and must not be folded away by the maximum
        torch.maximum(self.status, status, out=self.status)
        torch.minimum(self.status_min, status, out=self.status_min)
        # a step whose coordinate structures overflowed (or saw no valid batch) must not train: route the condition
        # into the optimiser kernel's skip flag (the GradScaler inf-skip path), evaluated on the device
        bad = (status < 0) | torch.where(self.caps_dev > 0, status > self.caps_dev, status != 0)
        self.tr.opt.found_inf.copy_(bad.any().to(torch.float32).reshape(1))

    def _tail(self):
        self.tr.exchange_gradients()
        self.tr.opt.step_from_device()

    def _body(self):
        self._forward_backward()
        if self.capture_collective:
            self._tail()

    # ------------------------------------------------------------------ capture / replay
    def capture(self, warmup=2):
        """Warm up on a side stream (kernel attributes, allocator, cuBLAS handles), then capture."""
        tr = self.tr
        tr.model.train()
        tr.opt.upload_hyper()                     # valid hyper-parameters for the warm-up steps
        tr.opt.step_count -= 1
        snapshot = [t.clone() for t in (tr.opt.flat_param, tr.opt.exp_avg, tr.opt.exp_avg_var)]
        buffers = [(b, b.clone()) for b in tr.model.buffers()]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._forward_backward()
                self._tail()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        # the warm-up steps must not count as training: restore parameters, optimiser state and BN buffers
        for dst, src in zip((tr.opt.flat_param, tr.opt.exp_avg, tr.opt.exp_avg_var), snapshot):
            dst.copy_(src)
        for b, saved in buffers:
            b.copy_(saved)
        self.status.zero_()
        self.status_min.zero_()
        calls0 = L.launch_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        self.launches_per_step = L.launch_count - calls0
        if not self.capture_collective:
            self.graph_tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_tail):
                self.tr.opt.step_from_device()
        return self

    def load(self, host_batch, non_blocking=True):
        """Copy one collated batch (pinned host tensors or device tensors) into the static input buffers."""
        for k, dst in self.inp.items():
            src = host_batch[k]
            assert src.shape == dst.shape, f"{k}: {tuple(src.shape)} does not match the captured shape {tuple(dst.shape)}"
            dst.copy_(src, non_blocking=non_blocking)

    # ------------------------------------------------------------------ overlapped input feed
    def prefetch(self, host_batch):
        """Start the host -> device copies of the NEXT step's batch (pinned host tensors) on a copy stream, into a
        staging set of device buffers, while the current step computes; ``take_prefetched()`` then moves the staged
        batch into the captured input buffers (a 16 MB device copy, ~5 us) at the start of the next step.  The copy
        stream waits for the previous ``take_prefetched()`` before it overwrites the staging set."""
        if self._stage is None:
            self._stage = {k: torch.empty_like(v) for k, v in self.inp.items()}
            self._copy_stream = torch.cuda.Stream()
            self._staged, self._taken = torch.cuda.Event(), torch.cuda.Event()
            self._taken.record()
        self._copy_stream.wait_event(self._taken)
        with torch.cuda.stream(self._copy_stream):
            for k, dst in self._stage.items():
                src = host_batch[k]
                assert src.shape == dst.shape, f"{k}: {tuple(src.shape)} does not match the captured shape"
                dst.copy_(src, non_blocking=True)
            self._staged.record()

    def take_prefetched(self):
        """Make the current stream wait for the staged batch and copy it into the captured input buffers."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        for k, dst in self.inp.items():
            dst.copy_(self._stage[k], non_blocking=True)
        self._taken.record()

    def step_async(self):
        """``step()`` + an asynchronous device -> host copy of the loss into a pinned slot; returns a handle whose
        ``result()`` blocks until THAT step's loss is on the host.  A loop that reads the previous step's handle after
        launching the current step keeps the device busy while the host prepares the next launch."""
        loss = self.step()
        if self._loss_ring is None:
            self._loss_ring = [(torch.zeros((), dtype=torch.float32).pin_memory(), torch.cuda.Event()) for _ in range(4)]
            self._loss_i = 0
        buf, ev = self._loss_ring[self._loss_i]
        self._loss_i = (self._loss_i + 1) % len(self._loss_ring)
        buf.copy_(loss, non_blocking=True)
        ev.record()
        return _LossHandle(buf, ev)

    def step(self):
        """Replay the captured step on the data currently in the input buffers; returns the on-device loss."""
        tr = self.tr
        tr.opt.upload_hyper()                      # this step runs at the current lr ...
        if self.drops:
            vals = [m.draw(self.B) if tr.model.training else [1.0] * self.B for m in self.drops]
            self.drop_ring.upload(torch.tensor(vals, dtype=torch.float32), self.drop_dev)
        self.graph.replay()
        if not self.capture_collective:
            tr.exchange_gradients()
            self.graph_tail.replay()
        # ... then the scheduler steps with the un-incremented batch counter, as the reference does
        tr.opt.lr = tr.sched.lr_at(tr.num_batches / tr.batches_per_epoch)
        tr.num_batches += 1
        return self.loss

    def release(self):
        """Destroy the captured graphs.  With the gradient all-reduce captured, NCCL keeps the communicator alive
        until every graph that references it is gone: ``dist.destroy_process_group()`` blocks (measured: forever)
        unless this runs first."""
        torch.cuda.synchronize()
        self.graph = None
        self.graph_tail = None
        import gc
        gc.collect()
        torch.cuda.synchronize()

    def verify(self):
        """One host read: raises if any replay since the last call exceeded a row capacity or left the packed
        coordinate range; returns {description: largest value seen}."""
        both = torch.stack([self.status, self.status_min]).tolist()
        vals, lows = both
        self.status.zero_()
        self.status_min.zero_()
        out = {}
        for (what, cap), v, lo in zip(self.status_meta, vals, lows):
            out[what] = v
            if lo < 0:
                raise L.B2SError(f"{what}: the device reported {lo} (a point outside the voxel bounds / an invalid "
                                 f"batch); the optimiser update of that step was skipped")
            if cap == 0:
                if v != 0:
                    raise L.B2SError(f"{what} raised on the device (B2S_EOVERFLOW)")
            elif v > cap:
                raise L.B2SError(f"{what}: {v} exceeds the planned capacity {cap}; re-plan with larger capacities")
        return out


class GraphForward:
    """Captured EVAL forward of a batch of plots: raw points -> quantise -> maps -> network -> prediction, one CUDA
    graph, no host sync (BASELINE.json configs[0] on the GPU: MSENet14 inference on one 16k-point plot; the reference
    runs it as ``eval.py`` -> ``model.forward`` under ``torch.no_grad``, models/base_model.py:153-160).  Same static
    coordinate manager as :class:`GraphStep`; ``verify()`` reports capacity overflows."""

    def __init__(self, model, ME, gs: GridSampling3D, num_plots, n_points, bounds, capacities, feat_dim=3, out_dim=2):
        self.model, self.ME, self.gs = model, ME, gs
        self.B, self.n_points, self.bounds = int(num_plots), int(n_points), bounds
        self.capacities = dict(capacities)
        dev = next(model.parameters()).device
        self.dev = dev
        self.inp = {
            "pos": torch.zeros((n_points, 3), dtype=torch.float32, device=dev),
            "feats": torch.zeros((n_points, feat_dim), dtype=torch.float32, device=dev),
            "batch": torch.zeros(n_points, dtype=torch.int32, device=dev),
            "perm": torch.arange(n_points, dtype=torch.int32, device=dev),
        }
        self.n_points_dev = torch.full((1,), n_points, dtype=torch.int32, device=dev)
        self.pred = torch.zeros((num_plots, out_dim), dtype=torch.float32, device=dev)
        self.pred_host = torch.zeros((num_plots, out_dim), dtype=torch.float32).pin_memory()
        self.graph = None
        self.status = self.status_min = None
        self.status_meta = []
        self.launches_per_step = 0

    def _body(self):
        vox = self.gs(self.inp["pos"], self.inp["batch"], tensors=(self.inp["feats"],), order=self.inp["perm"],
                      num_plots=self.B, bounds=self.bounds, capacity=self.capacities[1],
                      n_points_dev=self.n_points_dev)
        x = self.ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"], num_rows=vox["num_rows"],
                                 capacities=self.capacities, num_batches=self.B, dense_index=vox["index"])
        with torch.no_grad():
            pred = self.model(x)
        self.pred.copy_(pred.reshape(self.pred.shape))
        cm = x.coordinate_manager
        self.status_meta = [(what, cap) for what, cap, _ in cm.checks]
        status = torch.cat([t.reshape(-1)[:1] for _, _, t in cm.checks])
        if self.status is None:
            self.status, self.status_min = torch.zeros_like(status), torch.zeros_like(status)
        torch.maximum(self.status, status, out=self.status)
        torch.minimum(self.status_min, status, out=self.status_min)

    def capture(self, warmup=2):
        self.model.eval()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._body()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.status.zero_()
        self.status_min.zero_()
        calls0 = L.launch_count
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        self.launches_per_step = L.launch_count - calls0
        return self

    def load(self, host_batch, non_blocking=True):
        for k, dst in self.inp.items():
            dst.copy_(host_batch[k], non_blocking=non_blocking)

    def step(self):
        """Replay on the loaded points; returns the on-device prediction ``[B, out_dim]``."""
        self.graph.replay()
        return self.pred

    def verify(self):
        vals, lows = torch.stack([self.status, self.status_min]).tolist()
        self.status.zero_()
        self.status_min.zero_()
        out = {}
        for (what, cap), v, lo in zip(self.status_meta, vals, lows):
            out[what] = v
            if lo < 0:
                raise L.B2SError(f"{what}: the device reported {lo} (a point outside the voxel bounds)")
            if (cap == 0 and v != 0) or (cap > 0 and v > cap):
                raise L.B2SError(f"{what}: {v} against capacity {cap} (0 = flag)")
        return out

    def release(self):
        torch.cuda.synchronize()
        self.graph = None


class _Slot:
    """One of the two buffer sets of :class:`PipelinedGraphStep`: input buffers, the coordinate structures built from
    them, the two captured graphs and the events that order them."""

    def __init__(self, proto, n_points):
        self.inp = {k: torch.zeros_like(v) for k, v in proto.items()}
        self.inp["perm"] = proto["perm"].clone()
        self.n_points_dev = torch.full((1,), n_points, dtype=torch.int32, device=proto["pos"].device)
        self.x = None
        self.g_pre = self.g_main = None
        self.prep_done, self.train_done = torch.cuda.Event(), torch.cuda.Event()


class PipelinedGraphStep(GraphStep):
    """The captured step split in two graphs per buffer set and software-pipelined ACROSS steps: the coordinate
    pipeline of batch i+1 (voxel quantisation, coordinate hash, every strided / kernel map, the stem's x-line table --
    everything that depends on the points only) replays on a second stream while batch i trains (forward, backward,
    exchange, optimiser).  Nothing is cached across steps -- every batch still builds all of its structures from its
    own points; they are built one step early, into the other of two buffer sets.

        step.feed(batch)                     # host / device -> input buffers of the next set + its coordinate graph
        for nxt in batches: step.feed(nxt); loss = step.step()      # train the set fed one call earlier

    Same kernels, same operands, same results as :class:`GraphStep` (tests/test_gpu_graph.py)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.slots = [_Slot(self.inp, self.n_points), _Slot(self.inp, self.n_points)]
        self.prep_stream = torch.cuda.Stream()
        self._fed, self._cur = 0, 0          # slot the next feed() fills / the next step() trains
        self._outstanding = 0                # batches fed and not yet trained (0, 1 or 2)

    # ---- the two halves of GraphStep._forward_backward
    def _prep(self, slot):
        tr, ME = self.tr, self.tr.ME
        i = slot.inp
        vox = self.gs(i["pos"], i["batch"], tensors=(i["feats"],), order=i["perm"], num_plots=self.B,
                      bounds=self.bounds, capacity=self.capacities[1], n_points_dev=slot.n_points_dev)
        x = ME.SparseTensor(features=vox["tensors"][0], coordinates=vox["coords"], num_rows=vox["num_rows"],
                            capacities=self.capacities, num_batches=self.B, dense_index=vox["index"])
        if self.map_journal is not None:
            x.coordinate_manager.build_all(self.map_journal)
        slot.x, slot.vox = x, vox

    def _train(self, slot):
        tr = self.tr
        tr.opt.zero_grad()
        tr.prepare_weight_images()
        x = slot.x
        cm = x.coordinate_manager
        with tr.deferred_counters():
            pred = tr.model(x)
        loss = T.reg_loss(pred, slot.inp["target"], tr.center, tr.scale)
        with tr.direct_grads():
            loss.backward()
        tr.join_weight_images()
        if self.map_journal is None:
            self.map_journal = list(cm.journal)
        cm._side, cm._side_events, cm._built = None, {}, {}
        self.loss.copy_(loss.detach())
        self.status_meta = [(what, cap) for what, cap, _ in cm.checks]
        status = torch.cat([t.reshape(-1)[:1] for _, _, t in cm.checks])
        if self.status is None:
            self.status = torch.zeros_like(status)
            self.status_min = torch.zeros_like(status)
            self.caps_dev = torch.tensor([cap for _, cap in self.status_meta], dtype=torch.int32, device=status.device)
        torch.maximum(self.status, status, out=self.status)
        torch.minimum(self.status_min, status, out=self.status_min)
        bad = (status < 0) | torch.where(self.caps_dev > 0, status > self.caps_dev, status != 0)
        tr.opt.found_inf.copy_(bad.any().to(torch.float32).reshape(1))
        if self.capture_collective:
            self._tail()

    def capture(self, warmup=2):
        tr = self.tr
        tr.model.train()
        tr.opt.upload_hyper()
        tr.opt.step_count -= 1
        snapshot = [t.clone() for t in (tr.opt.flat_param, tr.opt.exp_avg, tr.opt.exp_avg_var)]
        buffers = [(b, b.clone()) for b in tr.model.buffers()]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                for slot in self.slots:
                    self._prep(slot)
                    self._train(slot)
                    if not self.capture_collective:
                        self._tail()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for dst, src in zip((tr.opt.flat_param, tr.opt.exp_avg, tr.opt.exp_avg_var), snapshot):
            dst.copy_(src)
        for b, saved in buffers:
            b.copy_(saved)
        self.status.zero_()
        self.status_min.zero_()
        calls0 = L.launch_count
        for slot in self.slots:                      # the coordinate graphs first: the training graphs read their tensors
            slot.g_pre = torch.cuda.CUDAGraph()
            with torch.cuda.graph(slot.g_pre):
                self._prep(slot)
        for slot in self.slots:
            slot.g_main = torch.cuda.CUDAGraph()
            with torch.cuda.graph(slot.g_main):
                self._train(slot)
        self.launches_per_step = (L.launch_count - calls0) // 2
        if not self.capture_collective:
            self.graph_tail = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_tail):
                self.tr.opt.step_from_device()
        torch.cuda.synchronize()
        for slot in self.slots:
            slot.train_done.record()
        self.graph = True                            # truthy: captured
        return self

    # ---- replay
    def feed(self, batch, non_blocking=True):
        """Copy ``batch`` (pinned host or device tensors) into the next buffer set and replay its coordinate graph on
        the prep stream -- beside whatever the main stream is training."""
        if self._outstanding >= 2:
            raise RuntimeError("PipelinedGraphStep.feed(): both buffer sets hold a batch that has not been trained; "
                               "call step() first")
        slot = self.slots[self._fed]
        self._fed ^= 1
        self._outstanding += 1
        ps = self.prep_stream
        ps.wait_event(slot.train_done)               # the step that last trained on this set has finished with it
        with torch.cuda.stream(ps):
            for k, dst in slot.inp.items():
                src = batch[k]
                assert src.shape == dst.shape, f"{k}: {tuple(src.shape)} does not match the captured shape {tuple(dst.shape)}"
                dst.copy_(src, non_blocking=non_blocking)
            slot.g_pre.replay()
            slot.prep_done.record(ps)

    def load(self, host_batch, non_blocking=True):
        self.feed(host_batch, non_blocking)

    def reset_feed(self):
        """Forget a batch that was fed but not trained (end of a loop that feeds one ahead)."""
        self._fed = self._cur
        self._outstanding = 0

    def step(self):
        """Train the set fed by the oldest outstanding ``feed``; returns the on-device loss."""
        tr = self.tr
        if self._outstanding == 0:
            raise RuntimeError("PipelinedGraphStep.step(): no batch has been fed (feed() one batch ahead of step())")
        self._outstanding -= 1
        slot = self.slots[self._cur]
        self._cur ^= 1
        cur = torch.cuda.current_stream()
        cur.wait_event(slot.prep_done)
        tr.opt.upload_hyper()
        if self.drops:
            vals = [m.draw(self.B) if tr.model.training else [1.0] * self.B for m in self.drops]
            self.drop_ring.upload(torch.tensor(vals, dtype=torch.float32), self.drop_dev)
        slot.g_main.replay()
        if not self.capture_collective:
            tr.exchange_gradients()
            self.graph_tail.replay()
        slot.train_done.record(cur)
        tr.opt.lr = tr.sched.lr_at(tr.num_batches / tr.batches_per_epoch)
        tr.num_batches += 1
        return self.loss

    def release(self):
        torch.cuda.synchronize()
        for slot in self.slots:
            slot.g_pre = slot.g_main = None
        self.graph_tail = None
        import gc
        gc.collect()
        torch.cuda.synchronize()

"""Coordinate manager: GPU coordinate hash, strided maps, kernel maps, per-plot row info.

Two modes.  DYNAMIC (default): every map is sized exactly; creating a map costs one host sync (its row count).
STATIC (``capacities=...``): every map is allocated at a fixed row capacity per tensor stride and its actual row
count stays on the device (``CoordMap.n_dev``); no call synchronises, so a whole training step can be captured
into one CUDA graph (``dpcr_agb_b200.graph_step``).  The kernels take both (capacity, device count), see
``include/b200sparse.h`` "row counts".

Mirrors what ``ME.SparseTensor(...)`` / strided ops create inside MinkowskiEngine's
``CoordinateManager`` (call sites: ``torch_points3d/models/instance/minkowski.py:74``,
``modules/MinkowskiEngine/SENet.py:53,94-97``, ``resnet_block.py:48-54`` of the reference).
All state lives in torch CUDA tensors owned by this object; every kernel is a C-ABI call
(``include/b200sparse.h``).  Host syncs: one per new coordinate map (its row count).
"""
from __future__ import annotations

import os

import torch

from dpcr_agb_b200 import lib as L


USE_DENSE_INDEX = True   # tests flip this to compare the occupancy-index kernel map with the hash one
USE_LINES = True         # tests flip this to compare the x-line convolution kernels with the table-driven ones

# bench.py sets this to a dict to collect the ALGORITHMIC bytes of the integer stages per entry point (SURVEY.md 8d:
# kernel map 16 N_out + 8 P, strided map 16 N_in + 16 N_out + 4 N_in); pair counts cost a host sync per map, so this
# is only ever enabled in an untimed statistics pass
MAP_STATS = None

# prebuild(): consumers wait for the side-stream operation they need instead of for the whole side stream
PER_OP_JOIN = os.environ.get("B2S_PER_OP_JOIN", "1") == "1"


def _map_account(name, nbytes):
    if MAP_STATS is not None:
        st = MAP_STATS.setdefault(name, {"calls": 0, "bytes": 0})
        st["calls"] += 1
        st["bytes"] += int(nbytes)


def _triple(v):
    if isinstance(v, torch.Tensor):
        v = v.tolist()
    if isinstance(v, (list, tuple)):
        if len(v) == 1:
            return (int(v[0]),) * 3
        assert len(v) == 3, "only D=3 is supported"
        return tuple(int(a) for a in v)
    return (int(v),) * 3


class CoordinateMapKey:
    """(tensor_stride, tag) -- hashable identity of a coordinate map inside one manager."""

    __slots__ = ("tensor_stride", "tag")

    def __init__(self, tensor_stride, tag=""):
        self.tensor_stride = _triple(tensor_stride)
        self.tag = tag

    def get_tensor_stride(self):
        return list(self.tensor_stride)

    def get_key(self):
        return (list(self.tensor_stride), self.tag)

    def __eq__(self, other):
        return isinstance(other, CoordinateMapKey) and self.tensor_stride == other.tensor_stride \
            and self.tag == other.tag

    def __hash__(self):
        return hash((self.tensor_stride, self.tag))

    def __repr__(self):
        return f"CoordinateMapKey(tensor_stride={list(self.tensor_stride)}, tag={self.tag!r})"


class CoordMap:
    """One coordinate map: rows ``int32 [N,4]`` (batch,x,y,z) + its open-addressing hash table."""

    __slots__ = ("coords", "table", "capacity", "n", "n_dev", "info", "dense", "_inv_counts", "_counts_host")

    def __init__(self, coords, table, capacity, n_dev=None, info=None):
        self.coords = coords
        self.table = table
        self.capacity = capacity
        self.n = coords.shape[0]          # rows allocated: the exact count (dynamic) or the capacity (static)
        self.n_dev = n_dev                # int32 [1] device tensor with the live row count, or None
        self.info = info                  # int32 [4] device tensor of b2s_coordmap_insert (static mode checks)
        self.dense = None                 # (workspace, lo, dims, num_plots): the quantiser's occupancy index, if the
        #                                   rows of this map are exactly the quantiser's output rows
        self._inv_counts = None
        self._counts_host = None


class KernelMap:
    """Neighbour table ``nbr int32 [K^3, N_out]`` (+ lazily the transposed table for dgrad)."""

    def __init__(self, manager, in_key, out_key, kernel_size, step, nbr, n_in, n_out, n_in_dev=None, n_out_dev=None):
        self.manager, self.in_key, self.out_key = manager, in_key, out_key
        self.kernel_size, self.step = kernel_size, step
        self._nbr, self.n_in, self.n_out = nbr, n_in, n_out
        self.n_in_dev, self.n_out_dev = n_in_dev, n_out_dev
        self.k3 = kernel_size[0] * kernel_size[1] * kernel_size[2]
        # stride-1 odd kernels are point-symmetric: the transposed table is nbr with k reversed
        self.symmetric = in_key == out_key and all(k % 2 == 1 for k in kernel_size)
        self._inv = None
        self._lines = None

    @property
    def nbr(self):
        """Neighbour table ``[K^3, N_out]``.  Maps that have the x-line form (``lines_ok``) build it on first use only:
        the k7 stem never asks for it."""
        self.manager._sync_builds()
        if self._nbr is None:
            cm = self.manager
            self._nbr = cm._probe(cm.maps[self.out_key], cm.maps[self.in_key], self.kernel_size, self.step, +1)
            cm._mark_built()
        return self._nbr

    @nbr.setter
    def nbr(self, value):
        self._nbr = value

    @property
    def lines_ok(self):
        """The x-line form exists: a stride-1 map of the quantiser's own rows (sorted by cell), at most 8 offsets along
        x with an x step of 1, fewer than 2^24 rows (C ABI ``b2s_kernel_map_lines``)."""
        m = self.manager.maps[self.in_key]
        return (USE_LINES and USE_DENSE_INDEX and self.in_key == self.out_key and m.dense is not None
                and self.kernel_size[0] <= 8 and self.step[0] == 1 and m.n < (1 << 24))

    @property
    def lines(self):
        """uint32-in-int32 ``[K1*K2, N]`` line words ``(base << 8) | mask`` (include/b200sparse.h), built on first use."""
        self.manager._sync_builds()
        if self._lines is None:
            assert self.lines_ok
            cm = self.manager
            cm._note(("lines", self.in_key, self.out_key, self.kernel_size))
            m = cm.maps[self.in_key]
            ws, lo, dims, num_plots = m.dense
            nl = self.kernel_size[1] * self.kernel_size[2]
            self._lines = torch.empty((nl, m.n), dtype=torch.int32, device=m.coords.device)
            L.call("b2s_kernel_map_lines", m.coords, m.n, m.n_dev, ws, num_plots, L.host_i32(*lo), L.host_i32(*dims),
                   L.host_i32(*self.kernel_size), L.host_i32(*self.step), self._lines)
            if MAP_STATS is not None:
                _map_account("b2s_kernel_map_lines", 16 * m.n + 8 * self.num_pairs())
            cm._mark_built()
        return self._lines

    def num_pairs(self, n_rows=None):
        """Number of (in, out) pairs of the map (host sync; statistics only)."""
        n_rows = self.n_out if n_rows is None else n_rows
        if self._nbr is None and self._lines is not None:
            masks = (self._lines[:, :n_rows] & 0xFF).to(torch.uint8)
            bits = torch.zeros_like(masks, dtype=torch.int64)
            for b in range(8):
                bits += (masks >> b) & 1
            return int(bits.sum().item())
        return int((self.nbr[:, :n_rows] >= 0).sum().item())

    @property
    def inv(self):
        """Transposed table ``[K^3, N_in]``: ``inv[k, i] = o`` iff ``nbr[k, o] = i`` (strided maps only)."""
        self.manager._join(("inv", self.in_key, self.out_key, self.kernel_size))
        if self._inv is None:
            cm = self.manager
            cm._note(("inv", self.in_key, self.out_key, self.kernel_size))
            self._inv = cm._probe(cm.maps[self.in_key], cm.maps[self.out_key], self.kernel_size, self.step, -1)
            cm._mark_built()
        return self._inv

    @property
    def parity_plan(self):
        """(perm, bounds) of ``b2s_parity_plan`` for the fine (input) rows of a stride-2 map, or None when the map is
        not a stride-2 map with kernel sizes 1 or 3 (dgrad then takes the dense transposed table)."""
        self.manager._join(("plan", self.in_key, self.out_key, self.kernel_size))
        if getattr(self, "_plan", None) is None:
            cm = self.manager
            cm._note(("plan", self.in_key, self.out_key, self.kernel_size))
            tin, tout = self.in_key.tensor_stride, self.out_key.tensor_stride
            ok = all(o == 2 * i for i, o in zip(tin, tout)) and all(k in (1, 3) for k in self.kernel_size) \
                and all(s == t for s, t in zip(self.step, tin))
            if not ok:
                self._plan = False
            elif (self.in_key, self.out_key) in cm.parity_plans:
                # the plan sorts the fine rows by their position in the coarse cell: it depends on the two maps only,
                # so the k3 convolution and the k1 downsample convolution of a block share it
                self._plan = cm.parity_plans[(self.in_key, self.out_key)]
            else:
                imap = cm.maps[self.in_key]
                rows = L.query("b2s_parity_plan_rows", imap.n)
                dev = imap.coords.device
                perm = torch.empty(rows, dtype=torch.int32, device=dev)
                bounds = torch.empty(9, dtype=torch.int32, device=dev)
                scratch = torch.empty(16, dtype=torch.int32, device=dev)
                L.call("b2s_parity_plan", imap.coords, imap.n, imap.n_dev, L.host_i32(*tout), perm, bounds, scratch)
                self._plan = (perm, bounds)
                cm.parity_plans[(self.in_key, self.out_key)] = self._plan
                cm._mark_built()
        return self._plan or None

    def transposed(self):
        """The same pairs seen from the other side: a KernelMap whose "out" rows are this map's in rows -- the map of
        ``MinkowskiConvolutionTranspose`` / ``MinkowskiPoolingTranspose`` from the coarse rows back onto the fine map
        (MinkowskiEngine builds the forward kernel map fine -> coarse and swaps its sides).  ``nbr`` of the view is this
        map's transposed table, and vice versa; the offset index k keeps its meaning."""
        if getattr(self, "_tview", None) is None:
            tv = KernelMap.__new__(KernelMap)
            tv.manager, tv.in_key, tv.out_key = self.manager, self.out_key, self.in_key
            tv.kernel_size, tv.step, tv.k3 = self.kernel_size, self.step, self.k3
            tv.n_in, tv.n_out, tv.n_in_dev, tv.n_out_dev = self.n_out, self.n_in, self.n_out_dev, self.n_in_dev
            tv._nbr, tv._inv, tv._lines = self.inv, self.nbr, None
            tv.symmetric = False
            tv._plan = False                 # no parity plan: dgrad of the view walks this map's forward table
            tv._tview = self
            self._tview = tv
        return self._tview

    def pairs(self):
        """MinkowskiEngine's pair-list form: (in_idx, out_idx, offsets[K^3+1]); sorted by out row per offset."""
        assert self.n_out_dev is None, "pairs() needs exact row counts (dynamic mode)"
        counts = torch.empty(self.k3, dtype=torch.int32, device=self.nbr.device)
        L.call("b2s_kernel_map_pair_counts", self.nbr, self.k3, self.n_out, counts)
        offsets = torch.zeros(self.k3 + 1, dtype=torch.int64, device=self.nbr.device)
        offsets[1:] = torch.cumsum(counts.long(), 0)
        total = int(offsets[-1].item())
        in_idx = torch.empty(total, dtype=torch.int32, device=self.nbr.device)
        out_idx = torch.empty(total, dtype=torch.int32, device=self.nbr.device)
        if total:
            L.call("b2s_kernel_map_pairs_fill", self.nbr, self.k3, self.n_out, offsets, in_idx, out_idx)
        return in_idx, out_idx, offsets


class CoordinateManager:
    def __init__(self, D=3, device=None, capacities=None, num_batches=None):
        """``capacities``: {tensor_stride (int, the x stride): row capacity} switches the manager to STATIC mode;
        ``num_batches`` must then be given too (it cannot be read back without a sync)."""
        assert D == 3, "the B200 path implements the D=3 case the reference uses"
        self.D = D
        self.device = device
        self.maps = {}
        self.kernel_maps = {}
        self.parity_plans = {}
        # Map journal: every map-building operation of a step in the order it happened.  A later step with the same
        # network can replay it up front on a side stream (prebuild), so that the integer pipeline -- hash inserts,
        # kernel maps, transposed tables, parity plans, per-plot counts -- overlaps the stem convolution instead of
        # sitting between the layers.  _side: the stream whose work the first consumer of a prebuilt map must join.
        self.journal = []
        self._side = None
        self._side_events = {}      # prebuild: event recorded behind every side-stream operation, by _join key
        self._built = {}            # stream id -> event behind the latest map build on that stream
        self._replaying = False
        self._main_built = set()
        self.capacities = dict(capacities) if capacities is not None else None
        self.static = capacities is not None
        if self.static:
            assert num_batches is not None, "static mode needs num_batches"
        self.num_batches = int(num_batches) if num_batches is not None else 0
        self.checks = []     # static mode: (description, capacity, int32 device tensor [>=1]) to verify after a replay

    # ------------------------------------------------------------------ journal / prebuild
    def _note(self, op):
        if not self._replaying:
            self.journal.append(op)

    # ------------------------------------------------------------------ builds on more than one stream
    # A network may run a branch on its own stream (dpcr_agb_b200.msenet.BRANCH_STREAM); whichever stream asks first
    # builds a map that the other one then finds in the cache.  Every build is followed by an event on its stream
    # (_mark_built) and every accessor first makes its stream wait for the builds of the OTHER streams (_sync_builds).
    def _mark_built(self):
        if self._replaying:              # prebuild(): the per-operation events of the side stream cover these
            return
        s = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(s)
        self._built[s.cuda_stream] = ev

    def _sync_builds(self):
        if not self._built:
            return
        cur = torch.cuda.current_stream()
        for sid, ev in self._built.items():
            if sid != cur.cuda_stream:
                cur.wait_event(ev)

    def _join(self, what):
        """Called by every map accessor: a request for something the side stream built makes the caller's stream wait
        for THAT operation (an event recorded behind it on the side stream; the journal is replayed in order of first
        use, so the first consumer waits for the first few operations only, not for every map of the step); anything
        else that was not built on the caller's stream waits for the whole side stream."""
        self._sync_builds()
        if self._side is None or self._replaying or what in self._main_built:
            return
        ev = self._side_events.get(what)
        if ev is not None:                 # kept: a consumer on another stream (a branch of the network) waits again
            torch.cuda.current_stream().wait_event(ev)
            return
        torch.cuda.current_stream().wait_stream(self._side)

    @staticmethod
    def _what_of(op):
        """The ``_join`` key of a journal operation."""
        kind = op[0]
        if kind == "stride":
            ts = tuple(a * b for a, b in zip(op[1].tensor_stride, _triple(op[2])))
            return ("map", CoordinateMapKey(ts, op[1].tag))
        if kind == "kmap":
            return ("kmap", (op[1], op[2], op[3], op[4]))
        if kind in ("inv", "plan", "lines"):
            return (kind, op[1], op[2], op[3])
        return (kind, op[1])

    def _replay(self, op):
        kind = op[0]
        if kind == "stride":
            self.stride(op[1], op[2])
        elif kind == "kmap":
            self.kernel_map(op[1], op[2], op[3], op[4])
        elif kind in ("inv", "plan", "lines"):
            km = next((k for k in self.kernel_maps.values()
                       if (k.in_key, k.out_key, k.kernel_size) == (op[1], op[2], op[3])), None)
            if km is not None:
                _ = km.inv if kind == "inv" else (km.parity_plan if kind == "plan" else km.lines)
        elif kind == "invc":
            self.inv_counts(op[1])

    def prebuild(self, journal, side_stream, main_first=1):
        """Replay a previous step's map journal: the first ``main_first`` operations (the stem's kernel map, which the
        first convolution needs at once) on the current stream, the rest on ``side_stream`` forked from it here and
        joined by the first consumer (``_join``).  Results are identical to building on demand -- same kernels, same
        inputs -- only their place in the stream order changes."""
        self._replaying = True
        try:
            for op in journal[:main_first]:
                self._replay(op)
                if op[0] == "kmap":
                    self._main_built.add(("kmap", (op[1], op[2], op[3], op[4])))
                elif op[0] == "stride":
                    ts = tuple(a * b for a, b in zip(op[1].tensor_stride, _triple(op[2])))
                    self._main_built.add(("map", CoordinateMapKey(ts, op[1].tag)))
            main = torch.cuda.current_stream()
            side_stream.wait_stream(main)
            with torch.cuda.stream(side_stream):
                for op in journal[main_first:]:
                    self._replay(op)
                    if PER_OP_JOIN:
                        ev = torch.cuda.Event()
                        ev.record(side_stream)
                        what = self._what_of(op)
                        self._side_events[what] = ev
        finally:
            self._replaying = False
        self._side = side_stream

    def build_all(self, journal):
        """Replay a previous step's whole map journal on the CURRENT stream and leave no stream bookkeeping behind: the
        form a separately captured "coordinate pipeline" graph uses (graph_step.PipelinedGraphStep) -- its consumers
        run in another graph, ordered behind it by an event between the two replays."""
        self._replaying = True
        try:
            for op in journal:
                self._replay(op)
        finally:
            self._replaying = False
        self._side, self._side_events, self._built = None, {}, {}

    def join_side(self):
        """Make the current stream wait for any outstanding prebuild work (end of a step)."""
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_events.clear()
            self._side = None

    def capacity_of(self, tensor_stride):
        ts = tensor_stride[0]
        if ts not in self.capacities:
            raise L.B2SError(f"no row capacity planned for tensor stride {ts} (have {sorted(self.capacities)})")
        return int(self.capacities[ts])

    def insert_static(self, coords: torch.Tensor, n_dev: torch.Tensor, tensor_stride=(1, 1, 1), tag="",
                      dense_index=None):
        """STATIC mode: register ``coords`` int32 [capacity, 4] whose first ``n_dev[0]`` rows are live, unique and
        batch-sorted (the quantiser's output contract) -- hash build only, no host sync."""
        assert self.static and coords.dtype == torch.int32 and coords.is_contiguous() and coords.shape[1] == 4
        cap_rows = coords.shape[0]
        dev = coords.device
        key = CoordinateMapKey(tensor_stride, tag)
        self.device = dev
        self.checks.append((f"rows at tensor stride {key.tensor_stride[0]}", cap_rows, n_dev))
        if dense_index is not None and USE_DENSE_INDEX:
            # every kernel map over these rows goes through the quantiser's occupancy index and the strided map
            # below builds its own table: the hash of the finest map (the largest of the step) would never be probed.
            # It is built on first use (_table_of); the coordinate range is bounded by the quantiser's box.
            cmap = CoordMap(coords, None, 0, n_dev=n_dev, info=None)
            cmap.dense = dense_index
            self.maps[key] = cmap
            return key
        cmap = CoordMap(coords, None, 0, n_dev=n_dev, info=None)
        self._table_of(cmap)
        cmap.dense = dense_index
        self.maps[key] = cmap
        return key

    def _table_of(self, cmap: CoordMap):
        """Hash table of a map whose rows are unique (table values == rows); built on first use."""
        if cmap.table is None:
            cap_rows, dev = cmap.coords.shape[0], cmap.coords.device
            hcap = L.query("b2s_hash_capacity", cap_rows)
            table = torch.empty(hcap * 16, dtype=torch.uint8, device=dev)
            slot = torch.empty(max(cap_rows, 1), dtype=torch.int32, device=dev)
            rank = torch.empty(max(cap_rows, 1), dtype=torch.int32, device=dev)
            info = torch.empty(4, dtype=torch.int32, device=dev)
            scan_ws = torch.empty(L.query("b2s_scan_workspace_bytes", cap_rows), dtype=torch.uint8, device=dev)
            L.call("b2s_coordmap_insert", cmap.coords, cap_rows, cmap.n_dev, L.host_i32(1, 1, 1), table, hcap, slot,
                   rank, info, scan_ws)
            cmap.table, cmap.capacity, cmap.info = table, hcap, info
            if self.static:
                self.checks.append(("coordinate range flag", 0, info[1:2]))
            self._mark_built()
        return cmap.table

    # ------------------------------------------------------------------ map construction
    def _build(self, coords: torch.Tensor, ts_floor, want_in2out=True):
        """Hash-insert ``coords`` floored to ``ts_floor``; returns (CoordMap, in2out or None, n_unique)."""
        n = coords.shape[0]
        dev = coords.device
        cap = L.query("b2s_hash_capacity", n)
        table = torch.empty(cap * 16, dtype=torch.uint8, device=dev)
        slot = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        rank = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        info = torch.empty(4, dtype=torch.int32, device=dev)
        scan_ws = torch.empty(L.query("b2s_scan_workspace_bytes", n), dtype=torch.uint8, device=dev)
        ts = L.host_i32(*ts_floor)
        L.call("b2s_coordmap_insert", coords, n, None, ts, table, cap, slot, rank, info, scan_ws)
        n_unique, overflow, max_batch, _ = info.tolist()          # the one host sync of a new map
        if overflow:
            raise L.B2SError("coordinate outside the packed 16-bit range (|c| < 32000, 0 <= batch < 65535): "
                             "B2S_EOVERFLOW")
        self.num_batches = max(self.num_batches, max_batch + 1)
        if n_unique == n and tuple(ts_floor) == (1, 1, 1):
            return CoordMap(coords, table, cap), None, n       # already unique: table values are the rows
        out = torch.empty((n_unique, 4), dtype=torch.int32, device=dev)
        in2out = torch.empty(max(n, 1), dtype=torch.int32, device=dev) if want_in2out else None
        L.call("b2s_coordmap_fill", coords, n, None, ts, table, cap, slot, rank, out, n_unique, in2out)
        _map_account("strided map (b2s_coordmap_insert + _fill)", 16 * n + 16 * n_unique + 4 * n)
        return CoordMap(out, table, cap), (in2out[:n] if want_in2out else None), n_unique

    def insert(self, coords: torch.Tensor, tensor_stride=(1, 1, 1), tag="", dense_index=None):
        """Create the map of a new SparseTensor.  Returns (key, unique_index or None).  ``dense_index``: the
        ``"index"`` entry of the GridSampling3D output these coordinates came from (optional accelerator)."""
        coords = coords.to(torch.int32).contiguous()
        assert coords.dim() == 2 and coords.shape[1] == 4, "coordinates must be [N, 1+3] (batch first)"
        key = CoordinateMapKey(tensor_stride, tag)
        cmap, in2out, n_unique = self._build(coords, (1, 1, 1))
        self.maps[key] = cmap
        self.device = coords.device
        unique_index = None
        if in2out is None:
            cmap.dense = dense_index                            # rows are the quantiser's rows, unchanged
        if in2out is not None:                                  # duplicates: keep the first occurrence
            n = coords.shape[0]
            first = torch.full((n_unique,), n, dtype=torch.int64, device=coords.device)
            first.scatter_reduce_(0, in2out.long(), torch.arange(n, device=coords.device), reduce="amin")
            unique_index = first
        return key, unique_index

    def stride(self, in_key: CoordinateMapKey, stride):
        stride = _triple(stride)
        ts = tuple(a * b for a, b in zip(in_key.tensor_stride, stride))
        out_key = CoordinateMapKey(ts, in_key.tag)
        self._join(("map", out_key))
        if out_key not in self.maps:
            self._note(("stride", in_key, stride))
            if self.static:
                self.maps[out_key] = self._build_static(self.maps[in_key], ts)
            else:
                cmap, _, _ = self._build(self.maps[in_key].coords, ts, want_in2out=False)
                self.maps[out_key] = cmap
            self._mark_built()
        return out_key

    def _build_static(self, imap: CoordMap, ts):
        """Strided map at fixed capacity: the unique count stays on the device (info[0])."""
        n, dev = imap.n, imap.coords.device
        cap_out = self.capacity_of(ts)
        hcap = L.query("b2s_hash_capacity", n)
        table = torch.empty(hcap * 16, dtype=torch.uint8, device=dev)
        slot = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        rank = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        info = torch.empty(4, dtype=torch.int32, device=dev)
        scan_ws = torch.empty(L.query("b2s_scan_workspace_bytes", n), dtype=torch.uint8, device=dev)
        tsh = L.host_i32(*ts)
        L.call("b2s_coordmap_insert", imap.coords, n, imap.n_dev, tsh, table, hcap, slot, rank, info, scan_ws)
        out = torch.empty((cap_out, 4), dtype=torch.int32, device=dev)
        L.call("b2s_coordmap_fill", imap.coords, n, imap.n_dev, tsh, table, hcap, slot, rank, out, cap_out, None)
        self.checks.append((f"rows at tensor stride {ts[0]}", cap_out, info[0:1]))
        self.checks.append((f"coordinate range flag at tensor stride {ts[0]}", 0, info[1:2]))
        return CoordMap(out, table, hcap, n_dev=info[0:1], info=info)

    def union(self, key_a, key_b):
        """Union map of two maps of the same tensor stride (``SparseTensor.__add__`` across coordinate maps): the rows
        of ``key_a`` in their order, then the rows only ``key_b`` has.  Returns (key, rows of a, rows of b) as int64
        index tensors.  Dynamic mode only (one host sync for the row count)."""
        if self.static:
            raise L.B2SError("union maps are built in dynamic mode only")
        assert key_a.tensor_stride == key_b.tensor_stride, "union of maps with different tensor strides"
        a, b = self.maps[key_a], self.maps[key_b]
        both = torch.cat([a.coords, b.coords])
        cmap, in2out, _ = self._build(both, (1, 1, 1))
        if in2out is None:          # disjoint: the concatenation is already unique
            in2out = torch.arange(both.shape[0], dtype=torch.int32, device=both.device)
        key = CoordinateMapKey(key_a.tensor_stride, f"union({key_a.tag}|{key_b.tag}|{len(self.maps)})")
        self.maps[key] = cmap
        return key, in2out[:a.n].long(), in2out[a.n:].long()

    def origin(self, key=None):
        """Key of the per-plot origin map (one row per batch id), as MinkowskiGlobalPooling returns."""
        key = CoordinateMapKey((0, 0, 0), "origin")
        if key not in self.maps:
            c = torch.zeros((self.num_batches, 4), dtype=torch.int32, device=self.device)
            c[:, 0] = torch.arange(self.num_batches, device=self.device, dtype=torch.int32)
            self.maps[key] = CoordMap(c, None, 0)
        return key

    # ------------------------------------------------------------------ kernel maps
    def _probe(self, query_map: CoordMap, table_map: CoordMap, kernel_size, step, sign):
        n = query_map.n
        k3 = kernel_size[0] * kernel_size[1] * kernel_size[2]
        nbr = torch.empty((k3, n), dtype=torch.int32, device=query_map.coords.device)
        if table_map.dense is not None and USE_DENSE_INDEX:
            ws, lo, dims, num_plots = table_map.dense
            L.call("b2s_kernel_map_dense", query_map.coords, n, query_map.n_dev, ws, num_plots, L.host_i32(*lo),
                   L.host_i32(*dims), L.host_i32(*kernel_size), L.host_i32(*step), sign, nbr)
            if MAP_STATS is not None:
                _map_account("b2s_kernel_map_dense", 16 * n + 8 * int((nbr >= 0).sum().item()))
            return nbr
        table = self._table_of(table_map)
        L.call("b2s_kernel_map", query_map.coords, n, query_map.n_dev, table, table_map.capacity,
               L.host_i32(*kernel_size), L.host_i32(*step), sign, nbr)
        if MAP_STATS is not None:
            _map_account("b2s_kernel_map", 16 * n + 8 * int((nbr >= 0).sum().item()))
        return nbr

    def kernel_map(self, in_key, out_key, kernel_size, dilation=(1, 1, 1)) -> KernelMap:
        kernel_size, dilation = _triple(kernel_size), _triple(dilation)
        ck = (in_key, out_key, kernel_size, dilation)
        self._join(("kmap", ck))
        km = self.kernel_maps.get(ck)
        if km is None:
            self._note(("kmap", in_key, out_key, kernel_size, dilation))
            step = tuple(d * t for d, t in zip(dilation, in_key.tensor_stride))
            imap, omap = self.maps[in_key], self.maps[out_key]
            km = KernelMap(self, in_key, out_key, kernel_size, step, None, imap.n, omap.n, imap.n_dev, omap.n_dev)
            if not km.lines_ok:          # maps with an x-line form build either table on first use (KernelMap.nbr / .lines)
                km.nbr = self._probe(omap, imap, kernel_size, step, +1)
                self._mark_built()
            self.kernel_maps[ck] = km
        return km

    # ------------------------------------------------------------------ per-plot info (origin map)
    def coords(self, key):
        return self.maps[key].coords

    def n_dev(self, key):
        """Device row count of a map (None in dynamic mode and for the origin map)."""
        return self.maps[key].n_dev

    def verify(self):
        """STATIC mode: one host read of every recorded device count; raises if a capacity was exceeded or a
        coordinate was out of the packed range.  Call after the step (or replay) has been enqueued."""
        if not self.checks:
            return {}
        vals = torch.cat([t.reshape(-1)[:1] for _, _, t in self.checks]).tolist()
        out = {}
        for (what, cap, _), v in zip(self.checks, vals):
            out[what] = v
            if v < 0:
                raise L.B2SError(f"{what}: the device reported {v} (a point outside the voxel bounds)")
            if cap == 0:
                if v != 0:
                    raise L.B2SError(f"{what} raised on the device (B2S_EOVERFLOW)")
            elif v > cap:
                raise L.B2SError(f"{what}: {v} exceeds the planned capacity {cap}; re-plan with larger capacities")
        return out

    def inv_counts(self, key):
        """float32 [B]: 1 / (rows of each plot) -- the average-pooling scale."""
        self._join(("invc", key))
        m = self.maps[key]
        if m._inv_counts is None:
            self._note(("invc", key))
            counts = torch.empty(self.num_batches, dtype=torch.int32, device=m.coords.device)
            L.call("b2s_batch_counts", m.coords, 4, m.n, m.n_dev, self.num_batches, counts)
            m._inv_counts = 1.0 / counts.clamp(min=1).float()
            m._counts_host = None
            self._mark_built()
        return m._inv_counts

    def rows_per_batch(self, key):
        """Host list of row counts per plot (syncs once per map; used by decomposed_coordinates)."""
        m = self.maps[key]
        if m._counts_host is None:
            counts = torch.empty(self.num_batches, dtype=torch.int32, device=m.coords.device)
            L.call("b2s_batch_counts", m.coords, 4, m.n, m.n_dev, self.num_batches, counts)
            m._counts_host = counts.tolist()
        return m._counts_host

"""Integer half of the oracle: voxel quantisation, strided coordinate maps, kernel maps.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED: restates the published
behaviour of pyg / torch-cluster / MinkowskiEngine, none of which is present in the reference tree.

Layout conventions (shared with the CUDA path):

* coordinates are ``int32 [N, 4] = [batch, x, y, z]`` -- batch index first, as built at
  ``torch_points3d/models/instance/minkowski.py:69`` of the reference;
* a kernel map is a *neighbour table* ``nbr int32 [K^3, N_out]`` with ``nbr[k, o] = i`` iff
  ``coords_in[i] == coords_out[o] + delta_k`` (same batch) and ``-1`` otherwise; the pair-list form
  MinkowskiEngine exposes is derived from it by :func:`table_to_pairs` (per offset, sorted by out row);
* kernel offset index ``k = ix + K*iy + K*K*iz`` (x fastest), ``delta = (i - K//2) * step`` for odd
  K and ``i * step`` for even K, ``step = dilation * tensor_stride_in``.
"""
from __future__ import annotations

import numpy as np

_BIAS = 1 << 15
_SPAN = 1 << 16


# --------------------------------------------------------------------------------------------
# key packing (same field layout as the CUDA hash: batch | z | y | x, 16 bits each, biased)
# --------------------------------------------------------------------------------------------
def pack_keys(coords: np.ndarray) -> np.ndarray:
    """int32 [N,4] (b,x,y,z) -> int64 [N]; ordering of keys == lexicographic (b, z, y, x)."""
    c = np.asarray(coords, dtype=np.int64)
    if c.size and (np.abs(c[:, 1:]).max() >= _BIAS - 8 or c[:, 0].min() < 0 or c[:, 0].max() >= _SPAN - 1):
        raise OverflowError("coordinate outside the 16-bit packed range")
    return ((c[:, 0] * _SPAN + (c[:, 3] + _BIAS)) * _SPAN + (c[:, 2] + _BIAS)) * _SPAN + (c[:, 1] + _BIAS)


# --------------------------------------------------------------------------------------------
# (a1) voxel quantisation -- GridSampling3D(mode="last", quantize_coords=True)
# --------------------------------------------------------------------------------------------
def quantize_points(pos: np.ndarray, size: float) -> np.ndarray:
    """``coords = torch.round(data.pos / grid_size)`` -- grid_transform.py:116.

    fp32 true division by the fp32-cast scalar, then round-half-to-even, kept as float32 exactly
    like the reference (the cast to int happens after the representative is chosen, :125)."""
    pos = np.asarray(pos, dtype=np.float32)
    return np.rint(pos / np.float32(size)).astype(np.float32)


def grid_cluster(coords_f: np.ndarray) -> np.ndarray:
    """torch_cluster.grid_cluster(coords, size=[1,1,1]) -- called at grid_transform.py:117-118.

    Published semantics (pytorch-cluster ``grid_cpu.cpp``): ``start = pos.min(0)``,
    ``end = pos.max(0)``, ``num_voxels_d = int((end_d - start_d) / size_d) + 1``,
    ``cluster = sum_d int((pos_d - start_d) / size_d) * prod_{j<d} num_voxels_j`` (x fastest)."""
    c = np.asarray(coords_f, dtype=np.float32)
    start = c.min(0)
    end = c.max(0)
    rel = (c - start).astype(np.int64)
    nvox = ((end - start)).astype(np.int64) + 1
    return rel[:, 0] + nvox[0] * (rel[:, 1] + nvox[1] * rel[:, 2])


def consecutive_cluster(cluster: np.ndarray):
    """torch_geometric ``consecutive_cluster`` -- called at grid_transform.py:121.

    ``unique, inv = torch.unique(src, sorted=True, return_inverse=True)``;
    ``perm = empty(len(unique)).scatter_(0, inv, arange(N))`` -- on CPU the scatter is sequential,
    so the LAST index of every cluster wins."""
    unique, inv = np.unique(cluster, return_inverse=True)
    perm = np.empty(unique.shape[0], dtype=np.int64)
    perm[inv] = np.arange(cluster.shape[0])  # numpy: repeated index -> last assignment wins
    return inv, perm


def quantize_plot(pos: np.ndarray, feats: np.ndarray, size: float, perm: np.ndarray | None):
    """GridSampling3D._process for one plot, mode="last" -- grid_transform.py:112-128.

    ``perm`` is the permutation ``shuffle_data`` draws with ``torch.randperm`` (:22-29); it is an
    explicit input so that the CUDA path can be compared on identical shuffles (``None`` = identity).
    Returns ``coords int32 [M,3]``, ``feats [M,F]``, ``pos [M,3]`` and ``src int64 [M]`` (index of the
    representative point in the ORIGINAL, unshuffled order).  Rows come out sorted by (z, y, x)."""
    pos = np.asarray(pos, dtype=np.float32)
    n = pos.shape[0]
    perm = np.arange(n, dtype=np.int64) if perm is None else np.asarray(perm, dtype=np.int64)
    pos_s = pos[perm]                                   # shuffle_data
    coords_f = quantize_points(pos_s, size)             # :116
    cluster = grid_cluster(coords_f)                    # :117-118
    _, rep = consecutive_cluster(cluster)               # :121
    src = perm[rep]
    coords = coords_f[rep].astype(np.int32)             # :124-125 (.int() truncation of exact ints)
    return coords, np.asarray(feats)[src], pos[src], src


def quantize_batch(pos_list, feat_list, size, perm_list=None):
    """Per-plot quantisation followed by the PyG collate + batch column of minkowski.py:69.

    Returns ``coords int32 [sum M, 4]`` (b,x,y,z), ``feats``, ``pos``, ``src`` (index into the
    concatenated original point array) and ``rows_per_plot``."""
    out_c, out_f, out_p, out_s, counts = [], [], [], [], []
    base = 0
    for b, (p, f) in enumerate(zip(pos_list, feat_list)):
        perm = None if perm_list is None else perm_list[b]
        c, ff, pp, s = quantize_plot(p, f, size, perm)
        out_c.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], 1))
        out_f.append(ff)
        out_p.append(pp)
        out_s.append(s + base)
        counts.append(c.shape[0])
        base += np.asarray(p).shape[0]
    return (np.concatenate(out_c), np.concatenate(out_f), np.concatenate(out_p),
            np.concatenate(out_s), np.asarray(counts, dtype=np.int64))


# --------------------------------------------------------------------------------------------
# (a2) first-occurrence unique -- SparseTensor creation on possibly duplicated coordinates
# --------------------------------------------------------------------------------------------
def unique_first(coords: np.ndarray):
    """Unique rows in FIRST-OCCURRENCE order (CPU-MinkowskiEngine ``insert_and_map`` order).

    Returns ``first_idx`` (input row of every unique row) and ``inverse`` (unique row of every input)."""
    keys = pack_keys(coords)
    _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")            # sorted-key rank -> first-occurrence rank
    rank = np.empty_like(order)
    rank[order] = np.arange(order.shape[0])
    return first[order], rank[inv]


# --------------------------------------------------------------------------------------------
# (a3) strided coordinate map
# --------------------------------------------------------------------------------------------
def stride_map(coords: np.ndarray, ts_out):
    """Coordinates of the coarser map a ``stride=s`` op produces (ME ``CoordinateMap::stride``).

    ``c'_d = floor(c_d / ts_out_d) * ts_out_d`` on the spatial columns (floor toward -inf), batch
    untouched, unique rows in first-occurrence order.  Returns (out_coords int32 [M,4], in2out [N])."""
    c = np.asarray(coords, dtype=np.int32)
    ts = np.broadcast_to(np.asarray(ts_out, dtype=np.int32), (3,))
    f = c.copy()
    f[:, 1:] = np.floor_divide(c[:, 1:], ts) * ts
    first, inv = unique_first(f)
    return f[first], inv.astype(np.int32)


# --------------------------------------------------------------------------------------------
# (a4) kernel maps
# --------------------------------------------------------------------------------------------
def kernel_offsets(kernel_size, step) -> np.ndarray:
    """int32 [K^3, 3] offsets (dx,dy,dz), x fastest; ME ``kernel_region`` iteration order."""
    K = np.broadcast_to(np.asarray(kernel_size, dtype=np.int64), (3,))
    st = np.broadcast_to(np.asarray(step, dtype=np.int64), (3,))
    axes = []
    for d in range(3):
        r = np.arange(K[d])
        if K[d] % 2 == 1:
            r = r - K[d] // 2
        axes.append(r * st[d])
    dz, dy, dx = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
    return np.stack([dx.ravel(), dy.ravel(), dz.ravel()], 1).astype(np.int32)


def kernel_map_table(in_coords: np.ndarray, out_coords: np.ndarray, kernel_size, step, sign: int = 1):
    """Neighbour table ``nbr int32 [K^3, N_out]``: ``nbr[k,o] = i`` iff ``in[i] == out[o] + sign*delta_k``.

    ``sign=+1`` is the forward map of MinkowskiConvolution / MinkowskiMaxPooling; ``sign=-1`` with the
    roles of the two coordinate sets swapped gives the transposed map used by dgrad and by
    MinkowskiConvolutionTranspose."""
    in_coords = np.asarray(in_coords, dtype=np.int32)
    out_coords = np.asarray(out_coords, dtype=np.int32)
    offs = kernel_offsets(kernel_size, step)
    keys = pack_keys(in_coords)
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]
    n_out = out_coords.shape[0]
    nbr = np.full((offs.shape[0], n_out), -1, dtype=np.int32)
    if skeys.shape[0] == 0 or n_out == 0:
        return nbr
    q = out_coords.astype(np.int64)
    for k, d in enumerate(offs):
        qq = q.copy()
        qq[:, 1:] += sign * d.astype(np.int64)
        qk = ((qq[:, 0] * _SPAN + (qq[:, 3] + _BIAS)) * _SPAN + (qq[:, 2] + _BIAS)) * _SPAN + (qq[:, 1] + _BIAS)
        pos = np.searchsorted(skeys, qk)
        pos_c = np.minimum(pos, skeys.shape[0] - 1)
        hit = skeys[pos_c] == qk
        nbr[k, hit] = order[pos_c[hit]].astype(np.int32)
    return nbr


def kernel_map_lines(nbr: np.ndarray, kernel_size) -> np.ndarray:
    """x-LINE form of a stride-1 neighbour table over rows sorted by (batch, z, y, x) -- the layout the k7 stem
    kernels read (include/b200sparse.h ``b2s_kernel_map_lines``; call site of the layer: ME/SENet.py:49-52).

    ``lines uint32 [K1*K2, N]``: line ``l = iy + K1*iz`` of row q packs its ``K0`` x-consecutive offsets as
    ``(base << 8) | mask`` with mask bit ix set iff ``nbr[ix + K0*l, q] >= 0`` and base = the smallest of those rows.
    Because the in rows are sorted with x fastest, the existing neighbours of a line are consecutive rows, so the
    table is recovered exactly by :func:`lines_to_table` (asserted here)."""
    K = np.broadcast_to(np.asarray(kernel_size, dtype=np.int64), (3,))
    k0, nl = int(K[0]), int(K[1] * K[2])
    assert k0 <= 8 and nbr.shape[0] == k0 * nl and nbr.shape[1] < (1 << 24)
    t = nbr.reshape(nl, k0, -1).astype(np.int64)
    valid = t >= 0
    mask = (valid * (1 << np.arange(k0, dtype=np.int64))[None, :, None]).sum(1)
    base = np.where(valid, t, np.iinfo(np.int64).max).min(1)
    base = np.where(mask != 0, base, 0)
    rank = np.cumsum(valid, 1) - valid                     # number of existing neighbours below ix
    assert np.all(np.where(valid, t == base[:, None, :] + rank, True)), "rows of a line are not consecutive"
    return ((base << 8) | mask).astype(np.uint32)


def lines_to_table(lines: np.ndarray, kernel_size) -> np.ndarray:
    """Inverse of :func:`kernel_map_lines`: ``nbr[ix + K0*l, q] = base + popcount(mask & ((1 << ix) - 1))``."""
    K = np.broadcast_to(np.asarray(kernel_size, dtype=np.int64), (3,))
    k0 = int(K[0])
    w = lines.astype(np.int64)
    mask, base = w & 0xFF, w >> 8
    out = np.full((lines.shape[0], k0, lines.shape[1]), -1, dtype=np.int32)
    below = np.zeros_like(mask)
    for ix in range(k0):
        bit = (mask >> ix) & 1
        out[:, ix, :] = np.where(bit == 1, base + below, -1)
        below = below + bit
    return out.reshape(lines.shape[0] * k0, lines.shape[1])


def table_to_pairs(nbr: np.ndarray):
    """Pair-list form: per offset k the (in,out) pairs sorted by out row; ``offsets int64 [K^3+1]``."""
    ins, outs, offsets = [], [], [0]
    for k in range(nbr.shape[0]):
        o = np.nonzero(nbr[k] >= 0)[0]
        ins.append(nbr[k, o])
        outs.append(o.astype(np.int32))
        offsets.append(offsets[-1] + o.shape[0])
    cat = (lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int32))
    return cat(ins).astype(np.int32), cat(outs).astype(np.int32), np.asarray(offsets, dtype=np.int64)


# --------------------------------------------------------------------------------------------
# (a5) origin map
# --------------------------------------------------------------------------------------------
def batch_rows(coords: np.ndarray, num_batches: int | None = None):
    """Row -> batch id and per-batch row counts (ME ``origin_map``)."""
    b = np.asarray(coords)[:, 0].astype(np.int64)
    nb = int(b.max()) + 1 if num_batches is None else num_batches
    return b.astype(np.int32), np.bincount(b, minlength=nb).astype(np.int64)

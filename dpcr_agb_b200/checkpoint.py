"""Checkpoint files in the reference's layout (``torch_points3d/metrics/model_checkpoint.py:24-61,182-193,237-243``).

The reference pickles one dict per run: ``models`` {weight name -> state_dict} (``"latest"`` plus ``best_<metric>``
entries), ``optimizer`` (class name, state_dict), ``schedulers``, ``grad_scale``, ``stats``, ``run_config``,
``dataset_properties``.  The model it saves is ``MinkowskiBaselineModel`` whose backbone sits under ``model.``
(``models/instance/minkowski.py:32-41``); :class:`dpcr_agb_b200.msenet.MSENet` has the same keys without that prefix
(same module attribute names, head included: ``final.linears.N``).  ``save`` / ``load`` translate between the two, so
a file written here opens with ``Checkpoint.load`` + ``model.load_state_dict`` there and a released reference
checkpoint loads into :class:`MSENet`.
"""
from __future__ import annotations

import torch

LATEST = "latest"
PREFIX = "model."


def to_reference_keys(state_dict):
    return {PREFIX + k: v for k, v in state_dict.items()}


def from_reference_keys(state_dict):
    """Accepts both layouts (with or without the ``model.`` prefix of ``MinkowskiBaselineModel``)."""
    return {(k[len(PREFIX):] if k.startswith(PREFIX) else k): v for k, v in state_dict.items()}


def save(path, model, optimizer=None, schedulers=None, stats=None, run_config=None, extra_models=None,
         dataset_properties=None):
    """Write ``model`` (and optionally the :class:`dpcr_agb_b200.train.FlatAdaBelief` state) as a reference-layout
    checkpoint.  ``extra_models``: {"best_<metric>": state_dict} entries next to ``latest``."""
    models = {LATEST: to_reference_keys({k: v.detach().cpu().clone() for k, v in model.state_dict().items()})}
    for name, sd in (extra_models or {}).items():
        models[name] = to_reference_keys({k: v.detach().cpu().clone() for k, v in sd.items()})
    opt = None
    if optimizer is not None:
        sd = optimizer.state_dict()
        for st in sd["state"].values():
            for k, v in list(st.items()):
                if torch.is_tensor(v):
                    st[k] = v.detach().cpu()
        opt = ("AdaBelief", sd)
    obj = {"run_config": run_config or {}, "models": models, "stats": stats or {"train": [], "test": [], "val": []},
           "optimizer": opt, "grad_scale": {}, "schedulers": schedulers or {},
           "dataset_properties": dataset_properties or {}}
    torch.save(obj, path)
    return obj


def load(path, model, optimizer=None, weight_name=LATEST, strict=True):
    """Load ``weight_name`` (``model_checkpoint.py:237-243`` falls back to ``latest`` when the name is missing) into
    ``model`` and, if given, the optimiser state.  Returns the checkpoint dict."""
    obj = torch.load(path, map_location="cpu", weights_only=False)
    models = obj["models"]
    if weight_name not in models:
        weight_name = LATEST
    model.load_state_dict(from_reference_keys(models[weight_name]), strict=strict)
    if optimizer is not None and obj.get("optimizer"):
        optimizer.load_state_dict(obj["optimizer"][1])
    return obj

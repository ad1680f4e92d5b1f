"""Validate and load reference checkpoints.

The reference saves plain ``state_dict`` files - ``torch.save(model.module.state_dict(), ...)`` under DataParallel,
``model.state_dict()`` otherwise (train_bid_pointconv.py:173-177, distilTrain.py:200-204) - and loads them with a bare
``load_state_dict(torch.load(path))`` (evaluate_bid_pointconv.py:87, distilTrain.py:104,121): a file saved from the
wrapped module (``module.`` prefixes), one wrapped by ``main_utils.save_checkpoint`` (main_utils.py:49-56: a dict with the
weights under ``state_dict``) or one from a different model variant fails there with a wall of key names.  ``load_reference_checkpoint``
accepts all three layouts, checks every tensor against the model (name, shape, dtype, finiteness) before touching it and
reports exactly what differs.  The models of this package keep the reference's parameter names and shapes
(tests/golden/state_dict_keys.json), so a valid reference checkpoint loads without any renaming.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Mapping, Tuple, Union

import torch

_WRAPPER_KEYS = ("state_dict", "model_state_dict", "model", "net")


@dataclass
class CheckpointReport:
    matched: int = 0
    missing: List[str] = field(default_factory=list)              # in the model, not in the file
    unexpected: List[str] = field(default_factory=list)           # in the file, not in the model
    shape_mismatch: List[Tuple[str, Tuple[int, ...], Tuple[int, ...]]] = field(default_factory=list)   # (name, model, file)
    non_finite: List[str] = field(default_factory=list)
    stripped_prefix: str = ""
    unwrapped_key: str = ""

    @property
    def ok(self) -> bool:
        return not (self.missing or self.unexpected or self.shape_mismatch or self.non_finite)

    def summary(self) -> str:
        lines = [f"{self.matched} tensors match"]
        if self.unwrapped_key:
            lines.append(f"weights found under '{self.unwrapped_key}'")
        if self.stripped_prefix:
            lines.append(f"prefix '{self.stripped_prefix}' stripped from every key")
        for title, items in (("missing (model has, file lacks)", self.missing), ("unexpected (file has, model lacks)", self.unexpected),
                             ("non-finite values", self.non_finite)):
            if items:
                lines.append(f"{len(items)} {title}: " + ", ".join(items[:8]) + (" ..." if len(items) > 8 else ""))
        if self.shape_mismatch:
            lines.append(f"{len(self.shape_mismatch)} shape mismatches: " +
                         ", ".join(f"{n} model{m} file{f}" for n, m, f in self.shape_mismatch[:8]))
        return "; ".join(lines)


def _unwrap(obj) -> Tuple[Mapping[str, torch.Tensor], str]:
    """The tensor dict inside whatever ``torch.load`` returned."""
    if isinstance(obj, torch.nn.Module):
        return obj.state_dict(), "<module>"
    if not isinstance(obj, Mapping):
        raise TypeError(f"kdpc: a checkpoint must be a state_dict or a dict holding one, got {type(obj).__name__}")
    if obj and all(torch.is_tensor(v) for v in obj.values()):
        return obj, ""
    for k in _WRAPPER_KEYS:
        if k in obj and isinstance(obj[k], Mapping):
            inner, _ = _unwrap(obj[k])
            return inner, k
    raise ValueError("kdpc: no state_dict found in the checkpoint (looked for a flat tensor dict or one under "
                     + ", ".join(_WRAPPER_KEYS) + ")")


def _strip_common_prefix(sd: Mapping[str, torch.Tensor], model_keys) -> Tuple[Dict[str, torch.Tensor], str]:
    """DataParallel / DistributedDataParallel ``module.`` (possibly nested) in front of EVERY key."""
    out, stripped = dict(sd), ""
    while out and all(k.startswith("module.") for k in out) and not any(k in model_keys for k in out):
        out = {k[len("module."):]: v for k, v in out.items()}
        stripped += "module."
    return out, stripped


def check_reference_checkpoint(model: torch.nn.Module, checkpoint: Union[str, Mapping, torch.nn.Module]):
    """Compare a checkpoint (path or loaded object) with ``model`` without modifying the model.
    Returns (report, clean_state_dict)."""
    obj = torch.load(checkpoint, map_location="cpu", weights_only=True) if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__") else checkpoint
    sd, unwrapped = _unwrap(obj)
    want = model.state_dict()
    sd, stripped = _strip_common_prefix(sd, want.keys())
    rep = CheckpointReport(stripped_prefix=stripped, unwrapped_key=unwrapped)
    for name, ref in want.items():
        if name not in sd:
            rep.missing.append(name)
            continue
        t = sd[name]
        if tuple(t.shape) != tuple(ref.shape):
            rep.shape_mismatch.append((name, tuple(ref.shape), tuple(t.shape)))
        elif t.is_floating_point() and not bool(torch.isfinite(t).all()):
            rep.non_finite.append(name)
        else:
            rep.matched += 1
    rep.unexpected = [k for k in sd if k not in want]
    return rep, sd


def load_reference_checkpoint(model: torch.nn.Module, checkpoint, strict: bool = True) -> CheckpointReport:
    """``model.load_state_dict`` for reference checkpoints, with the layouts above handled and a readable error.
    strict=False loads every tensor that matches by name and shape and reports the rest."""
    rep, sd = check_reference_checkpoint(model, checkpoint)
    if strict and not rep.ok:
        raise RuntimeError("kdpc: checkpoint does not fit the model: " + rep.summary())
    want = model.state_dict()
    bad = {n for n, _, _ in rep.shape_mismatch} | set(rep.non_finite)
    usable = {k: v.to(dtype=want[k].dtype) for k, v in sd.items() if k in want and k not in bad}
    model.load_state_dict(usable, strict=False)
    return rep

"""Checkpoint readers and name / layout mapping for the models of this package.

The models' `load_state_dict(state_dict)` takes REFERENCE-named parameters in PADDLE layouts (Linear weight `[in, out]`,
conv `[O, I, kh, kw]`), exactly what `ppdiffusers` / `paddlemix` checkpoints hold. This module gets such a dict from
the files a user of the reference has on disk:

* `.safetensors` (single file or the sharded `*.safetensors.index.json` layout) - read with a small pure-Python
  reader (the `safetensors` wheel is not a dependency): 8-byte little-endian header length, JSON header
  `{name: {dtype, shape, data_offsets}, "__metadata__": {...}}`, then the raw little-endian tensor bytes.
  The archive's `format` metadata decides the layout like the reference does
  (ppdiffusers/ppdiffusers/models/modeling_utils.py:150-176): "pt" (or absent) = torch layout, "pd" / "np" = paddle.
* `.pdparams` - `paddle.save` of a state dict is a pickle of `{name: numpy.ndarray}` (bfloat16 stored as uint16 plus a
  few bookkeeping keys); it unpickles without paddle (modeling_utils.py:177-208 `smart_load(..., return_numpy=True)`).
* torch-layout dicts are converted the way `convert_pytorch_state_dict_to_paddle` does
  (ppdiffusers/ppdiffusers/models/modeling_pytorch_paddle_utils.py:27-64): every `nn.Linear` weight is transposed,
  `position_ids` / `num_batches_tracked` are dropped. Which keys are Linear weights comes from the model itself
  (`linear_weight_keys(model)`: its 2-D `.weight` parameters except embedding tables).

Nothing here touches the GPU: tensors stay on the host until `model.load_state_dict` moves them.
"""
from __future__ import annotations

import json
import mmap
import os
import pickle
import struct
from typing import Any, Dict, Iterable, Optional, Tuple

import numpy as np
import torch

_ST_DTYPES = {
    "F64": (torch.float64, 8), "F32": (torch.float32, 4), "F16": (torch.float16, 2), "BF16": (torch.bfloat16, 2),
    "I64": (torch.int64, 8), "I32": (torch.int32, 4), "I16": (torch.int16, 2), "I8": (torch.int8, 1),
    "U8": (torch.uint8, 1), "BOOL": (torch.bool, 1),
}
_ST_NAMES = {v[0]: k for k, v in _ST_DTYPES.items()}

# 2-D `.weight` parameters that are lookup tables, not Linear layers (never transposed)
EMBEDDING_SUFFIXES = ("embed_tokens.weight", "wte.weight", "wpe.weight", "position_embedding.weight",
                      "token_embedding.weight", "class_embedding.weight")
IGNORED_KEY_PARTS = ("position_ids", ".num_batches_tracked", "StructuredToParameterName@@")


class CheckpointError(ValueError):
    pass


# ---------------------------------------------------------------------------------------------------------------------
# safetensors
# ---------------------------------------------------------------------------------------------------------------------
def read_safetensors_header(path: str) -> Tuple[Dict[str, Any], int]:
    """Returns (header dict, byte offset of the data section)."""
    with open(path, "rb") as f:
        raw = f.read(8)
        if len(raw) != 8:
            raise CheckpointError(f"{path}: too short to be a safetensors archive")
        (n,) = struct.unpack("<Q", raw)
        if n <= 0 or n > 100 * 1024 * 1024:
            raise CheckpointError(f"{path}: implausible safetensors header length {n}")
        try:
            header = json.loads(f.read(n).decode("utf-8"))
        except Exception as e:  # noqa: BLE001
            raise CheckpointError(f"{path}: safetensors header is not valid JSON ({e})") from e
    return header, 8 + n


def read_safetensors(path: str, keys: Optional[Iterable[str]] = None) -> Tuple[Dict[str, torch.Tensor], Dict[str, str]]:
    """Reads tensors (all, or `keys`) as host torch tensors; returns (tensors, metadata). The file is memory-mapped and
    each tensor is copied out, so the mapping does not outlive the call."""
    header, base = read_safetensors_header(path)
    meta = header.get("__metadata__") or {}
    want = None if keys is None else set(keys)
    out: Dict[str, torch.Tensor] = {}
    size = os.path.getsize(path)
    with open(path, "rb") as f, mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) as mm:
        for name, info in header.items():
            if name == "__metadata__" or (want is not None and name not in want):
                continue
            if info["dtype"] not in _ST_DTYPES:
                raise CheckpointError(f"{path}: tensor {name} has unsupported dtype {info['dtype']}")
            dtype, esz = _ST_DTYPES[info["dtype"]]
            b0, b1 = info["data_offsets"]
            shape = tuple(info["shape"])
            numel = int(np.prod(shape, dtype=np.int64)) if shape else 1
            if b1 - b0 != numel * esz or base + b1 > size or b0 < 0:
                raise CheckpointError(f"{path}: tensor {name} has inconsistent offsets {b0}:{b1} for shape {shape}")
            buf = bytearray(mm[base + b0:base + b1])  # private copy (torch.frombuffer needs a writable buffer)
            out[name] = (torch.frombuffer(buf, dtype=dtype).reshape(shape) if numel else torch.empty(shape, dtype=dtype))
    if want is not None and want - set(out):
        raise CheckpointError(f"{path}: missing tensors {sorted(want - set(out))[:3]}")
    return out, meta


def write_safetensors(path: str, tensors: Dict[str, torch.Tensor], metadata: Optional[Dict[str, str]] = None) -> None:
    """Writes a safetensors archive (used to export converted weights and by the tests)."""
    header: Dict[str, Any] = {}
    if metadata:
        header["__metadata__"] = {str(k): str(v) for k, v in metadata.items()}
    blobs = []
    off = 0
    for name in sorted(tensors):
        t = tensors[name].detach().cpu().contiguous()
        if t.dtype not in _ST_NAMES:
            raise CheckpointError(f"cannot store {name}: dtype {t.dtype}")
        raw = t.reshape(-1).view(torch.uint8).numpy().tobytes() if t.numel() else b""
        header[name] = {"dtype": _ST_NAMES[t.dtype], "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    hj = json.dumps(header, separators=(",", ":")).encode("utf-8")
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for raw in blobs:
            f.write(raw)


# ---------------------------------------------------------------------------------------------------------------------
# .pdparams (paddle.save pickle of numpy arrays)
# ---------------------------------------------------------------------------------------------------------------------
class _NumpyOnlyUnpickler(pickle.Unpickler):
    """A `.pdparams` file (paddle.save of a dict of numpy arrays) needs exactly these globals to unpickle; every other
    (module, name) pair is refused. Names are matched exactly - protocol-4 STACK_GLOBAL accepts dotted names
    (`numpy._core._methods` + `os.system` resolves through the submodule's attributes), so a package-level allow-list
    is an arbitrary-code-execution hole."""
    _ALLOWED = {
        ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
        ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"),
        ("numpy", "ndarray"), ("numpy", "dtype"),
        ("collections", "OrderedDict"), ("_codecs", "encode"),
    }

    def find_class(self, module, name):
        if (module, name) not in self._ALLOWED or "." in name:
            raise CheckpointError(f"refusing to unpickle {module}.{name} from a parameter file")
        return super().find_class(module, name)


def read_pdparams(path: str, bf16_keys_as_uint16: bool = True) -> Dict[str, torch.Tensor]:
    with open(path, "rb") as f:
        obj = _NumpyOnlyUnpickler(f).load()
    if not isinstance(obj, dict):
        raise CheckpointError(f"{path}: expected a pickled dict of arrays, got {type(obj).__name__}")
    out: Dict[str, torch.Tensor] = {}
    for k, v in obj.items():
        if any(part in k for part in IGNORED_KEY_PARTS):
            continue
        if isinstance(v, np.ndarray):
            if v.dtype == np.uint16 and bf16_keys_as_uint16:  # paddle stores bfloat16 as uint16 bit patterns
                out[k] = torch.from_numpy(v.astype(np.int16, copy=True)).view(torch.bfloat16)
            else:
                out[k] = torch.from_numpy(np.ascontiguousarray(v))
        elif isinstance(v, (int, float)):
            out[k] = torch.tensor(v)
        elif isinstance(v, dict) and not v:
            continue
        else:
            raise CheckpointError(f"{path}: entry {k} is a {type(v).__name__}, not an array")
    return out


# ---------------------------------------------------------------------------------------------------------------------
# layout / name mapping
# ---------------------------------------------------------------------------------------------------------------------
def linear_weight_keys(model) -> set:
    """Keys of `model.state_dict_shapes()` that are nn.Linear weights (2-D `.weight`, embedding tables excluded)."""
    extra = tuple(getattr(model, "EMBEDDING_KEYS", ()))
    return {k for k, shp in model.state_dict_shapes().items()
            if k.endswith(".weight") and len(shp) == 2 and not k.endswith(EMBEDDING_SUFFIXES + extra)}


def torch_to_paddle_layout(model, state_dict: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """convert_pytorch_state_dict_to_paddle (modeling_pytorch_paddle_utils.py:27-64): transpose Linear weights, drop
    bookkeeping keys, BatchNorm running_{mean,var} -> _{mean,variance}."""
    lin = linear_weight_keys(model)
    out = {}
    for k, v in state_dict.items():
        if any(part in k for part in IGNORED_KEY_PARTS):
            continue
        k2 = k.replace(".running_var", "._variance").replace(".running_mean", "._mean")
        out[k2] = v.t() if (k2 in lin and v.ndim == 2) else v
    return out


def check_against_model(model, state_dict: Dict[str, torch.Tensor], strict: bool = True) -> Dict[str, list]:
    """Compares names and shapes with `model.state_dict_shapes()`; raises with the first few offenders when `strict`."""
    want = model.state_dict_shapes()
    missing = [k for k in want if k not in state_dict]
    unexpected = [k for k in state_dict if k not in want]
    bad_shape = [(k, tuple(state_dict[k].shape), tuple(want[k])) for k in want
                 if k in state_dict and tuple(state_dict[k].shape) != tuple(want[k])]
    if strict and (missing or bad_shape):
        msg = []
        if missing:
            msg.append(f"{len(missing)} missing (e.g. {missing[:3]})")
        if bad_shape:
            msg.append(f"{len(bad_shape)} with the wrong shape (e.g. {bad_shape[:2]}; wrong `layout`?)")
        raise CheckpointError("checkpoint does not match the model: " + "; ".join(msg))
    return dict(missing=missing, unexpected=unexpected, bad_shape=bad_shape)


def read_checkpoint(path: str) -> Tuple[Dict[str, torch.Tensor], str]:
    """Reads a checkpoint file, a sharded-safetensors index, or a directory holding one of them.
    Returns (state dict on the host, layout) with layout "torch" or "paddle"."""
    if os.path.isdir(path):
        names = sorted(os.listdir(path))
        pick = ([n for n in names if n.endswith(".safetensors.index.json")] or
                [n for n in names if n.endswith(".safetensors")] or [n for n in names if n.endswith(".pdparams")])
        if not pick:
            raise CheckpointError(f"{path}: no .safetensors / .safetensors.index.json / .pdparams inside")
        if len(pick) > 1 and not pick[0].endswith(".index.json"):
            raise CheckpointError(f"{path}: several candidate files {pick[:4]}; pass one explicitly")
        path = os.path.join(path, pick[0])
    if path.endswith(".safetensors.index.json"):
        with open(path) as f:
            index = json.load(f)
        by_file: Dict[str, list] = {}
        for k, fn in index["weight_map"].items():
            by_file.setdefault(fn, []).append(k)
        sd: Dict[str, torch.Tensor] = {}
        fmt = None
        for fn, keys in sorted(by_file.items()):
            part, meta = read_safetensors(os.path.join(os.path.dirname(path), fn), keys)
            fmt = fmt or meta.get("format")
            sd.update(part)
        return sd, _layout_of(fmt, path)
    if path.endswith(".safetensors"):
        sd, meta = read_safetensors(path)
        return sd, _layout_of(meta.get("format"), path)
    if path.endswith(".pdparams"):
        return read_pdparams(path), "paddle"
    raise CheckpointError(f"{path}: unsupported checkpoint type (want .safetensors, .safetensors.index.json, .pdparams)")


def _layout_of(fmt: Optional[str], path: str) -> str:
    fmt = fmt or "pt"  # the reference's default for archives without metadata (modeling_utils.py:170)
    if fmt not in ("pt", "pd", "np"):
        raise CheckpointError(f"{path}: safetensors metadata format={fmt!r} (want pt, pd or np)")
    return "torch" if fmt == "pt" else "paddle"


def load_pretrained(model, path: str, device=0, layout: str = "auto", strict: bool = True, prefix: str = ""):
    """Reads `path`, brings it to the reference's Paddle layout, checks names / shapes against the model and hands it to
    `model.load_state_dict` (which builds the fused / transposed device tensors). `prefix` strips a leading scope such
    as "unet." or "transformer." from every key."""
    sd, detected = read_checkpoint(path)
    if prefix:
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    layout = detected if layout == "auto" else layout
    if layout not in ("torch", "paddle"):
        raise CheckpointError(f"layout must be auto, torch or paddle (got {layout!r})")
    if layout == "torch":
        sd = torch_to_paddle_layout(model, sd)
    # parameters the reference leaves at their constructor default when the archive does not hold them (Paddle-only
    # LayerNorm biases of torch-format SD3 archives, the tied lm_head of Qwen2-VL-2B): the model says how to fill them
    fill = getattr(model, "default_missing_parameters", None)
    defaulted = []
    if fill is not None:
        for k, v in fill(sd).items():
            if k not in sd:
                sd[k] = v
                defaulted.append(k)
    report = check_against_model(model, sd, strict=strict)
    if report["missing"]:  # strict=False: zero-fill what is still missing (reported), like set_state_dict does
        want = model.state_dict_shapes()
        for k in report["missing"]:
            sd[k] = torch.zeros(want[k])
    report["defaulted"] = defaulted
    model.load_state_dict(sd, device=device)
    return report

"""Weight serialisation for the entropy coder (reference: pipeline/entropy_coding/__init__.py).

Same functions as the reference — `linear_state_dict`, `compress_state_dict`, `decompress_state_dict` — but the byte
stream (fp16 tensors and uint8 k-means codes of `model.half()`, back to back in state_dict order) is assembled ON
THE DEVICE by one kernel (sirenb200_pack_stream) and leaves it in one device->host copy; the reverse path unpacks a
stream straight into a model's fp32 parameters on the device (weight = centroids[labels]), so `eval_epoch` can run
from the compressed representation without a host round trip.  The entropy coders themselves (zstd / lzma) stay on
the host and out of scope; "plain" (the reference's NumpyParser) and "lzma" (stdlib) work here, "zstd" needs the
third-party `zstandard` module.
"""
import ctypes
import json
import lzma
from collections import OrderedDict
from pathlib import Path

import numpy as np
import torch
from torch import nn

from ... import _lib

PACK_F32_TO_F16, PACK_I64_TO_U8, PACK_I64_TO_U16, PACK_RAW = 0, 1, 2, 3
UNPACK_F16_TO_F32, UNPACK_GATHER_U8, UNPACK_GATHER_U16 = 4, 5, 6


def _layout(model):
    """[(name, tensor, kind, numpy dtype name)] in the order of linear_state_dict(model.half()) (reference :15-41):
    state_dict order; quantised Linear layers drop `weight` and keep `centroids` (fp16) + `labeled_weight`
    (uint8, or uint16 when a label exceeds 2**8 — the reference's rule, :35)."""
    quantised = set()
    for name, module in model.named_modules():
        if isinstance(module, nn.Linear) and hasattr(module, "centroids") and hasattr(module, "labeled_weight"):
            quantised.add(name)
    items = []
    for key, t in model.state_dict().items():
        mod = key.rsplit(".", 1)[0]
        if mod in quantised and key.endswith(".weight"):
            continue
        if mod in quantised and key.endswith(".labeled_weight"):
            wide = bool((t.max() > 2 ** 8).item())
            items.append((key, t, PACK_I64_TO_U16 if wide else PACK_I64_TO_U8, "uint16" if wide else "uint8"))
        elif t.is_floating_point():
            items.append((key, t, PACK_F32_TO_F16, "float16"))
        else:
            raise _lib.SirenB200Error(f"unexpected non-float tensor {key} in the state dict")
    return items


def pack_state_dict(model):
    """Device-side `linear_state_dict(model.half())` + concatenation: returns (stream [bytes] uint8 CUDA tensor,
    meta OrderedDict like the reference's meta_data.json).  The model itself is left untouched (fp32)."""
    items = _layout(model)
    if not items:
        raise _lib.SirenB200Error("nothing to pack")
    dev = items[0][1].device
    _lib.require_cuda(items[0][1], "model parameters")
    lib = _lib.load()
    n = len(items)
    srcs, kinds, counts, offsets = [], [], [], []
    meta, off = OrderedDict(), 0
    keep = []
    for order, (name, t, kind, dt) in enumerate(items):
        if kind == PACK_F32_TO_F16:
            src = t.detach().to(torch.float32).contiguous()
        else:
            src = t.detach().to(torch.int64).contiguous()
        keep.append(src)
        srcs.append(src.data_ptr())
        kinds.append(kind)
        counts.append(src.numel())
        offsets.append(off)
        meta[order] = {"shape": list(t.shape), "dtype": dt, "name": name}
        off += src.numel() * np.dtype(dt).itemsize
    stream = torch.empty(off, dtype=torch.uint8, device=dev)
    _lib.check(lib.sirenb200_pack_stream(
        n, (ctypes.c_void_p * n)(*srcs), None, (ctypes.c_int32 * n)(*kinds), (ctypes.c_int64 * n)(*counts),
        (ctypes.c_int64 * n)(*offsets), None, stream.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return stream, meta


def unpack_into_model(stream, meta, model):
    """Device-side decompress_state_dict (:123-186) + load: fp16 tensors -> fp32 parameters, and for quantised layers
    weight = centroids[labels], written straight into `model`'s parameters (fp32, on the stream's device)."""
    lib = _lib.load()
    params = dict(model.named_parameters())
    by_name, off = {}, 0
    for order in sorted(meta):
        info = meta[order]
        cnt = int(np.prod(info["shape"])) if len(info["shape"]) else 1
        by_name[info["name"]] = (off, cnt, info["dtype"])
        off += cnt * np.dtype(info["dtype"]).itemsize
    dsts, kinds, counts, offsets, auxs = [], [], [], [], []
    for name, (o, cnt, dt) in by_name.items():
        if name.endswith(".centroids"):
            continue
        if name.endswith(".labeled_weight"):
            target = params[name.replace("labeled_weight", "weight")]
            kinds.append(UNPACK_GATHER_U16 if dt == "uint16" else UNPACK_GATHER_U8)
            auxs.append(by_name[name.replace("labeled_weight", "centroids")][0])
        else:
            target = params[name]
            kinds.append(UNPACK_F16_TO_F32)
            auxs.append(0)
        if target.dtype != torch.float32 or not target.is_contiguous() or target.numel() != cnt:
            raise _lib.SirenB200Error(f"cannot unpack {name} into a {target.dtype} tensor of {target.numel()} elements")
        dsts.append(target.data.data_ptr())
        counts.append(cnt)
        offsets.append(o)
    n = len(dsts)
    _lib.check(lib.sirenb200_pack_stream(
        n, None, (ctypes.c_void_p * n)(*dsts), (ctypes.c_int32 * n)(*kinds), (ctypes.c_int64 * n)(*counts),
        (ctypes.c_int64 * n)(*offsets), (ctypes.c_int64 * n)(*auxs), stream.data_ptr(),
        torch.cuda.current_stream().cuda_stream))
    return model


# ------------------------------------------------------------------------------------------------------------------
# the reference's interface (host files)
# ------------------------------------------------------------------------------------------------------------------
def linear_state_dict(model):
    """:15-41 — for callers that want the tensors: {name: CPU tensor} of the packed representation."""
    stream, meta = pack_state_dict(model)
    host = stream.cpu().numpy()
    out, off = OrderedDict(), 0
    for order in sorted(meta):
        info = meta[order]
        cnt = int(np.prod(info["shape"])) if len(info["shape"]) else 1
        nb = cnt * np.dtype(info["dtype"]).itemsize
        out[info["name"]] = torch.from_numpy(host[off:off + nb].view(info["dtype"]).reshape(info["shape"]).copy())
        off += nb
    return out


class _Plain:
    """The reference's NumpyParser (parsers.py:21-45): raw bytes."""

    def __init__(self, handler):
        self.handler = handler

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def write(self, data):
        return self.handler.write(data)

    def read(self):
        return self.handler.read()

    def flush(self):
        self.handler.flush()


class _Lzma(_Plain):
    def write(self, data):
        return self.handler.write(lzma.compress(bytes(data)))

    def read(self):
        return lzma.decompress(self.handler.read())


def _stream(name, **kwargs):
    if name == "plain":
        return _Plain, _Plain
    if name == "lzma":
        return _Lzma, _Lzma
    if name == "zstd":
        try:
            import zstandard
        except ImportError as e:
            raise _lib.SirenB200Error("stream 'zstd' needs the third-party `zstandard` module (host-side entropy "
                                      "coding is out of scope of the B200 path); use 'plain' or 'lzma'") from e
        return (zstandard.ZstdCompressor(level=kwargs.get("level", 3)).stream_writer,
                zstandard.ZstdDecompressor().stream_reader)
    raise NotImplementedError(f"stream writer {name} not found.")


def compress_state_dict(model, dir_name, stream_name="plain", **kwargs):
    """:70-120 — writes <dir>/compressed_weights.data + meta_data.json; returns the size of the data file."""
    writer, _ = _stream(stream_name, **kwargs)
    stream, meta = pack_state_dict(model)
    payload = stream.cpu().numpy().tobytes()  # ONE device->host copy
    dir_name = Path(dir_name)
    dir_name.mkdir(exist_ok=True, parents=True)
    binary_file = dir_name / "compressed_weights.data"
    with open(binary_file, "wb") as fh:
        with writer(fh) as compressor:
            compressor.write(payload)
            compressor.flush()
    with open(dir_name / "meta_data.json", "w") as f:
        f.write(json.dumps({k: {"shape": v["shape"], "dtype": v["dtype"], "name": v["name"]} for k, v in meta.items()},
                           indent=2, sort_keys=True))
    return binary_file.stat().st_size


def decompress_state_dict(dir_name, stream_name="plain", device=None, **kwargs):
    """:123-186 — {name: fp32 tensor}; quantised layers come back as `weight` = centroids[labels]."""
    _, reader = _stream(stream_name, **kwargs)
    dir_name = Path(dir_name)
    with open(dir_name / "meta_data.json") as f:
        meta = {int(k): v for k, v in json.load(f).items()}
    with open(dir_name / "compressed_weights.data", "rb") as fh:
        with reader(fh) as dec:
            raw = dec.read()
    arrays, off = {}, 0
    for order in sorted(meta):
        info = meta[order]
        cnt = int(np.prod(info["shape"])) if len(info["shape"]) else 1
        arr = np.frombuffer(raw, dtype=getattr(np, info["dtype"]), count=cnt, offset=off).reshape(info["shape"])
        arrays[info["name"]] = arr
        off += cnt * np.dtype(info["dtype"]).itemsize
    out = {}
    for name, arr in arrays.items():
        if "labeled_weight" in name:
            cent = arrays[name.replace("labeled_weight", "centroids")]
            out[name.replace("labeled_weight", "weight")] = torch.from_numpy(cent[arr].copy()).float()
        elif "centroids" not in name:
            out[name] = torch.from_numpy(arr.copy()).float()
    if device is not None:
        out = {k: v.to(device) for k, v in out.items()}
    return out


def load_stream(dir_name, stream_name="plain", device="cuda", **kwargs):
    """File -> (device uint8 stream, meta) for unpack_into_model."""
    _, reader = _stream(stream_name, **kwargs)
    dir_name = Path(dir_name)
    with open(dir_name / "meta_data.json") as f:
        meta = {int(k): v for k, v in json.load(f).items()}
    with open(dir_name / "compressed_weights.data", "rb") as fh:
        with reader(fh) as dec:
            raw = dec.read()
    return torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device), meta

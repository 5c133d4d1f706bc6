"""Read the published HILCodec ONNX graphs without the `onnx` package.

The reference ships its pretrained weights only as ONNX files
(`onnx/hil_{speech,music}_{enc,dec,vq{i},deq{i}}.onnx`, produced by the export
notebook `scripts/HILCodec Onnx.ipynb` cell 5).  The initializer names of the
`enc`/`dec` graphs are the `state_dict()` keys of `models/hilcodec/streaming.py`
`Encoder`/`Decoder` after `remove_weight_reparameterizations()`
(streaming.py:740-747), and every `vq{i}` graph carries `embed[1024,128]`
(streaming.py:46).  ONNX is protobuf; this module walks the wire format by hand
(varint + length-delimited fields only) and pulls out `graph.initializer[]`.

Field numbers used (onnx.proto3):
  ModelProto.graph = 7
  GraphProto.initializer = 5
  TensorProto: dims = 1, data_type = 2, float_data = 4, int64_data = 7,
               name = 8, raw_data = 9
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Iterator, Tuple

import numpy as np

_ONNX_DTYPES = {1: np.float32, 6: np.int32, 7: np.int64, 11: np.float64}


def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    out = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf: memoryview) -> Iterator[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value) for one protobuf message."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            val = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            val = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield field, wt, val


def _packed_varints(buf: memoryview) -> list:
    out, pos = [], 0
    while pos < len(buf):
        v, pos = _varint(buf, pos)
        out.append(v)
    return out


def _tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    dims, dtype, name, raw = [], 1, "", None
    floats, int64s = [], []
    for field, wt, val in _fields(buf):
        if field == 1:
            dims.extend(_packed_varints(val) if wt == 2 else [val])
        elif field == 2:
            dtype = val
        elif field == 4:
            if wt == 2:
                floats.extend(struct.unpack(f"<{len(val) // 4}f", bytes(val)))
            else:
                floats.append(struct.unpack("<f", val)[0])
        elif field == 7:
            int64s.extend(_packed_varints(val) if wt == 2 else [val])
        elif field == 8:
            name = bytes(val).decode("utf-8")
        elif field == 9:
            raw = bytes(val)
    np_dtype = _ONNX_DTYPES.get(dtype)
    if np_dtype is None:
        raise ValueError(f"initializer {name!r}: unsupported ONNX data_type {dtype}")
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(np_dtype).newbyteorder("<")).astype(np_dtype)
    elif floats:
        arr = np.asarray(floats, dtype=np_dtype)
    else:
        arr = np.asarray(int64s, dtype=np_dtype)
    return name, arr.reshape(dims).copy()


def read_initializers(path: str) -> Dict[str, np.ndarray]:
    """Return {initializer name: ndarray} of one .onnx file."""
    with open(path, "rb") as f:
        data = memoryview(f.read())
    out: Dict[str, np.ndarray] = {}
    for field, wt, val in _fields(data):
        if field == 7 and wt == 2:  # ModelProto.graph
            for gfield, gwt, gval in _fields(val):
                if gfield == 5 and gwt == 2:  # GraphProto.initializer
                    name, arr = _tensor(gval)
                    out[name] = arr
    return out


def load_onnx_model(onnx_dir: str, name: str) -> Dict[str, np.ndarray]:
    """Collect every tensor of one published model ("hil_speech" / "hil_music").

    Keys: `encoder.<streaming Encoder state_dict key>`,
    `decoder.<streaming Decoder state_dict key>`, `quantizer.layers.{i}.embed`.
    The dequantizer codebooks (`*_deq{i}.onnx`) are byte-identical to the
    quantizer ones and are checked, not stored twice.
    """
    out: Dict[str, np.ndarray] = {}
    for part, prefix in (("enc", "encoder."), ("dec", "decoder.")):
        inits = read_initializers(os.path.join(onnx_dir, f"{name}_{part}.onnx"))
        for k, v in inits.items():
            if v.dtype != np.float32 or k.startswith("onnx::") or "/" in k:
                continue  # shape constants / folded scalars, not state_dict tensors
            out[prefix + k] = v
    i = 0
    while os.path.exists(os.path.join(onnx_dir, f"{name}_vq{i}.onnx")):
        inits = read_initializers(os.path.join(onnx_dir, f"{name}_vq{i}.onnx"))
        embed = inits["embed"] if "embed" in inits else next(
            v for v in inits.values() if v.shape == (1024, 128))
        deq_path = os.path.join(onnx_dir, f"{name}_deq{i}.onnx")
        if os.path.exists(deq_path):
            deq = read_initializers(deq_path)
            demb = deq["embed"] if "embed" in deq else next(
                v for v in deq.values() if v.ndim == 2 and v.shape[1] == embed.shape[1])
            if not np.array_equal(demb, embed):
                raise ValueError(f"{name}: vq{i} and deq{i} codebooks differ")
        out[f"quantizer.layers.{i}.embed"] = embed.astype(np.float32)
        i += 1
    if i == 0:
        raise FileNotFoundError(f"no {name}_vq*.onnx under {onnx_dir}")
    return out


def convert(onnx_dir: str, name: str, out_path: str) -> Dict[str, np.ndarray]:
    """Extract one published model to an .npz the loader in `weights.py` reads."""
    tensors = load_onnx_model(onnx_dir, name)
    os.makedirs(os.path.dirname(os.path.abspath(out_path)), exist_ok=True)
    np.savez(out_path, **tensors)
    return tensors


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description=__doc__.splitlines()[0])
    ap.add_argument("onnx_dir")
    ap.add_argument("name", choices=["hil_speech", "hil_music"])
    ap.add_argument("out")
    a = ap.parse_args()
    t = convert(a.onnx_dir, a.name, a.out)
    print(f"{a.name}: {len(t)} tensors, {sum(v.size for v in t.values())} values -> {a.out}")

"""Drop-in replacements for the reference's deployment modules
(`models/hilcodec/streaming.py`): `Encoder` (:368), `Decoder` (:520), `ResidualVQ` (:75),
`Dequantizer` (:134) and `HILCodec` (:651) -- same constructor keywords, `forward()`
signatures, cache-list protocol and `state_dict()` key names, with every forward executed
by the sm_100a kernels in libhilcodec_b200.so through the C ABI (include/hilcodec_b200.h).

Differences kept deliberately small and documented:
  * weights are held *folded* (what `remove_weight_reparameterizations()` leaves behind,
    i.e. the tensors in the published ONNX files); a training-format state dict with
    `weight_g` / `weight_v` / `res_scale_param` / `scale_param` is folded on load
    (see `fold.py`);
  * tensors must live on a CUDA device -- there is no CPU path, calls with CPU tensors raise;
  * `HILCodec.forward` chains encoder -> quantizer -> dequantizer -> decoder (the reference's
    own `forward`, streaming.py:726-738, feeds indices to the decoder and cannot run).
"""
from __future__ import annotations

import ctypes as C
import math
import typing as tp
from collections import OrderedDict

import numpy as np
import torch
from torch import Tensor, nn

from . import _lib
from .weights import CodecConfig, check_weights, dft_basis, tensor_shapes


def _require_cuda(t: Tensor, what: str, dtype: torch.dtype = torch.float32) -> None:
    """The kernels reinterpret `data_ptr()` as `dtype`: anything else is rejected here (float32 for waveforms,
    latents and caches, int64 for indices only)."""
    if not isinstance(t, Tensor):
        raise TypeError(f"hilcodec_b200: {what} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(
            f"hilcodec_b200: {what} must be a CUDA tensor (this package has no CPU path; "
            f"move the module and its inputs to a B200 with .cuda())")
    if t.dtype != dtype:
        raise TypeError(f"hilcodec_b200: {what} must be {dtype}, got {t.dtype}")


def _check_wav(x: Tensor, hop: int) -> tp.Tuple[int, int]:
    _require_cuda(x, "x")
    if x.dim() != 3 or x.shape[1] != 1:
        raise ValueError(f"expected x of shape [B, 1, T], got {tuple(x.shape)}")
    B, _, T = x.shape
    if T % hop or T == 0:
        raise ValueError(f"T={T} must be a positive multiple of the hop length {hop}")
    return B, T


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _NativeCodec:
    """Owns one `hil_model` (+ per-batch `hil_state`s) per CUDA device for a weight set."""

    def __init__(self, cfg: CodecConfig, graph: int = _lib.HIL_GRAPH_DEPLOY):
        self.cfg = cfg
        self.graph = graph  # HIL_GRAPH_DEPLOY (streaming.py) or HIL_GRAPH_TRAIN (models.py / seanet.py)
        self.weights: "OrderedDict[str, Tensor]" = OrderedDict()  # folded fp32 CPU tensors
        self._models: tp.Dict[int, int] = {}
        self._states: tp.Dict[tp.Tuple[int, int], int] = {}
        self._lib_handle = None
        # bumped by invalidate(): a StreamState remembers the generation it was created in and refuses to run
        # against a later one (its hil_state belongs to a model that has been replaced)
        self.generation = 0
        # fp16-range guard (include/hilcodec_b200.h, hil_state_range_flag): after every module call read the state's
        # "non-finite output" flag (one 4-byte D2H + stream sync) and, if set, repeat the call on the FP32 kernels.  The
        # reference's fp32 convolutions have no |x| < 65504 limit; with this on, neither do the modules.  Set False
        # to keep calls asynchronous (latency-critical loops); StreamState.step never checks (see range_overflow()).
        self.check_range = True

    @property
    def _lib(self):
        if self._lib_handle is None:
            self._lib_handle = _lib.load()
        return self._lib_handle

    # -- weights ---------------------------------------------------------------------
    def set_weights(self, w: tp.Mapping[str, tp.Any]) -> None:
        for k, v in w.items():
            t = torch.as_tensor(np.asarray(v) if not isinstance(v, Tensor) else v).detach()
            self.weights[k] = t.to(device="cpu", dtype=torch.float32).contiguous().clone()
        self.invalidate()

    def set_graph(self, graph: int) -> None:
        if graph not in (_lib.HIL_GRAPH_DEPLOY, _lib.HIL_GRAPH_TRAIN):
            raise ValueError(f"Unknown graph: {graph}")
        if graph != self.graph:
            self.graph = graph
            self.invalidate()

    def invalidate(self) -> None:
        """Drop the native models (weights or graph changed).  The C side defers freeing a model until the last
        `hil_state` created from it is destroyed, so StreamStates that are still alive stay safe to reset /
        export / destroy; they raise on `step` / `codec_forward(state=...)` (see StreamState._check)."""
        self.generation += 1
        lib = self._lib_handle
        if lib is not None:
            for s in self._states.values():
                lib.hil_state_destroy(s)
            for m in self._models.values():
                lib.hil_model_destroy(m)
        self._states.clear()
        self._models.clear()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.invalidate()
        except Exception:
            pass

    # -- native handles ----------------------------------------------------------------
    def c_config(self) -> _lib.HilConfig:
        cfg = self.cfg
        c = _lib.HilConfig()
        c.channels_enc, c.channels_dec, c.n_fft_base = cfg.channels_enc, cfg.channels_dec, cfg.n_fft_base
        c.n_residual_enc, c.n_residual_dec = cfg.n_residual_enc, cfg.n_residual_dec
        c.res_scale_enc, c.res_scale_dec = cfg.res_scale_enc, cfg.res_scale_dec
        c.n_strides = len(cfg.strides)
        for i, r in enumerate(cfg.strides):
            c.strides[i] = r
        c.kernel_size, c.dim, c.codebook_size = cfg.kernel_size, cfg.dim, cfg.codebook_size
        c.num_quantizers = cfg.num_quantizers
        return c

    def model(self, device: torch.device) -> int:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = self._models.get(idx)
        if h is not None:
            return h
        lib = self._lib
        with torch.cuda.device(idx):
            handle = C.c_void_p()
            cfg = self.c_config()
            _lib.check(lib.hil_model_create(C.byref(cfg), C.byref(handle)))
            try:
                shapes = tensor_shapes(self.cfg)
                complete = [pre for pre in ("encoder.", "decoder.", "quantizer.")
                            if all(k in self.weights for k in shapes if k.startswith(pre))]
                if not complete:
                    raise RuntimeError("hilcodec_b200: no complete weight section loaded (encoder/decoder/quantizer)")
                for name, t in self.weights.items():
                    if not name.startswith(tuple(complete)):
                        continue
                    dims = (C.c_int64 * t.dim())(*t.shape)
                    _lib.check(lib.hil_model_set_tensor(handle, name.encode(), C.c_void_p(t.data_ptr()), dims, t.dim()))
                _lib.check(lib.hil_model_set_graph(handle, self.graph))
                _lib.check(lib.hil_model_finalize(handle))
            except Exception:
                lib.hil_model_destroy(handle)
                raise
        self._models[idx] = handle.value
        return handle.value

    def state(self, device: torch.device, batch: int) -> int:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = (idx, batch)
        s = self._states.get(key)
        if s is not None:
            return s
        m = self.model(device)
        with torch.cuda.device(idx):
            handle = C.c_void_p()
            _lib.check(self._lib.hil_state_create(m, batch, C.byref(handle)))
        self._states[key] = handle.value
        return handle.value

    def cache_shapes(self, which: int, batch: int, device: torch.device) -> tp.List[tp.Tuple[int, int, int]]:
        m = self.model(device)
        n = self._lib.hil_model_num_caches(m, which)
        out = []
        dims = (C.c_int64 * 3)()
        for i in range(n):
            _lib.check(self._lib.hil_model_cache_shape(m, which, i, batch, dims))
            out.append((int(dims[0]), int(dims[1]), int(dims[2])))
        return out

    # -- calls ---------------------------------------------------------------------------
    def _guarded(self, state: int, dev: torch.device, call: tp.Callable[[], None],
                 rollback: tp.Tuple[int, int] = (0, 0)) -> None:
        """Run `call` (a closure over one C-ABI forward); if the range guard trips, undo the cache-generation advance
        (`rollback` = (encoder, decoder) sides the call advances inside the state) and repeat it on the FP32 kernels."""
        call()
        if not self.check_range:
            return
        lib = self._lib
        flag = C.c_int32(0)
        _lib.check(lib.hil_state_range_flag(state, 1, _stream_ptr(dev), C.byref(flag)))
        if not flag.value:
            return
        if any(rollback):
            _lib.check(lib.hil_state_rollback(state, *rollback))
        prev = lib.hil_set_exact_fp32(1)
        try:
            call()
            _lib.check(lib.hil_state_range_flag(state, 1, _stream_ptr(dev), C.byref(flag)))
        finally:
            lib.hil_set_exact_fp32(prev)

    def _cache_io(self, which: int, caches: tp.Sequence[Tensor], batch: int, device: torch.device):
        shapes = self.cache_shapes(which, batch, device)
        if len(caches) != len(shapes):
            raise ValueError(f"expected {len(shapes)} cache tensors, got {len(caches)}")
        ins = []
        for c, shp in zip(caches, shapes):
            _require_cuda(c, "cache")
            if tuple(c.shape) != shp:
                raise ValueError(f"cache shape {tuple(c.shape)} != {shp}")
            ins.append(c.contiguous())
        sizes = [s[0] * s[1] * s[2] for s in shapes]
        offs = np.cumsum([0] + [(n + 3) // 4 * 4 for n in sizes])
        flat = torch.empty(int(offs[-1]), dtype=torch.float32, device=device)
        outs = [flat[int(o):int(o) + n].view(shp) for o, n, shp in zip(offs[:-1], sizes, shapes)]
        n = len(shapes)
        pin = (C.c_void_p * n)(*[t.data_ptr() for t in ins])
        pout = (C.c_void_p * n)(*[t.data_ptr() for t in outs])
        return ins, outs, pin, pout

    def encode(self, x: Tensor, caches: tp.Sequence[Tensor]) -> tp.Tuple[Tensor, tp.List[Tensor]]:
        _require_cuda(x, "x")
        if x.dim() != 3 or x.shape[1] != 1:
            raise ValueError(f"expected x of shape [B, 1, T], got {tuple(x.shape)}")
        B, _, T = x.shape
        hop = self.cfg.hop
        if T % hop or T == 0:
            raise ValueError(f"T={T} must be a positive multiple of the hop length {hop}")
        dev = x.device
        with torch.cuda.device(dev):
            x = x.contiguous()
            ins, outs, pin, pout = self._cache_io(_lib.HIL_ENCODER, caches, B, dev)
            z = torch.empty(B, T // hop, self.cfg.dim, dtype=torch.float32, device=dev)
            model, state = self.model(dev), self.state(dev, B)
            self._guarded(state, dev, lambda: _lib.check(self._lib.hil_encode_caches(
                model, state, x.data_ptr(), B, T, z.data_ptr(), pin, pout, _stream_ptr(dev))))
        return z, outs

    def encode_ragged(self, x: Tensor) -> Tensor:
        """One-shot training-graph encoder call for any T >= 1 (`hil_encode_ragged`): [B,1,T] -> [B,ceil(T/hop),dim]."""
        _require_cuda(x, "x")
        if x.dim() != 3 or x.shape[1] != 1:
            raise ValueError(f"expected x of shape [B, 1, T], got {tuple(x.shape)}")
        B, _, T = x.shape
        if T == 0:
            raise ValueError("empty input")
        dev = x.device
        with torch.cuda.device(dev):
            x = x.contiguous()
            z = torch.empty(B, -(-T // self.cfg.hop), self.cfg.dim, dtype=torch.float32, device=dev)
            model, state = self.model(dev), self.state(dev, B)
            self._guarded(state, dev, lambda: _lib.check(self._lib.hil_encode_ragged(
                model, state, x.data_ptr(), B, T, z.data_ptr(), _stream_ptr(dev))))
        return z

    def decode(self, q: Tensor, caches: tp.Sequence[Tensor]) -> tp.Tuple[Tensor, tp.List[Tensor]]:
        _require_cuda(q, "x")
        if q.dim() != 3 or q.shape[2] != self.cfg.dim:
            raise ValueError(f"expected x of shape [B, T', {self.cfg.dim}], got {tuple(q.shape)}")
        B, F, _ = q.shape
        if F == 0:
            raise ValueError("empty input")
        dev = q.device
        with torch.cuda.device(dev):
            q = q.contiguous()
            ins, outs, pin, pout = self._cache_io(_lib.HIL_DECODER, caches, B, dev)
            wav = torch.empty(B, 1, F * self.cfg.hop, dtype=torch.float32, device=dev)
            model, state = self.model(dev), self.state(dev, B)
            self._guarded(state, dev, lambda: _lib.check(self._lib.hil_decode_caches(
                model, state, q.data_ptr(), B, F, wav.data_ptr(), pin, pout, _stream_ptr(dev))))
        return wav, outs

    def rvq_encode(self, z: Tensor, n: int, with_sum: bool = False):
        _require_cuda(z, "x")
        if z.dim() != 3 or z.shape[2] != self.cfg.dim:
            raise ValueError(f"expected x of shape [B, T, {self.cfg.dim}], got {tuple(z.shape)}")
        assert 1 <= n <= self.cfg.num_quantizers, "n must satisfy 1 <= n <= num_quantizers"
        B, F, _ = z.shape
        dev = z.device
        with torch.cuda.device(dev):
            z = z.contiguous()
            idx = torch.empty(n, B, F, dtype=torch.int64, device=dev)
            qsum = torch.empty_like(z) if with_sum else None
            _lib.check(self._lib.hil_rvq_encode(
                self.model(dev), z.data_ptr(), B, F, n, idx.data_ptr(),
                qsum.data_ptr() if with_sum else None, _stream_ptr(dev)))
        return (idx, qsum) if with_sum else idx

    def rvq_decode(self, idx: Tensor, n: int) -> Tensor:
        _require_cuda(idx, "indices", torch.int64)
        if idx.dim() != 3:
            raise TypeError("indices must be an int64 tensor of shape [n, B, T]")
        assert 1 <= n <= self.cfg.num_quantizers, "n must satisfy 1 <= n <= num_quantizers"
        if idx.shape[0] < n:
            raise IndexError(f"indices has {idx.shape[0]} stages, n={n}")
        _, B, F = idx.shape
        dev = idx.device
        with torch.cuda.device(dev):
            idx = idx[:n].contiguous()
            q = torch.empty(B, F, self.cfg.dim, dtype=torch.float32, device=dev)
            _lib.check(self._lib.hil_rvq_decode(self.model(dev), idx.data_ptr(), B, F, n, q.data_ptr(), _stream_ptr(dev)))
        return q

    def zero_caches(self, which: int, batch: int, device: torch.device) -> tp.List[Tensor]:
        return [torch.zeros(s, dtype=torch.float32, device=device) for s in self.cache_shapes(which, batch, device)]


class _Part(nn.Module):
    """Base of the drop-in modules: a view on a shared `_NativeCodec` under a key prefix."""

    _prefix = ""

    def __init__(self, core: _NativeCodec):
        super().__init__()
        object.__setattr__(self, "_core", core)
        self._device = torch.device("cpu")

    # state_dict protocol with the reference's key names ------------------------------
    def _own_keys(self) -> tp.List[str]:
        return [k for k in tensor_shapes(self._core.cfg) if k.startswith(self._prefix)]

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):  # type: ignore[override]
        out = destination if destination is not None else OrderedDict()
        for k in self._own_keys():
            if k in self._core.weights:
                out[prefix + k[len(self._prefix):]] = self._core.weights[k]
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):  # type: ignore[override]
        from .fold import fold_state_dict

        sd = fold_state_dict(OrderedDict(state_dict), self._core.cfg, part=self._prefix.rstrip("."))
        shapes = tensor_shapes(self._core.cfg)
        expected = set(self._own_keys())
        got: tp.Dict[str, Tensor] = {}
        unexpected = []
        for k, v in sd.items():
            full = self._prefix + k
            if full in expected:
                if tuple(v.shape) != shapes[full]:
                    raise RuntimeError(f"size mismatch for {k}: {tuple(v.shape)} vs {shapes[full]}")
                got[full] = v
            elif not k.endswith("ema_num"):
                unexpected.append(k)
        missing = [k[len(self._prefix):] for k in expected if k not in got and k not in self._core.weights]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing}, unexpected {unexpected}")
        self._core.set_weights(got)
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def _apply(self, fn, recurse=True):  # tracks .cuda()/.to(); weights are uploaded lazily per device
        probe = fn(torch.empty(0))
        self._device = probe.device
        return super()._apply(fn, recurse)


class Encoder(_Part):
    """streaming.py:368 `Encoder`: forward(x[B,1,T], *caches22) -> (z[B,T/hop,dim], caches22)."""

    _prefix = "encoder."

    def __init__(self, core_or_channels: tp.Union[_NativeCodec, int] = 1, dimension: int = 128, n_filters: int = 32,
                 n_fft_base: int = 64, n_residual_layers: int = 2, ratios: tp.List[int] = [8, 5, 4, 2],
                 activation: str = "ELU", activation_params: dict = {"alpha": 1.0}, norm: str = "weight_norm",
                 kernel_size: int = 5, last_kernel_size: int = 5, residual_kernel_size: int = 5,
                 dilation_base: int = 1, skip: str = "1x1", compress: int = 1, act_all: bool = False,
                 expansion: int = 1, groups: int = -1, l2norm: bool = True, bias: bool = True,
                 res_scale: float = 0.5, **_ignored):
        if isinstance(core_or_channels, _NativeCodec):
            core = core_or_channels
        else:
            _check_supported(activation, activation_params, norm, kernel_size, last_kernel_size, residual_kernel_size,
                             dilation_base, compress, act_all, expansion, groups, bias)
            if not l2norm:
                raise NotImplementedError("hilcodec_b200 builds the encoder_l2norm=True graph of both published configs")
            core = _NativeCodec(CodecConfig(channels_enc=n_filters, n_fft_base=n_fft_base,
                                            n_residual_enc=n_residual_layers, res_scale_enc=res_scale,
                                            strides=tuple(ratios), kernel_size=kernel_size, dim=dimension))
        super().__init__(core)
        cfg = core.cfg
        self.dimension = cfg.dim
        self.n_filters = cfg.channels_enc
        self.ratios = list(reversed(cfg.strides))
        self.n_residual_layers = cfg.n_residual_enc
        self.hop_length = int(np.prod(self.ratios))
        self.num_cache = 2 + len(cfg.strides) * (2 * cfg.n_residual_enc + 1)
        self.merged = True

    def initialize_cache(self, x: Tensor) -> tp.List[Tensor]:
        _require_cuda(x, "x")
        return self._core.zero_caches(_lib.HIL_ENCODER, x.size(0), x.device)

    def merge_scaling(self) -> None:
        return None  # weights are stored folded

    def forward(self, x: Tensor, *args) -> tp.Tuple[Tensor, tp.List[Tensor]]:
        return self._core.encode(x, list(args))


class Decoder(_Part):
    """streaming.py:520 `Decoder`: forward(q[B,T',dim], *caches30) -> (wav[B,1,hop*T'], caches30)."""

    _prefix = "decoder."

    def __init__(self, core_or_channels: tp.Union[_NativeCodec, int] = 1, dimension: int = 128, n_filters: int = 32,
                 n_residual_layers: int = 1, ratios: tp.List[int] = [8, 5, 4, 2], activation: str = "ELU",
                 activation_params: dict = {"alpha": 1.0}, norm: str = "weight_norm", kernel_size: int = 7,
                 last_kernel_size: int = 7, residual_kernel_size: int = 3, dilation_base: int = 2, skip: str = "1x1",
                 compress: int = 2, final_activation: tp.Optional[str] = None,
                 final_activation_params: tp.Optional[dict] = None, act_all: bool = False, expansion: int = 1,
                 groups: int = -1, bias: bool = True, res_scale: tp.Optional[float] = None, **_ignored):
        if isinstance(core_or_channels, _NativeCodec):
            core = core_or_channels
        else:
            _check_supported(activation, activation_params, norm, kernel_size, last_kernel_size, residual_kernel_size,
                             dilation_base, compress, act_all, expansion, groups, bias)
            if final_activation != "Tanh":
                raise NotImplementedError("hilcodec_b200 builds the final_activation='Tanh' graph of both published configs")
            core = _NativeCodec(CodecConfig(channels_dec=n_filters, n_residual_dec=n_residual_layers,
                                            res_scale_dec=res_scale, strides=tuple(ratios), kernel_size=kernel_size,
                                            dim=dimension))
        super().__init__(core)
        cfg = core.cfg
        self.dimension = cfg.dim
        self.channels = 1
        self.n_filters = cfg.channels_dec
        self.ratios = list(cfg.strides)
        self.n_residual_layers = cfg.n_residual_dec
        self.hop_length = int(np.prod(self.ratios))
        self.num_cache = 2 + len(cfg.strides) * (2 * cfg.n_residual_dec + 1)
        self.merged = True

    def initialize_cache(self, x: Tensor) -> tp.List[Tensor]:
        _require_cuda(x, "x")
        return self._core.zero_caches(_lib.HIL_DECODER, x.size(0), x.device)

    def merge_scaling(self) -> None:
        return None

    def forward(self, x: Tensor, *args) -> tp.Tuple[Tensor, tp.List[Tensor]]:
        return self._core.decode(x, list(args))


class _Layers:
    """`quantizer.layers` look-alike: len() and `[i].embed`."""

    def __init__(self, core: _NativeCodec):
        self._core = core

    def __len__(self) -> int:
        return self._core.cfg.num_quantizers

    def __getitem__(self, i: int):
        core = self._core

        class _Layer:
            @property
            def embed(self_inner) -> Tensor:
                return core.weights[f"quantizer.layers.{i}.embed"]
        if not -len(self) <= i < len(self):
            raise IndexError(i)
        return _Layer()


class ResidualVQ(_Part):
    """streaming.py:75 `ResidualVQ`: forward(x[B,T,C], n) -> indices[n,B,T] int64."""

    _prefix = "quantizer."

    def __init__(self, core: tp.Optional[_NativeCodec] = None, num_quantizers: int = 16, dropout: bool = False,
                 dropout_index: tp.Optional[tp.List[int]] = None, dim: int = 128, codebook_size: int = 1024, **_ignored):
        if not isinstance(core, _NativeCodec):
            core = _NativeCodec(CodecConfig(dim=dim, codebook_size=codebook_size, num_quantizers=num_quantizers))
        super().__init__(core)
        self.layers = _Layers(core)

    def forward(self, x: Tensor, n: int) -> Tensor:
        return self._core.rvq_encode(x, n)

    def quantize(self, x: Tensor, n: int) -> tp.Tuple[Tensor, Tensor]:
        """indices and the dequantised sum in one kernel (what encoder->decoder chaining needs)."""
        return self._core.rvq_encode(x, n, with_sum=True)


class Dequantizer(ResidualVQ):
    """streaming.py:134 `Dequantizer`: forward(indices[n,B,T], n) -> q[B,T,C]."""

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):  # keys are layers.{i}.embed too
        return super().state_dict(*args, destination=destination, prefix=prefix, keep_vars=keep_vars)

    def forward(self, indices: Tensor, n: int) -> Tensor:  # type: ignore[override]
        return self._core.rvq_decode(indices, n)


def _check_supported(activation, activation_params, norm, kernel_size, last_kernel_size, residual_kernel_size,
                     dilation_base, compress, act_all, expansion, groups, bias) -> None:
    if norm != "weight_norm":
        raise ValueError(f"Unknown norm: {norm}")  # causal_layers.py:200-203
    bad = []
    if activation != "ELU" or float(activation_params.get("alpha", 1.0)) != 1.0:
        bad.append("activation")
    if not (kernel_size == last_kernel_size == residual_kernel_size == 5):
        bad.append("kernel sizes (5)")
    if dilation_base != 1:
        bad.append("dilation_base (1)")
    if compress != 1 or act_all or expansion != 1 or groups != -1 or not bias:
        bad.append("compress/act_all/expansion/groups/bias")
    if bad:
        raise NotImplementedError(
            "hilcodec_b200 builds the graph of configs/hilcodec_{speech,music}.yaml; unsupported: " + ", ".join(bad))


class HILCodec(nn.Module):
    """streaming.py:651 `HILCodec`, same constructor keywords."""

    def __init__(self, sample_rate: int = 16_000, channels_audio: int = 1, channels_enc: int = 64,
                 channels_dec: int = 96, n_fft_base: int = 64, n_residual_enc: int = 2, n_residual_dec: int = 3,
                 res_scale_enc: tp.Optional[float] = 0.5773502691896258,
                 res_scale_dec: tp.Optional[float] = 0.5773502691896258, strides: tp.List[int] = [8, 5, 4, 2],
                 activation: str = "ELU", activation_kwargs: dict = {"alpha": 1.0}, norm: str = "weight_norm",
                 kernel_size: int = 5, last_kernel_size: int = 5, residual_kernel_size: int = 5,
                 dilation_base: int = 1, skip: str = "identity", compress: int = 1,
                 final_activation: tp.Optional[str] = "Tanh", use_vq: bool = True, vq: str = "ResidualVQ",
                 vq_kwargs: tp.Dict[str, tp.Any] = dict(dim=128), act_all: bool = False, expansion: int = 1,
                 groups: int = -1, encoder_l2norm: bool = True, bias: bool = True, spec: str = "stft",
                 spec_compression: str = "log", zero_init: bool = True, inout_norm: bool = True, **_ignored):
        assert spec == "stft"
        assert spec_compression == "log"
        assert skip == "identity", skip
        assert zero_init is True
        assert inout_norm is True
        if expansion != 1 and groups != -1:
            raise RuntimeError(f"Both expansion({expansion}) and groups({groups}) are set. "
                               f"Either set expansion=1 or set groups=-1")
        if channels_audio != 1 or vq != "ResidualVQ" or not encoder_l2norm or final_activation != "Tanh":
            raise NotImplementedError("hilcodec_b200 builds the mono ResidualVQ/l2norm/Tanh graph of both published configs")
        _check_supported(activation, activation_kwargs, norm, kernel_size, last_kernel_size, residual_kernel_size,
                         dilation_base, compress, act_all, expansion, groups, bias)
        super().__init__()
        cfg = CodecConfig(
            channels_enc=channels_enc, channels_dec=channels_dec, n_fft_base=n_fft_base,
            n_residual_enc=n_residual_enc, n_residual_dec=n_residual_dec, res_scale_enc=res_scale_enc,
            res_scale_dec=res_scale_dec, strides=tuple(strides), kernel_size=kernel_size,
            dim=vq_kwargs.get("dim", 128), codebook_size=vq_kwargs.get("codebook_size", 1024),
            num_quantizers=vq_kwargs.get("num_quantizers", 16))
        self.cfg = cfg
        self.norm = norm
        core = _NativeCodec(cfg)
        object.__setattr__(self, "_core", core)
        # fixed (non-learned) DFT bases are known at construction, like the reference's registered buffer
        dft = {k: dft_basis(s[2]) for k, s in tensor_shapes(cfg).items() if k.endswith("spec.weight")}
        core.set_weights(dft)
        self.encoder = Encoder(core)
        self.decoder = Decoder(core)
        self.quantizer = ResidualVQ(core)
        self.dequantizer = Dequantizer(core)
        self.sample_rate = sample_rate
        self.channels = channels_audio

    # -- weights -------------------------------------------------------------------------
    @classmethod
    def from_weights(cls, weights: tp.Mapping[str, tp.Any], num_quantizers: int, sample_rate: int = 24_000,
                     **kwargs) -> "HILCodec":
        model = cls(sample_rate, vq_kwargs=dict(dim=128, codebook_size=1024, num_quantizers=num_quantizers), **kwargs)
        check_weights(model.cfg, weights)
        model._core.set_weights(weights)
        return model

    @classmethod
    def from_pretrained(cls, name: str) -> "HILCodec":
        """`hil_speech` / `hil_music` from weights/<name>.npz (extracted from the published ONNX files)."""
        from .weights import CONFIGS, load_pretrained
        return cls.from_weights(load_pretrained(name), CONFIGS[name].num_quantizers)

    def state_dict(self, *args, destination=None, prefix="", keep_vars=False):  # type: ignore[override]
        out = destination if destination is not None else OrderedDict()
        for k, v in self._core.weights.items():
            out[prefix + k] = v
            if k.startswith("quantizer."):
                out[prefix + "de" + k] = v
        return out

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):  # type: ignore[override]
        from .fold import fold_state_dict

        sd = fold_state_dict(OrderedDict(state_dict), self.cfg, part="")
        shapes = tensor_shapes(self.cfg)
        got, unexpected = {}, []
        for k, v in sd.items():
            kk = k[2:] if k.startswith("dequantizer.") else k
            if kk in shapes:
                if tuple(v.shape) != shapes[kk]:
                    raise RuntimeError(f"size mismatch for {k}: {tuple(v.shape)} vs {shapes[kk]}")
                got[kk] = v
            elif not k.endswith("ema_num"):
                unexpected.append(k)
        missing = [k for k in shapes if k not in got and k not in self._core.weights]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing}, unexpected {unexpected}")
        self._core.set_weights(got)
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def remove_weight_reparameterizations(self) -> None:
        """streaming.py:740-747.  Weights are folded when loaded, so nothing is left to do."""
        return None

    # -- forward ---------------------------------------------------------------------------
    def initialize_cache(self, x: Tensor) -> tp.Tuple[tp.List[Tensor], tp.List[Tensor]]:
        return self.encoder.initialize_cache(x), self.decoder.initialize_cache(x)

    def forward(self, x: Tensor, n: int, *args) -> tp.Tuple[Tensor, tp.List[Tensor], tp.List[Tensor]]:
        cache = [*args]
        cache_enc = cache[:self.encoder.num_cache]
        cache_dec = cache[self.encoder.num_cache:]
        z, cache_enc = self.encoder(x, *cache_enc)
        _, q = self.quantizer.quantize(z, n)
        y, cache_dec = self.decoder(q, *cache_dec)
        return y, cache_enc, cache_dec

    @torch.no_grad()
    def codec_forward(self, x: Tensor, n: int, state: tp.Optional["StreamState"] = None):
        """Fused enc -> RVQ -> dec in one C-ABI call (`hil_codec_forward`).  Returns
        (indices[n,B,F], wav[B,1,T]).  With `state` the caches persist on the GPU between calls
        (streaming); without it the call is one-shot from zero caches."""
        B, T = _check_wav(x, self.cfg.hop)
        assert 1 <= n <= self.cfg.num_quantizers, "n must satisfy 1 <= n <= num_quantizers"
        core, dev = self._core, x.device
        with torch.cuda.device(dev):
            x = x.contiguous()
            lib = _lib.load()
            model = core.model(dev)
            if state is None:
                st = core.state(dev, B)
                _lib.check(lib.hil_state_reset(st, _stream_ptr(dev)))
            else:
                if state._core is not core:
                    raise ValueError("state belongs to another model")
                st = state.handle(B)
            idx = torch.empty(n, B, T // self.cfg.hop, dtype=torch.int64, device=dev)
            y = torch.empty(B, 1, T, dtype=torch.float32, device=dev)
            core._guarded(st, dev, lambda: _lib.check(lib.hil_codec_forward(
                model, st, x.data_ptr(), B, T, n, None, idx.data_ptr(), y.data_ptr(), _stream_ptr(dev))), rollback=(1, 1))
        return idx, y

    def new_stream_state(self, batch: int, device: tp.Union[str, torch.device] = "cuda") -> "StreamState":
        return StreamState(self, batch, torch.device(device))


class StreamState:
    """GPU-resident caches of `batch` parallel streams (an opaque `hil_state`), with lossless
    export/import to the reference's list-of-tensors protocol."""

    def __init__(self, model: HILCodec, batch: int, device: torch.device):
        if device.type != "cuda":
            raise RuntimeError("hilcodec_b200: StreamState needs a CUDA device")
        self._core = model._core
        self.batch = batch
        self.device = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        lib = _lib.load()
        with torch.cuda.device(self.device):
            h = C.c_void_p()
            _lib.check(lib.hil_state_create(self._core.model(self.device), batch, C.byref(h)))
        self._h = h.value
        self._lib = lib
        self._generation = self._core.generation
        self._io: tp.Dict[tp.Tuple[int, int], tp.Tuple[Tensor, Tensor, Tensor]] = {}

    def _check(self) -> None:
        if self._generation != self._core.generation:
            raise RuntimeError(
                "hilcodec_b200: the model's weights or graph changed after this StreamState was created "
                "(set_weights / load_state_dict / set_graph); create a new one with model.new_stream_state()")

    def handle(self, batch: int) -> int:
        self._check()
        if batch != self.batch:
            raise ValueError(f"state was created for batch {self.batch}, got {batch}")
        return self._h

    def reset(self) -> None:
        self._check()
        _lib.check(self._lib.hil_state_reset(self._h, _stream_ptr(self.device)))

    @torch.no_grad()
    def step(self, x: Tensor, n: int) -> tp.Tuple[Tensor, Tensor]:
        """One streaming chunk through the CUDA-graph executor (`hil_codec_forward_graph`).  The chunk is copied
        into a buffer owned by this state and the results are returned as views of state-owned buffers that the
        next `step` overwrites (clone them to keep them): fixed pointers are what lets the graph be replayed."""
        self._check()
        hop = self._core.cfg.hop
        B, T = _check_wav(x, hop)
        if B != self.batch:
            raise ValueError(f"state was created for batch {self.batch}, got {B}")
        assert 1 <= n <= self._core.cfg.num_quantizers, "n must satisfy 1 <= n <= num_quantizers"
        key = (T, n)
        bufs = self._io.get(key)
        if bufs is None:
            bufs = (torch.empty(B, 1, T, dtype=torch.float32, device=self.device),
                    torch.empty(n, B, T // hop, dtype=torch.int64, device=self.device),
                    torch.empty(B, 1, T, dtype=torch.float32, device=self.device))
            self._io[key] = bufs
        xin, idx, y = bufs
        with torch.cuda.device(self.device):
            xin.copy_(x)
            _lib.check(self._lib.hil_codec_forward_graph(self._core.model(self.device), self._h, xin.data_ptr(), B, T, n,
                                                         idx.data_ptr(), y.data_ptr(), _stream_ptr(self.device)))
        return idx, y

    def range_overflow(self, clear: bool = True) -> bool:
        """True if any step since the last (cleared) check produced a non-finite latent or PCM sample -- an activation
        left the fp16 range of the tensor-core kernels (`hil_state_range_flag`; synchronises the stream).  `step` does not
        check by itself: a streaming loop stays asynchronous and polls this when it wants to."""
        self._check()
        flag = C.c_int32(0)
        _lib.check(self._lib.hil_state_range_flag(self._h, 1 if clear else 0, _stream_ptr(self.device), C.byref(flag)))
        return bool(flag.value)

    def export(self) -> tp.Tuple[tp.List[Tensor], tp.List[Tensor]]:
        self._check()
        out = []
        for which in (_lib.HIL_ENCODER, _lib.HIL_DECODER):
            lst = []
            for i, shp in enumerate(self._core.cache_shapes(which, self.batch, self.device)):
                t = torch.empty(shp, dtype=torch.float32, device=self.device)
                _lib.check(self._lib.hil_state_export_cache(self._h, which, i, t.data_ptr(), _stream_ptr(self.device)))
                lst.append(t)
            out.append(lst)
        return out[0], out[1]

    def load(self, cache_enc: tp.Sequence[Tensor], cache_dec: tp.Sequence[Tensor]) -> None:
        self._check()
        for which, lst in ((_lib.HIL_ENCODER, cache_enc), (_lib.HIL_DECODER, cache_dec)):
            shapes = self._core.cache_shapes(which, self.batch, self.device)
            if len(lst) != len(shapes):
                raise ValueError(f"expected {len(shapes)} cache tensors, got {len(lst)}")
            for i, (t, shp) in enumerate(zip(lst, shapes)):
                _require_cuda(t, "cache")
                if tuple(t.shape) != shp:
                    raise ValueError(f"cache shape {tuple(t.shape)} != {shp}")
                t = t.contiguous()
                _lib.check(self._lib.hil_state_import_cache(self._h, which, i, t.data_ptr(), _stream_ptr(self.device)))

    def __del__(self):  # pragma: no cover
        try:
            if self._h:
                self._lib.hil_state_destroy(self._h)
                self._h = None
        except Exception:
            pass

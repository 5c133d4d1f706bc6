"""hilcodec_b200 -- B200-native (sm_100a) HILCodec encode -> RVQ -> decode.

Drop-in modules live in `hilcodec_b200.streaming` (mirrors the reference's
`models/hilcodec/streaming.py`); the compute is in `libhilcodec_b200.so`
(C ABI: include/hilcodec_b200.h), built by `hilcodec_b200/build.py`.
"""
from .weights import CONFIGS, HIL_MUSIC, HIL_SPEECH, CodecConfig  # noqa: F401

__all__ = ["CodecConfig", "HIL_SPEECH", "HIL_MUSIC", "CONFIGS", "streaming"]


def __getattr__(name):  # lazy: importing the package must not need torch/CUDA
    if name == "streaming":
        import importlib
        return importlib.import_module(".streaming", __name__)
    raise AttributeError(name)

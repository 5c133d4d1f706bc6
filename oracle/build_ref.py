"""Recipe that puts the REFERENCE ITSELF next to the oracle: copies the few files of the reference's deployment
path, unmodified, from where they lie under /root/reference into `oracle/_ref/` (git-ignored, NOT gpurun-ignored:
it travels to the GPU box like the built .so, and never enters the history).

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The copy is what `bench.py --impl reference` and `bench.py`'s
`cpu_baseline` leg time (`kind: "reference"`: the reference's own `streaming.py` classes with the published
weights on the box's host cores), and what the `-m gpu` parity tests may import as a second checker beside the
oracle port.  Nothing under `hilcodec_b200/` imports it.

    python oracle/build_ref.py            # run by __graft_entry__.build() whenever /root/reference exists

Files (all the deployment graph needs; `functional` / `utils` are imported by streaming.py at module scope for one
unused symbol each and are stubbed by oracle/ref_shim.py when they are absent):
  models/hilcodec/streaming.py       Encoder / Decoder / ResidualVQ / Dequantizer / HILCodec   (the contract)
  models/hilcodec/causal_layers.py   CausalConv1d / CausalConvTranspose1d / CausalSTFT
  configs/hilcodec_{speech,music}.yaml   model_kwargs
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = os.environ.get("HILCODEC_REFERENCE", "/root/reference")
FILES = [
    os.path.join("models", "hilcodec", "streaming.py"),
    os.path.join("models", "hilcodec", "causal_layers.py"),
    os.path.join("configs", "hilcodec_speech.yaml"),
    os.path.join("configs", "hilcodec_music.yaml"),
]


def build_ref(verbose: bool = False) -> bool:
    """Returns True when oracle/_ref holds a complete copy afterwards."""
    if os.path.isfile(os.path.join(SRC, FILES[0])):
        for rel in FILES:
            src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
                shutil.copyfile(src, dst)
                if verbose:
                    print("copied", rel)
    return all(os.path.isfile(os.path.join(DST, rel)) for rel in FILES)


if __name__ == "__main__":
    ok = build_ref(verbose=True)
    print("oracle/_ref", "complete" if ok else "INCOMPLETE (no reference tree here and no earlier copy)")
    sys.exit(0 if ok else 1)

"""Host statement of the index bitstream format (no GPU)."""
import numpy as np

from hilcodec_b200 import bitstream
from oracle import bitstream_oracle


def test_numpy_pack_roundtrip_and_rates():
    rng = np.random.default_rng(0)
    for n in (1, 2, 8, 12):
        idx = rng.integers(0, 1024, size=(n, 2, 9))
        packed = bitstream_oracle.pack_numpy(idx)
        assert packed.shape == (2, 9, bitstream.bytes_per_frame(n))
        assert np.array_equal(bitstream_oracle.unpack_numpy(packed, n), idx)
    # known answer: two 10-bit values 0x3FF, 0x001 -> bytes FF 07 00 (LSB first)
    kat = np.array([[[0x3FF]], [[0x001]]])
    assert bitstream_oracle.pack_numpy(kat)[0, 0].tolist() == [0xFF, 0x07, 0x00]
    assert bitstream_oracle.bytes_per_frame(12) == bitstream.bytes_per_frame(12) == 15
    # 0.75 kbps per codebook at 75 frames/s
    assert bitstream.bytes_per_frame(8) * 8 * 75 == 6000
    assert bitstream.bytes_per_frame(12) * 8 * 75 == 9000

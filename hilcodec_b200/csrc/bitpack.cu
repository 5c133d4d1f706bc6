// Bitstream packing of the RVQ indices: the step after the quantizer on a real wire.
// The reference stores indices as int16 .npy (test_onnx.py:99); HILCodec's nominal rate is
// log2(1024) = 10 bits per codebook per frame (0.75 kbps per codebook at 75 frames/s), so a frame with n
// codebooks packs into ceil(10 n / 8) bytes: 10 bytes at n = 8 (6 kbps), 15 bytes at n = 12 (9 kbps).
// Layout: frame-major [B*F][bytes_per_frame]; inside a frame the n indices are concatenated LSB first.
#include "common.cuh"

namespace hil {

__global__ void pack_indices_kernel(const int64_t* __restrict__ idx, long long frames, int n, int bits,
                                    int bytes_per_frame, uint8_t* __restrict__ out) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    uint8_t* o = out + f * bytes_per_frame;
    unsigned long long acc = 0;
    int have = 0, pos = 0;
    for (int s = 0; s < n; ++s) {
        const unsigned long long v = (unsigned long long)idx[(size_t)s * frames + f] & ((1ull << bits) - 1);
        acc |= v << have;
        have += bits;
        while (have >= 8) {
            o[pos++] = (uint8_t)(acc & 0xff);
            acc >>= 8;
            have -= 8;
        }
    }
    if (have > 0) o[pos++] = (uint8_t)(acc & 0xff);
}

__global__ void unpack_indices_kernel(const uint8_t* __restrict__ in, long long frames, int n, int bits,
                                      int bytes_per_frame, int64_t* __restrict__ idx) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= frames) return;
    const uint8_t* src = in + f * bytes_per_frame;
    unsigned long long acc = 0;
    int have = 0, pos = 0;
    for (int s = 0; s < n; ++s) {
        while (have < bits) {
            acc |= (unsigned long long)src[pos++] << have;
            have += 8;
        }
        idx[(size_t)s * frames + f] = (int64_t)(acc & ((1ull << bits) - 1));
        acc >>= bits;
        have -= bits;
    }
}

cudaError_t launch_pack_indices(const int64_t* idx, long long frames, int n, int bits, uint8_t* out, cudaStream_t st) {
    if (frames == 0 || n == 0) return cudaSuccess;
    if (bits < 1 || bits > 16) return cudaErrorInvalidValue;
    const int bpf = (n * bits + 7) / 8;
    pack_indices_kernel<<<(unsigned)((frames + 127) / 128), 128, 0, st>>>(idx, frames, n, bits, bpf, out);
    return cudaGetLastError();
}

cudaError_t launch_unpack_indices(const uint8_t* in, long long frames, int n, int bits, int64_t* idx, cudaStream_t st) {
    if (frames == 0 || n == 0) return cudaSuccess;
    if (bits < 1 || bits > 16) return cudaErrorInvalidValue;
    const int bpf = (n * bits + 7) / 8;
    unpack_indices_kernel<<<(unsigned)((frames + 127) / 128), 128, 0, st>>>(in, frames, n, bits, bpf, idx);
    return cudaGetLastError();
}

}  // namespace hil

// convert.cu -- decoded frames to 8-bit luma on the device (ingest row f-3).
//
// The dataset players of the reference decode a PNG and call `DynamicImage::to_luma8()` before the frame
// reaches the tracker (src/datasets/euroc_player.rs:227-231, tum_vi_player.rs:227-231,
// fourseasons_player.rs:226-230): an identity copy for 8-bit grey (EuRoC), a 16 -> 8 bit reduction for the
// TUM-VI PNGs, a luma transform for colour input.  Here the raw decoded buffer is uploaded and converted by a
// kernel that writes straight into the tracker's input slot, so the CPU pass and one host copy disappear.
// The arithmetic is the `image` crate's (0.25, not vendored: PARITY UNPINNED), restated:
//   Luma<u16> -> Luma<u8>:  (c + 128) / 257                      (FromPrimitive<u16> for u8, rounds c / 257)
//   Rgb<u8>   -> Luma<u8>:  (2126 r + 7152 g + 722 b) / 10000    (rgb_to_luma, integer division)
#include "common.cuh"

namespace rsvio {

namespace {

__global__ void __launch_bounds__(256)
to_luma8_kernel(const uint8_t *__restrict__ src, size_t stride_bytes, int w, int h, int fmt, uint8_t *__restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t *row = src + (size_t)y * stride_bytes;
    uint32_t v;
    if (fmt == RSVIO_PIXEL_LUMA16) {
        const uint32_t c = (uint32_t)row[2 * x] | ((uint32_t)row[2 * x + 1] << 8);  // native (little-endian) u16
        v = (c + 128u) / 257u;
    } else if (fmt == RSVIO_PIXEL_RGB8) {
        v = (2126u * row[3 * x] + 7152u * row[3 * x + 1] + 722u * row[3 * x + 2]) / 10000u;
    } else {
        v = row[x];
    }
    dst[(size_t)y * w + x] = (uint8_t)v;
}

}  // namespace

int pixel_bytes(int fmt) {
    return fmt == RSVIO_PIXEL_LUMA16 ? 2 : (fmt == RSVIO_PIXEL_RGB8 ? 3 : (fmt == RSVIO_PIXEL_LUMA8 ? 1 : 0));
}

cudaError_t convert_to_luma8(const uint8_t *d_src, size_t stride_bytes, int w, int h, int fmt, uint8_t *d_dst,
                             cudaStream_t st) {
    if (w <= 0 || h <= 0) return cudaSuccess;
    to_luma8_kernel<<<dim3((unsigned)((w + 255) / 256), (unsigned)h, 1), 256, 0, st>>>(d_src, stride_bytes, w, h, fmt, d_dst);
    return cudaGetLastError();
}

}  // namespace rsvio

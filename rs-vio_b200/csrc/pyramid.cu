// pyramid.cu -- u8 image pyramid, every level resampled from the base image, bit-exact.
//
// Replaces build_image_pyramid (reference src/feature_tracker/feature_tracker.rs:209-220), i.e.
// image 0.25 `imageops::resize(img, w0 >> i, h0 >> i, FilterType::Triangle)` per level: vertical
// pass first into an un-rounded f32 intermediate, then the horizontal pass, clamp, round half away
// from zero.  Weights are tent(|i - c| / ratio) normalised by their f32 sum and accumulated as
// `t += pix * w` in tap order with separately rounded multiply and add.
//
// Four kernels share the work (DESIGN.md "pyramid"):
//
//  pyr_fast_kernel    levels 0..3 of all images in ONE pass over the base image (HBM-bound part).
//      When w0 and h0 are divisible by 2^i the level-i filter is the dyadic tent
//      (1,3,..,2r-1,2r-1,..,3,1)/(2r^2), r = 2^i, and for interior outputs every f32 operation
//      of the reference is exact up to level 3 (<= 22 significant bits), so the result equals
//      round_half_up(N / 4r^4) with N the integer 2-D tent sum.  The tents obey
//          T_2r[n] = T_r[n] + 2 T_r[n - r] + T_r[n - 2r],   T_1 = (1, 1)
//      so with G_r[m] = (T_r * p)(m r) ("even phase") and O_r[o] = (T_r * p)(o r - r/2) (the
//      level output, "odd phase"):
//          O_2r[o] = G_r[2o-1] + 2 G_r[2o] + G_r[2o+1],   G_2r[m] = G_r[2m] + 2 G_r[2m+1] + G_r[2m+2]
//      i.e. one 3x3 binomial per level on a 4x smaller array.  Level 1 (all the full-resolution
//      work) is 4 DP4A per output: the (1,3,3,1)x(1,3,3,1) weights fit signed bytes and the 4
//      row dot-products chain through the DP4A accumulator.  A CTA stages a 256x64 tile (+halo)
//      in shared memory with 16-byte loads, writes level 0 from it with 16-byte stores and keeps
//      G_2 / G_4 in shared memory; only u8 levels (and, for deeper pyramids, the 1/64-size G_8)
//      go to HBM: traffic == algorithmic bytes.
//  pyr_deep_kernel    levels 4 and 5 from G_8 (exact integers).  There the reference's horizontal
//      f32 accumulation is no longer exact (26 > 24 bits), but its error is < 2^-10 grey levels,
//      so round(N / D) is the reference's byte unless frac(N / D) is within 2^-10 of 1/2; those
//      rare outputs go to a queue that extra CTAs of the following pyr_generic_kernel launch drain as
//      1x1 tiles with the faithful f32 chains.
//  pyr_generic_kernel faithful f32 chains, one CTA per tile descriptor: the outermost ring of every
//      dyadic level (truncated windows: weights are renormalised by a non-power-of-two sum,
//      nothing is exact) as 1-row / 1-column tiles, whole levels whose ratio is not a power of
//      two (w0 or h0 not divisible by 2^i), levels >= 6, and everything when the fast path's
//      alignment requirements (16-byte aligned rows, w0 % 16 == 0) do not hold.
//
// Every multiply/add that can round is an explicit __fmul_rn/__fadd_rn and the library is built
// with -fmad=false.
#include "common.cuh"

#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef PYR_VEXACT
#define PYR_VEXACT 1  // generic kernel: vertical pass of complete dyadic windows (levels 4, 5) as exact integer tents
#endif

namespace rsvio {

namespace {

// ------------------------------------------------------------------------------------------
// tap tables (faithful weights) -- image 0.25 sample.rs weight construction
// ------------------------------------------------------------------------------------------
__global__ void taps_kernel(int n_in, int n_out, int kmax, int *left, int *cnt, float *w) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= n_out) return;
    const float ratio = __fdiv_rn((float)n_in, (float)n_out);
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float support = __fmul_rn(1.0f, sratio);
    float c = __fmul_rn(__fadd_rn((float)o, 0.5f), ratio);
    long long l = (long long)floorf(__fsub_rn(c, support));
    l = l < 0 ? 0 : (l > (long long)n_in - 1 ? (long long)n_in - 1 : l);
    long long r = (long long)ceilf(__fadd_rn(c, support));
    r = r < l + 1 ? l + 1 : (r > (long long)n_in ? (long long)n_in : r);
    c = __fsub_rn(c, 0.5f);
    const int n = (int)(r - l);
    left[o] = (int)l;
    cnt[o] = n;
    float *wo = w + (size_t)o * kmax;
    float sum = 0.0f;
    for (int k = 0; k < n; ++k) {
        const float x = fabsf(__fdiv_rn(__fsub_rn((float)(l + k), c), sratio));
        const float t = x < 1.0f ? __fsub_rn(1.0f, x) : 0.0f;
        wo[k] = t;
        sum = __fadd_rn(sum, t);
    }
    for (int k = 0; k < n; ++k) wo[k] = __fdiv_rn(wo[k], sum);
    for (int k = n; k < kmax; ++k) wo[k] = 0.0f;
}

// u8 -> f32 (exact).  Behind an opaque 32-bit value the conversion is I2FP.F32.U32 on the FMA pipe; left to itself the
// compiler narrows it to I2F.U16 on the quarter-rate XU pipe (measured on the LK kernel: 3.4 % of its time).
__device__ __forceinline__ float u8f(uint8_t b) {
    uint32_t v = b;
    asm volatile("mov.b32 %0, %0;" : "+r"(v));
    return __uint2float_rn(v);
}

__device__ __forceinline__ const uint8_t *image_ptr(const uint8_t *src0, const uint8_t *src1, int split, int j,
                                                    size_t pitch) {
    return j < split ? src0 + (size_t)j * pitch : src1 + (size_t)(j - split) * pitch;
}

// ------------------------------------------------------------------------------------------
// generic faithful tile kernel: one CTA per tile descriptor (whole non-dyadic levels, and the
// border ring of the dyadic ones as 1-row / 1-column tiles)
// ------------------------------------------------------------------------------------------
constexpr int kGenThreads = 256;
constexpr int kGenSmemFloats = 5376;  // f32 intermediate (ncols x noy) per CTA

struct TileDesc {
    int16_t lev, nox, noy, pad;  // lev == 0: copy rows [oy0, oy0 + noy)
    int32_t ox0, oy0;
};

// a near-tie output of the deep dyadic kernel, queued for the faithful chains: x = image, y = lev << 28 | oy << 14 | ox
typedef uint2 TieEntry;

// Vertical pass of a COMPLETE window on an axis whose ratio is the power of two RY: the reference's weights are
// (2k + 1) / (2 RY^2), mirrored -- exact in f32 -- and so is every product and partial sum of its chain (< 2^24 in units
// of 1 / (2 RY^2) for RY <= 128), so the chain's value is the integer tent sum scaled by a power of two.  Four columns
// per 32-bit load; the byte of column c is selected by putting the weight into byte c of the DP4A operand: one
// instruction per column and tap instead of the chain's PRMT + FADD + FMUL + FADD.
template <int RY>
__device__ __forceinline__ float4 vtent_exact4(const uint8_t *__restrict__ p, size_t stride) {
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int k = 0; k < 2 * RY; ++k) {
        const uint32_t q = __ldg(reinterpret_cast<const uint32_t *>(p + (size_t)k * stride));
        const uint32_t wk = 2u * (uint32_t)(k < RY ? k : 2 * RY - 1 - k) + 1u;
        a0 = __dp4a(q, wk, a0);
        a1 = __dp4a(q, wk << 8, a1);
        a2 = __dp4a(q, wk << 16, a2);
        a3 = __dp4a(q, wk << 24, a3);
    }
    constexpr float sc = 1.0f / (2.0f * RY * RY);
    return make_float4(__fmul_rn((float)a0, sc), __fmul_rn((float)a1, sc), __fmul_rn((float)a2, sc), __fmul_rn((float)a3, sc));
}

__device__ __forceinline__ void generic_tile(const PyrDesc &pd, const PyrTaps &taps, const TileDesc td, int img,
                                             const uint8_t *__restrict__ src0, const uint8_t *__restrict__ src1, int split,
                                             size_t stride, size_t pitch, uint8_t *__restrict__ dst, float *__restrict__ v) {
    const uint8_t *__restrict__ src = image_ptr(src0, src1, split, img, pitch);
    uint8_t *__restrict__ out = dst + (size_t)img * pd.pyr_stride;
    const int lev = td.lev;
    const int w0 = pd.lv[0].w;

    if (lev == 0) {  // same size: plain copy (image `resize` short-circuit)
        const int y0 = td.oy0, ny = td.noy;
        const bool vec = ((w0 & 15) == 0) && ((stride & 15) == 0) && ((((size_t)src) & 15) == 0);
        if (vec) {
            const int wv = w0 >> 4;
            for (int idx = threadIdx.x; idx < wv * ny; idx += kGenThreads) {
                const int y = idx / wv, x = idx - y * wv;
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)(y0 + y) * stride) + x);
                reinterpret_cast<uint4 *>(out + (size_t)(y0 + y) * w0)[x] = q;
            }
        } else {
            for (int idx = threadIdx.x; idx < w0 * ny; idx += kGenThreads) {
                const int y = idx / w0, x = idx - y * w0;
                out[(size_t)(y0 + y) * w0 + x] = __ldg(src + (size_t)(y0 + y) * stride + x);
            }
        }
        return;
    }

    const int wl = pd.lv[lev].w;
    const TapTable tx = taps.x[lev], ty = taps.y[lev];
    const int ox0 = td.ox0, oy0 = td.oy0, nox = td.nox, noy = td.noy;
    const int xl = tx.left[ox0];
    const int xr = tx.left[ox0 + nox - 1] + tx.cnt[ox0 + nox - 1];
    const int xa = xl & ~3;                    // 4-aligned first column
    const int ngrp = (xr - xa + 3) >> 2;       // groups of 4 columns, one 32-bit load per tap
    const int ncols = ngrp * 4;                // intermediate columns kept (>= xr - xl)
    const bool vec4 = ((((size_t)src) | stride) & 3) == 0;
    // vertical ratio 2^lev exactly (h0 divisible): complete windows (2 ry taps) take the exact integer path
    const int ry = (pd.lv[0].h & ((1 << lev) - 1)) == 0 ? (1 << lev) : 0;

    // vertical pass: f32, tap order, multiply and add rounded separately; 4 columns per thread
    for (int idx = threadIdx.x; idx < ngrp * noy; idx += kGenThreads) {
        const int cy = idx / ngrp, g = idx - cy * ngrp;
        const int oy = oy0 + cy;
        const int top = ty.left[oy], n = ty.cnt[oy];
        const float *__restrict__ wy = ty.w + (size_t)oy * ty.kmax;
        const int x = xa + 4 * g;
        RSVIO_CHECK(top >= 0 && top + n <= pd.lv[0].h && x >= 0 && x < w0 && cy * ncols + 4 * g + 3 < kGenSmemFloats);
        const uint8_t *__restrict__ p = src + (size_t)top * stride + x;
        float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f, t3 = 0.0f;
        if (PYR_VEXACT && vec4 && x + 4 <= w0 && n == 2 * ry && (ry == 16 || ry == 32)) {
            const float4 t = ry == 32 ? vtent_exact4<32>(p, stride) : vtent_exact4<16>(p, stride);
            t0 = t.x, t1 = t.y, t2 = t.z, t3 = t.w;
        } else if (vec4 && x + 4 <= w0) {
            // u8 -> f32 without the quarter-rate I2F: one PRMT drops byte j into the mantissa of 2^23
            // (0x4B0000bb == 8388608 + b), one exact FADD removes the 2^23.  This loop is conversion-bound otherwise.
            constexpr float kTwo23 = 8388608.0f;
#pragma unroll 16  // the loads do not depend on the f32 chains: 16 in flight per thread
            for (int k = 0; k < n; ++k) {
                const uint32_t q = __ldg(reinterpret_cast<const uint32_t *>(p + (size_t)k * stride));
                const float wk = __ldg(wy + k);
                const float p0 = __fadd_rn(__uint_as_float(__byte_perm(q, 0x4B000000u, 0x7440)), -kTwo23);
                const float p1 = __fadd_rn(__uint_as_float(__byte_perm(q, 0x4B000000u, 0x7441)), -kTwo23);
                const float p2 = __fadd_rn(__uint_as_float(__byte_perm(q, 0x4B000000u, 0x7442)), -kTwo23);
                const float p3 = __fadd_rn(__uint_as_float(__byte_perm(q, 0x4B000000u, 0x7443)), -kTwo23);
                t0 = __fadd_rn(t0, __fmul_rn(p0, wk));
                t1 = __fadd_rn(t1, __fmul_rn(p1, wk));
                t2 = __fadd_rn(t2, __fmul_rn(p2, wk));
                t3 = __fadd_rn(t3, __fmul_rn(p3, wk));
            }
        } else {
            for (int k = 0; k < n; ++k) {
                const uint8_t *__restrict__ r = p + (size_t)k * stride;
                const float wk = __ldg(wy + k);
                if (x + 0 < w0) t0 = __fadd_rn(t0, __fmul_rn(u8f(__ldg(r + 0)), wk));
                if (x + 1 < w0) t1 = __fadd_rn(t1, __fmul_rn(u8f(__ldg(r + 1)), wk));
                if (x + 2 < w0) t2 = __fadd_rn(t2, __fmul_rn(u8f(__ldg(r + 2)), wk));
                if (x + 3 < w0) t3 = __fadd_rn(t3, __fmul_rn(u8f(__ldg(r + 3)), wk));
            }
        }
        *reinterpret_cast<float4 *>(&v[cy * ncols + 4 * g]) = make_float4(t0, t1, t2, t3);
    }
    __syncthreads();
    // horizontal pass, clamp, round half away from zero
    uint8_t *__restrict__ lo = out + pd.lv[lev].off;
    for (int idx = threadIdx.x; idx < nox * noy; idx += kGenThreads) {
        const int cy = idx / nox, cx = idx - cy * nox;
        const int ox = ox0 + cx;
        const int left = tx.left[ox] - xa, n = tx.cnt[ox];
        const float *__restrict__ wx = tx.w + (size_t)ox * tx.kmax;
        const float *__restrict__ vr = &v[cy * ncols + left];
        RSVIO_CHECK(left >= 0 && left + n <= ncols && ox < wl && oy0 + cy < pd.lv[lev].h);
        float t = 0.0f;
#pragma unroll 16  // weights and column values of 16 taps in flight; only the adds are serial
        for (int k = 0; k < n; ++k) t = __fadd_rn(t, __fmul_rn(vr[k], __ldg(wx + k)));
        t = fminf(fmaxf(t, 0.0f), 255.0f);
        lo[(size_t)(oy0 + cy) * wl + ox] = (uint8_t)__float2uint_rz(roundf(t));
    }
}

// One CTA per static tile descriptor (blockIdx.x < n_static, image blockIdx.y); the remaining CTAs of the grid drain
// the tie queue of pyr_deep_kernel (any image).
__global__ void __launch_bounds__(kGenThreads, 5)
pyr_generic_kernel(const __grid_constant__ PyrDesc pd, const __grid_constant__ PyrTaps taps,
                   const TileDesc *__restrict__ tiles, int n_static, const TieEntry *__restrict__ tie_queue,
                   const unsigned *__restrict__ tie_count, const uint8_t *__restrict__ src0,
                   const uint8_t *__restrict__ src1, int split, size_t stride, size_t pitch,
                   uint8_t *__restrict__ dst) {
    __shared__ __align__(16) float v[kGenSmemFloats];
    if ((int)blockIdx.x < n_static) {
        generic_tile(pd, taps, tiles[blockIdx.x], blockIdx.y, src0, src1, split, stride, pitch, dst, v);
        return;
    }
    // Tie workers: one WARP per queued output (an interior output of dyadic level 4 or 5: complete 2r x 2r window).
    // A lane runs the vertical f32 chain of window column `lane` (and lane + 32 at level 5) with all its loads in
    // flight at once, then every lane walks the horizontal chain in tap order over shuffled column values.
    const int lane = threadIdx.x & 31, wpb = kGenThreads / 32;
    const unsigned workers = (gridDim.x - n_static) * gridDim.y * wpb;
    const unsigned cnt = *tie_count;
    for (unsigned i = (blockIdx.y * (gridDim.x - n_static) + (blockIdx.x - n_static)) * wpb + (threadIdx.x >> 5); i < cnt;
         i += workers) {
        const TieEntry e = tie_queue[i];
        const int lev = (int)(e.y >> 28), ox = (int)(e.y & 0x3fffu), oy = (int)((e.y >> 14) & 0x3fffu), img = (int)e.x;
        const uint8_t *__restrict__ src = image_ptr(src0, src1, split, img, pitch);
        const TapTable tx = taps.x[lev], ty = taps.y[lev];
        const int left = __ldg(tx.left + ox), nx = __ldg(tx.cnt + ox), top = __ldg(ty.left + oy), ny = __ldg(ty.cnt + oy);
        const float *__restrict__ wy = ty.w + (size_t)oy * ty.kmax;
        const float *__restrict__ wx = tx.w + (size_t)ox * tx.kmax;
        RSVIO_CHECK((lev == 4 || lev == 5) && nx <= 64 && ny <= 64 && left >= 0 && left + nx <= pd.lv[0].w && top >= 0 &&
                    top + ny <= pd.lv[0].h && ox < pd.lv[lev].w && oy < pd.lv[lev].h);
        float vc[2], wc[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int kx = lane + 32 * h;
            float t = 0.0f;
            if (kx < nx) {
                const uint8_t *__restrict__ p = src + (size_t)top * stride + (left + kx);
#pragma unroll 16
                for (int ky = 0; ky < ny; ++ky) t = __fadd_rn(t, __fmul_rn(u8f(__ldg(p + (size_t)ky * stride)), __ldg(wy + ky)));
            }
            vc[h] = t;
            wc[h] = kx < nx ? __ldg(wx + kx) : 0.0f;
        }
        float t = 0.0f;
        for (int k = 0; k < nx; ++k) {  // nx <= 64 (levels 4, 5)
            const float vk = __shfl_sync(0xffffffffu, k < 32 ? vc[0] : vc[1], k & 31);
            const float wk = __shfl_sync(0xffffffffu, k < 32 ? wc[0] : wc[1], k & 31);
            t = __fadd_rn(t, __fmul_rn(vk, wk));
        }
        if (lane == 0) {
            t = fminf(fmaxf(t, 0.0f), 255.0f);
            dst[(size_t)img * pd.pyr_stride + pd.lv[lev].off + (size_t)oy * pd.lv[lev].w + ox] = (uint8_t)__float2uint_rz(roundf(t));
        }
    }
}

// ------------------------------------------------------------------------------------------
// fast dyadic kernel: levels 0..NA (NA <= 3) from one shared-memory tile
// ------------------------------------------------------------------------------------------
// experiment switches (A/B builds: make VARIANT=x EXTRA="-DPYR_STAGE0=0 ...")
#ifndef PYR_STAGE0
#define PYR_STAGE0 1        // 1: unpredicated level-0 copy of full-height tiles, loads batched before the stores
#endif
#ifndef PYR_RING_NOINLINE
#define PYR_RING_NOINLINE 0 // 1: border ring as an out-of-line function
#endif
#ifndef PYR_OROUND
#define PYR_OROUND 1        // 1: rounding constant folded into the O accumulators
#endif
#ifndef PYR_SPANSTORE
#define PYR_SPANSTORE 1     // 1: level-1 store predicates as one unsigned compare against a per-lane row span
#endif
#ifndef PYR_TH
#define PYR_TH 48
#endif
// base-image tile owned by one CTA.  TH = 48: 480 and 1200 rows divide evenly (no half-empty bottom tiles) and the
// 3-level variant fits 8 CTAs per SM instead of 7; measured against 64 / 32 on 128 x 752x480: 36.7 / 39.5 / 37.4 us,
// no shape slower (gpurun_out/r2/pyr1/ab.log).
constexpr int TW = 256, TH = PYR_TH;
constexpr int PITCH = TW + 32;            // tile columns [X0 - 16, X0 + TW + 16), bytes per smem row
constexpr int NCOL = TW / 8;              // stage-1 column threads: 4 level-1 columns (8 px) each = one warp
constexpr int NSEG = 4;                   // stage-1 row segments = warps per CTA
constexpr int NTA = NCOL * NSEG;          // 128 threads per CTA
constexpr int G2P = TW / 2 + 8;           // G_2 columns kept in smem (u16): j in [X0/2 - 2, X0/2 + TW/2 + 6)
constexpr int kRingTaps = 16;             // dyadic levels <= 3: at most 2r = 16 taps

template <int NA>
struct FastCfg {
    static constexpr int et = NA == 3 ? 2 : (NA == 2 ? 1 : 0);  // extra G_2 rows above the O range
    static constexpr int eb = NA == 3 ? 3 : 0;                  // ... and below
    static constexpr int ytop = 2 * et + 1, ybot = 2 * eb + 2;  // tile rows [Y0 - ytop, Y0 + TH + ybot)
    static constexpr int ROWS = TH + ytop + ybot;
    static constexpr int NE = TH / 2 + et + eb;                 // level-1 rows emitted per tile (O and/or G)
    static constexpr int G2ROWS = NA >= 2 ? NE : 1;
    static constexpr int GX = NA == 3 ? 5 : (NA == 2 ? 2 : 0);  // G_2 columns right of the warp's 128
    static constexpr int G4C = TW / 4 + 2, G4R = TH / 4 + 2;    // G_4[X0/4 - 1 .. X0/4 + TW/4]
};

// faithful f32 weights of the truncated border windows of a dyadic level (ratio r = 2^lev): they
// depend on r only, so one set serves rows and columns.  Copied from the device tap tables.
struct RingWeights {
    float first[kRingTaps];   // output 0:       taps [0, 1.5 r)
    float inner[kRingTaps];   // interior output: taps [o r - r/2, o r + 1.5 r)
    float last[kRingTaps];    // output n-1:     taps [n_in - 1.5 r, n_in)
};

struct FastParams {
    PyrDesc pd;
    RingWeights ring[4];      // index = level (1..3)
    int split, n_images;  // images [0, split) come through tensor map 0, the rest through map 1
    uint8_t *dst;
    uint32_t *g8;         // [img][h0/8][w0/8] exported when deeper levels follow, else nullptr
    int tiles_x, tiles_y;
    unsigned magic_tpi, magic_tx;  // ceil(2^32 / d) for d = tiles per image, tiles_x (0: plain division)
    int nbuf;             // tile buffers per CTA: 2 = next tile prefetched during compute, 1 = more CTAs per SM
    unsigned *tile_counter;  // single-buffer mode: tiles are handed out dynamically from this counter ...
    unsigned *tile_counter_next;  // ... and the other one of the pair is cleared for the next launch
    unsigned *tie_count;     // deep levels follow: their tie queue is emptied here (stream order: before pyr_deep_kernel)
};

// n / d with the host-made multiplier m = ceil(2^32 / d); exact while n * d < 2^32 (the host checks)
__device__ __forceinline__ int magic_div(int n, int d, unsigned m) {
    return m ? (int)__umulhi((unsigned)n, m) : n / d;
}
inline unsigned magic_for(unsigned d, uint64_t n_max) {
    if (d < 2 || n_max * d >= (1ull << 32)) return 0u;
    return (unsigned)(((1ull << 32) + d - 1) / d);
}

__device__ __forceinline__ uint32_t dp4a_u(uint32_t a, uint32_t b, uint32_t c) {
    return __dp4a(a, b, c);  // u8 x u8 (weights <= 9), u32 accumulate
}

// window of output o of a dyadic level along one axis (n_out outputs, ratio r)
struct RingTap {
    int left, cnt;
    const float *w;
};
__device__ __forceinline__ RingTap ring_tap(const RingWeights &rw, int o, int n_out, int r) {
    RingTap t;
    if (o == 0) {
        t.left = 0;
        t.cnt = r + (r >> 1);
        t.w = rw.first;
    } else if (o == n_out - 1) {
        t.left = n_out * r - (r + (r >> 1));
        t.cnt = r + (r >> 1);
        t.w = rw.last;
    } else {
        t.left = o * r - (r >> 1);
        t.cnt = 2 * r;
        t.w = rw.inner;
    }
    return t;
}

// One base-image row of stage 1.  A thread holds 4 level-1 columns j0..j0+3 (O, odd phase) and
// the 4 even-phase columns j0-2..j0+1 (G); wr points at the tile word of pixels 2*j0..2*j0+3.
// FIRST: this row is row 2t-1 of the (1,3,3,1) vertical tent of O_t (tap weights 1|3 for the
// accumulators), else row 2t.  Flags select which accumulators are live (peeled first/last rows).
// GS = false: the G columns are j0-2 .. j0+1 (words wm1, w01.x, w01.y); GS = true: j0 .. j0+3 (words
// w01.x, w01.y, wp2), the streaming kernel's mapping.
// the O accumulators start from the rounding constant of `(N + 32) >> 6` (round half up of N / 64): the store is a shift
constexpr uint32_t kORound = PYR_OROUND ? 32u : 0u;
template <bool ROW_A, bool ON, bool OO, bool GN, bool GM, bool GO, bool GS = false>
__device__ __forceinline__ void stage1_words(uint32_t wm1, uint2 w01, uint32_t wp2, uint32_t (&aOn)[4],
                                             uint32_t (&aOo)[4], uint32_t (&aGn)[4], uint32_t (&aGm)[4],
                                             uint32_t (&aGo)[4]) {
    const uint32_t k1 = 0x01030301u, k3 = 0x03090903u;
    if (ON || OO) {
        const uint32_t v0 = __funnelshift_r(wm1, w01.x, 24), v1 = __funnelshift_r(w01.x, w01.y, 8);
        const uint32_t v2 = __funnelshift_r(w01.x, w01.y, 24), v3 = __funnelshift_r(w01.y, wp2, 8);
        if (ON) {  // O_t: row 2t-1 starts it with weight 1 (on top of the rounding constant), row 2t adds weight 3
            aOn[0] = dp4a_u(v0, ROW_A ? k1 : k3, ROW_A ? kORound : aOn[0]);
            aOn[1] = dp4a_u(v1, ROW_A ? k1 : k3, ROW_A ? kORound : aOn[1]);
            aOn[2] = dp4a_u(v2, ROW_A ? k1 : k3, ROW_A ? kORound : aOn[2]);
            aOn[3] = dp4a_u(v3, ROW_A ? k1 : k3, ROW_A ? kORound : aOn[3]);
        }
        if (OO) {  // O_{t-1}: rows 2t-1 (weight 3) and 2t (weight 1, completes it)
            aOo[0] = dp4a_u(v0, ROW_A ? k3 : k1, aOo[0]);
            aOo[1] = dp4a_u(v1, ROW_A ? k3 : k1, aOo[1]);
            aOo[2] = dp4a_u(v2, ROW_A ? k3 : k1, aOo[2]);
            aOo[3] = dp4a_u(v3, ROW_A ? k3 : k1, aOo[3]);
        }
    }
    if (GN || GM || GO) {
        const uint32_t ga = GS ? w01.x : wm1, gb = GS ? w01.y : w01.x, gc = GS ? wp2 : w01.y;
        const uint32_t g1 = __funnelshift_r(ga, gb, 16), g3 = __funnelshift_r(gb, gc, 16);
        if (GN && !ROW_A) {  // G_t starts on row 2t (weight 1)
            aGn[0] = dp4a_u(ga, k1, 0u);
            aGn[1] = dp4a_u(g1, k1, 0u);
            aGn[2] = dp4a_u(gb, k1, 0u);
            aGn[3] = dp4a_u(g3, k1, 0u);
        }
        if (GM) {  // G_{t-1}: rows 2t-1 and 2t carry weight 3
            aGm[0] = dp4a_u(ga, k3, aGm[0]);
            aGm[1] = dp4a_u(g1, k3, aGm[1]);
            aGm[2] = dp4a_u(gb, k3, aGm[2]);
            aGm[3] = dp4a_u(g3, k3, aGm[3]);
        }
        if (GO && ROW_A) {  // G_{t-2}: row 2t-1 is its last row (weight 1)
            aGo[0] = dp4a_u(ga, k1, aGo[0]);
            aGo[1] = dp4a_u(g1, k1, aGo[1]);
            aGo[2] = dp4a_u(gb, k1, aGo[2]);
            aGo[3] = dp4a_u(g3, k1, aGo[3]);
        }
    }
}

template <bool ROW_A, bool ON, bool OO, bool GN, bool GM, bool GO>
__device__ __forceinline__ void stage1_row(const uint32_t *__restrict__ wr, uint32_t (&aOn)[4], uint32_t (&aOo)[4],
                                           uint32_t (&aGn)[4], uint32_t (&aGm)[4], uint32_t (&aGo)[4]) {
    const uint32_t wm1 = wr[-1];
    const uint2 w01 = *reinterpret_cast<const uint2 *>(wr);
    const uint32_t wp2 = (ON || OO) ? wr[2] : 0u;
    stage1_words<ROW_A, ON, OO, GN, GM, GO>(wm1, w01, wp2, aOn, aOo, aGn, aGm, aGo);
}

// ---- TMA / mbarrier primitives (sm_100a) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(mbar), "r"(parity) : "memory");
}
// 3-D tiled bulk tensor load global -> shared (zero fill outside the tensor), completes on mbar
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int x, int y, int z, uint32_t mbar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(mbar) : "memory");
}

// The tile [X0 - 16, X0 + TW + 16) x [Y0 - ytop, Y0 + TH + ybot) is wider than the 256-element TMA
// box limit, so it is fetched as two overlapping boxes of RW columns: region A = tile columns
// [0, RW), region B = [RB, RB + RW).  Every 16-byte window a stage-1 lane needs lies in one of them.
constexpr int RW = 160, RB = 128;
constexpr int kSplitCol = 144;  // tile columns below this are read from region A, the rest from B

// ---- border ring helpers of the tile kernel -------------------------------------------------
// A tile in shared memory: tile column c of row ry is byte a[(c >= kSplitCol ? delta : 0) + ry * RW + c].
struct RingTile {
    const uint8_t *a;
    int delta, X0, Y0, ytop;
    __device__ __forceinline__ const uint8_t *at(int ry, int col) const {
        RSVIO_CHECK(ry >= 0 && col >= 0 && col < RB + RW);  // the caller's tap loops stay below the tile's row count
        return a + (col >= kSplitCol ? delta : 0) + ry * RW + col;
    }
};
__device__ __forceinline__ uint8_t ring_round(float t) {
    t = fminf(fmaxf(t, 0.0f), 255.0f);
    return (uint8_t)__float2uint_rz(roundf(t));
}

// vertical chains of border row `side` (0: oy = 0, 1: oy = hl - 1) of level LEV, one thread per tile column
template <int LEV>
__device__ __forceinline__ void ring_rows_vertical(const FastParams &P, const RingTile &rt, float *__restrict__ vrow,
                                                   int side, int tid) {
    constexpr int r = 1 << LEV, CNT = r + (r >> 1);
    const int hl = P.pd.lv[LEV].h, w0 = P.pd.lv[0].w;
    const int oy = side == 0 ? 0 : hl - 1;
    if (oy * r < rt.Y0 || oy * r >= rt.Y0 + TH) return;
    float wv[CNT];
#pragma unroll
    for (int k = 0; k < CNT; ++k) wv[k] = side == 0 ? P.ring[LEV].first[k] : P.ring[LEV].last[k];
    const int row0 = (side == 0 ? 0 : hl * r - CNT) - rt.Y0 + rt.ytop;
    const int clo = max(rt.X0 - (r >> 1), 0) - rt.X0 + 16, chi = min(rt.X0 + TW + (r >> 1), w0) - rt.X0 + 16;
    for (int c = clo + tid; c < chi; c += NTA) {
        const uint8_t *__restrict__ s = rt.at(row0, c);
        float v = 0.0f;
#pragma unroll
        for (int k = 0; k < CNT; ++k) v = __fadd_rn(v, __fmul_rn(u8f(s[k * RW]), wv[k]));
        vrow[c] = v;
    }
}

// horizontal chains of that row: one thread per output column of the tile
template <int LEV>
__device__ __forceinline__ void ring_rows_horizontal(const FastParams &P, const RingTile &rt,
                                                     const float *__restrict__ vrow, uint8_t *__restrict__ out, int side,
                                                     int t) {
    constexpr int r = 1 << LEV, CIN = 2 * r;
    const int wl = P.pd.lv[LEV].w, hl = P.pd.lv[LEV].h;
    const int oy = side == 0 ? 0 : hl - 1;
    if (oy * r < rt.Y0 || oy * r >= rt.Y0 + TH) return;
    for (; t < TW / r; t += NTA) {
        const int ox = rt.X0 / r + t;
        if (ox >= wl) return;
        float acc = 0.0f;
        if (ox > 0 && ox < wl - 1) {
            const float *__restrict__ v = vrow + (ox * r - (r >> 1) - rt.X0 + 16);
#pragma unroll
            for (int k = 0; k < CIN; ++k) acc = __fadd_rn(acc, __fmul_rn(v[k], P.ring[LEV].inner[k]));
        } else {
            const RingTap tx = ring_tap(P.ring[LEV], ox, wl, r);
            for (int k = 0; k < tx.cnt; ++k) acc = __fadd_rn(acc, __fmul_rn(vrow[tx.left + k - rt.X0 + 16], tx.w[k]));
        }
        out[P.pd.lv[LEV].off + (size_t)oy * wl + ox] = ring_round(acc);
    }
}

// border column `side` (0: ox = 0, 1: ox = wl - 1) of level LEV, interior rows only.  A group of GL adjacent lanes
// owns one output: lane kx runs the vertical chain of window column kx (2r taps), then the CX column values are
// combined in tap order through shuffles.  All NTA threads work (one pass per level and side), the lanes of a group
// read the same shared-memory words (broadcast), and no block barrier is needed.  (The first version gave each
// output to one thread: 48 serial taps at level 2, 192 at level 3, 8- to 15-way bank conflicts, and every border
// tile waited for that one warp: 23 % of the kernel's time.)
template <int LEV>
__device__ __forceinline__ void ring_cols(const FastParams &P, const RingTile &rt, uint8_t *__restrict__ out, int side,
                                          int tid) {
    constexpr int r = 1 << LEV, CX = r + (r >> 1), CY = 2 * r, ROWS = TH / r;
    constexpr int GL = CX <= 4 ? 4 : (CX <= 8 ? 8 : 16);  // lanes per output
    static_assert((ROWS * GL) % 32 == 0 && 32 % GL == 0, "groups must not straddle warps");
    const int wl = P.pd.lv[LEV].w, hl = P.pd.lv[LEV].h, w0 = P.pd.lv[0].w;
    const int ox = side == 0 ? 0 : wl - 1;
    if (ox * r < rt.X0 || ox * r >= rt.X0 + TW) return;  // uniform
    const int c0 = (side == 0 ? 0 : w0 - CX) - rt.X0 + 16;
    for (int item = tid; item < ROWS * GL; item += NTA) {  // whole warps enter and leave together
        const int ly = item / GL, kx = item % GL;
        const int oy = rt.Y0 / r + ly;
        const bool live = oy >= 1 && oy <= hl - 2;
        // The vertical window of an interior row is complete: its weights are the dyadic tent (2k+1) / (2 r^2), every
        // product and partial sum of the reference's f32 chain is exact (< 2^15 in units of 1 / (2 r^2)), so the chain
        // is the integer sum below, scaled once -- one IMAD per tap instead of convert, multiply, add.
        uint32_t nv = 0;
        if (live && kx < CX) {
            const uint8_t *__restrict__ s = rt.at(oy * r - (r >> 1) - rt.Y0 + rt.ytop, c0 + kx);
#pragma unroll
            for (int ky = 0; ky < CY; ++ky) nv += (uint32_t)(ky < r ? 2 * ky + 1 : 2 * (CY - 1 - ky) + 1) * s[ky * RW];
        }
        const float v = __fmul_rn((float)nv, 1.0f / (float)(2 * r * r));
        const int leader = (tid & 31) & ~(GL - 1);
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < CX; ++k) {
            const float vk = __shfl_sync(0xffffffffu, v, leader + k);
            t = __fadd_rn(t, __fmul_rn(vk, side == 0 ? P.ring[LEV].first[k] : P.ring[LEV].last[k]));
        }
        if (live && kx == 0) out[P.pd.lv[LEV].off + (size_t)oy * wl + ox] = ring_round(t);
    }
}

// The border ring of one tile: the reference's f32 chains (truncated windows are renormalised by a non-dyadic
// sum, so nothing is exact there), read from the staged tile.  sides: bit 0 top, 1 bottom, 2 left, 3 right.
#if PYR_RING_NOINLINE
#define PYR_RING_ATTR __noinline__
#else
#define PYR_RING_ATTR __forceinline__
#endif
template <int NA>
__device__ PYR_RING_ATTR void ring_of_tile(const FastParams &P, const uint8_t *tA, int delta, int X0, int Y0,
                                          float (*vbuf)[PITCH], uint8_t *__restrict__ out, int sides, int tid) {
    using C = FastCfg<NA>;
    const bool top = sides & 1, bottom = sides & 2, left = sides & 4, right = sides & 8;
    const RingTile rt{tA, delta, X0, Y0, C::ytop};
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {  // border rows: oy = 0 (top tiles), oy = hl - 1 (bottom tiles)
        if (side == 0 ? !top : !bottom) continue;
        // vbuf lives in the bytes of g2: the previous tile's readers passed the end-of-tile barrier; only a tile
        // that is both top and bottom reuses vbuf within this block
        if (side == 1 && top) __syncthreads();
        ring_rows_vertical<1>(P, rt, vbuf[0], side, tid);
        if (NA >= 2) ring_rows_vertical<2>(P, rt, vbuf[NA >= 2 ? 1 : 0], side, tid);
        if (NA >= 3) ring_rows_vertical<3>(P, rt, vbuf[NA >= 3 ? 2 : 0], side, tid);
        __syncthreads();
        ring_rows_horizontal<1>(P, rt, vbuf[0], out, side, tid);
        if (NA >= 2) ring_rows_horizontal<2>(P, rt, vbuf[NA >= 2 ? 1 : 0], out, side, (tid + 64) % NTA);
        if (NA >= 3) ring_rows_horizontal<3>(P, rt, vbuf[NA >= 3 ? 2 : 0], out, side, (tid + 32) % NTA);
    }
    // border columns (ox = 0 for left tiles, ox = wl - 1 for right tiles) without the corner rows
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
        if (side == 0 ? !left : !right) continue;
        ring_cols<1>(P, rt, out, side, tid);
        if (NA >= 2) ring_cols<2>(P, rt, out, side, tid);
        if (NA >= 3) ring_cols<3>(P, rt, out, side, tid);
    }
    if (top || bottom) __syncthreads();  // vbuf shares its bytes with g2, which stage 1 writes next
}

template <int NA>
struct FastSmem {
    using C = FastCfg<NA>;
    static constexpr int region = (C::ROWS * RW + 127) & ~127;
    static constexpr int off_tile = 0;                                   // [2 buffers][A, B][ROWS][RW]
    static constexpr int off_g2 = off_tile + 4 * region;
    static constexpr int off_g4 = (off_g2 + C::G2ROWS * G2P * 2 + 15) & ~15;
    static constexpr int off_vbuf = off_g2;  // ring scratch, dead before stage 1 writes g2
    static constexpr int end_g = (off_g4 + (NA >= 3 ? C::G4R * C::G4C * 4 : 4) + 15) & ~15;
    static constexpr int end_v = (off_vbuf + NA * PITCH * 4 + 15) & ~15;
    static constexpr int off_mbar = end_g > end_v ? end_g : end_v;
    static constexpr int bytes = off_mbar + 16 + 128;                    // + slack for the 128-B base alignment
};

#ifdef PYR_MINB
#define PYR_BOUNDS __launch_bounds__(NTA, PYR_MINB)
#else
#define PYR_BOUNDS __launch_bounds__(NTA)
#endif
template <int NA>
__global__ void PYR_BOUNDS
pyr_fast_kernel(const __grid_constant__ FastParams P, const __grid_constant__ CUtensorMap map0,
                const __grid_constant__ CUtensorMap map1) {
    using C = FastCfg<NA>;
    using S = FastSmem<NA>;
    extern __shared__ uint8_t smem_raw[];
    // 128-byte alignment for the TMA destinations; offset arithmetic on the array keeps the pointer in the
    // shared address space (LDS/STS with 32-bit addresses instead of generic loads)
    uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    // layout: [nbuf x (region A, region B)] [g2 | vbuf] [g4] [mbarriers]
    const int shrink = (2 - P.nbuf) * 2 * S::region;
    uint16_t(*g2)[G2P] = reinterpret_cast<uint16_t(*)[G2P]>(smem + S::off_g2 - shrink);
    uint32_t(*g4)[C::G4C] = reinterpret_cast<uint32_t(*)[C::G4C]>(smem + S::off_g4 - shrink);
    float(*vbuf)[PITCH] = reinterpret_cast<float(*)[PITCH]>(smem + S::off_vbuf - shrink);
    const uint32_t mbar0 = smem_u32(smem + S::off_mbar - shrink);

    const int tid = threadIdx.x;
    const int w0 = P.pd.lv[0].w, h0 = P.pd.lv[0].h;
    const int w1 = P.pd.lv[1].w, h1 = P.pd.lv[1].h;
    const int tiles_per_img = P.tiles_x * P.tiles_y;
    const int total = tiles_per_img * P.n_images;

    auto issue_load = [&](int tile_id, int buf) {  // one elected thread: both boxes of one tile
        const int img = magic_div(tile_id, tiles_per_img, P.magic_tpi), tl = tile_id - img * tiles_per_img;
        const int ty = magic_div(tl, P.tiles_x, P.magic_tx), tx = tl - ty * P.tiles_x;
        const CUtensorMap *map = img < P.split ? &map0 : &map1;
        const int z = img < P.split ? img : img - P.split;
        const uint32_t mbar = mbar0 + 8u * buf;
        const uint32_t dst = smem_u32(smem + S::off_tile + 2 * buf * S::region);
        mbar_expect_tx(mbar, 2u * C::ROWS * RW);
        tma_load_3d(dst, map, tx * TW - 16, ty * TH - C::ytop, z, mbar);
        tma_load_3d(dst + S::region, map, tx * TW - 16 + RB, ty * TH - C::ytop, z, mbar);
    };

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map0)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map1)) : "memory");
        mbar_init(mbar0, 1);
        mbar_init(mbar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (blockIdx.x == 0 && P.tie_count) *P.tie_count = 0u;
    }
    __syncthreads();
    const bool dbl = P.nbuf == 2;
    if (dbl && tid == 0 && (int)blockIdx.x < total) issue_load(blockIdx.x, 0);

    // single-buffer mode: tiles are handed out dynamically (border tiles carry extra ring work, so a
    // static assignment would let them decide the makespan).  Thread 0 draws the index of the next tile
    // at the top of a tile and issues its load as soon as the last reader of the buffer is done.
    __shared__ int next_tile;
    int my_next = 0;
    const bool boss = tid == NTA - 32;  // lane 0 of the warp with the shortest stage-1 segment
    auto hand_out = [&](int t) {
        next_tile = t;
        if (t < total) issue_load(t, 0);
    };
    // the first tile of a CTA is its block index (no round trip to the counter before the first load);
    // the counter hands out tiles gridDim.x, gridDim.x + 1, ...
    if (!dbl) {
        if (boss) {
            if (blockIdx.x == 0) *P.tile_counter_next = 0u;
            hand_out((int)blockIdx.x);
        }
        __syncthreads();
    }
    int it = 0;
    for (int tile_id = blockIdx.x;; ++it) {
    if (!dbl) tile_id = next_tile;
    else if (it > 0) tile_id += gridDim.x;
    if (tile_id >= total) break;
    const int buf = dbl ? (it & 1) : 0;
    if (!dbl) {
        if (boss) my_next = (int)(gridDim.x + atomicAdd(P.tile_counter, 1u));  // used after stage 0: latency hidden
    } else if (tid == 0 && tile_id + (int)gridDim.x < total) {
        issue_load(tile_id + gridDim.x, buf ^ 1);  // prefetch
    }
    // Plain image-major, row-major tile order.  Tried and dropped: longest-processing-time-first orders (border tiles of
    // all images first; or position-major over cost-sorted positions): no gain at 128 images, 35-50 % slower at 640
    // (neighbouring tiles share halo rows and DRAM pages; concurrent CTAs a whole image apart share nothing).
    const int img = magic_div(tile_id, tiles_per_img, P.magic_tpi), tl = tile_id - img * tiles_per_img;
    const int tile_y = magic_div(tl, P.tiles_x, P.magic_tx), tile_x = tl - tile_y * P.tiles_x;
    const int X0 = tile_x * TW, Y0 = tile_y * TH;
    uint8_t *__restrict__ out = P.dst + (size_t)img * P.pd.pyr_stride;
    const uint8_t *__restrict__ tA = smem + S::off_tile + 2 * buf * S::region;   // tile columns [0, RW)
    const uint8_t *__restrict__ tB = tA + S::region - RB;                        // indexable by tile column >= RB
    auto trow = [&](int ry, int col) -> const uint8_t * {
        RSVIO_CHECK(ry >= 0 && ry < C::ROWS && col >= 0 && col < RB + RW);
        return (col < kSplitCol ? tA : tB) + ry * RW + col;
    };
    mbar_wait(mbar0 + 8u * buf, (dbl ? (it >> 1) : it) & 1);

    // ---- stage 0: level 0 == input, 16-byte copies from the staged tile.  A thread owns one 16-byte
    // column and every (NTA/16)-th row: constant row offsets on the shared side, one global pointer advanced by
    // a constant stride; full-height tiles (uniform) run without row predicates.
    {
        constexpr int VPT = TW / 16, RPP = NTA / VPT;  // vectors per row, rows per pass
        const int nrow = min(TH, h0 - Y0), nvec = min(VPT, (w0 - X0) >> 4);
        const int vx = tid % VPT, y0 = tid / VPT;
        if (vx < nvec) {
            const uint8_t *__restrict__ s0 = trow(y0 + C::ytop, 16 + vx * 16);
            uint8_t *__restrict__ d0 = out + (size_t)(Y0 + y0) * w0 + X0 + vx * 16;
            const size_t dstep = (size_t)RPP * w0;
#if PYR_STAGE0 == 1
            if (nrow == TH) {
                uint4 q[TH / RPP];
#pragma unroll
                for (int k = 0; k < TH / RPP; ++k) q[k] = *reinterpret_cast<const uint4 *>(s0 + k * RPP * RW);
#pragma unroll
                for (int k = 0; k < TH / RPP; ++k) {
                    *reinterpret_cast<uint4 *>(d0) = q[k];
                    d0 += dstep;
                }
            } else
#endif
            {
#pragma unroll
                for (int k = 0; k < TH / RPP; ++k)
                    if (y0 + k * RPP < nrow)
                        *reinterpret_cast<uint4 *>(d0 + k * dstep) = *reinterpret_cast<const uint4 *>(s0 + k * RPP * RW);
            }
        }
    }

    // ---- border ring of levels 1..NA owned by this tile (before stage 1: its scratch row vbuf lives in the
    // bytes of g2).  Out of line: only border tiles pay for it, and none of its address arithmetic leaks into the
    // common path.
    const bool top = tile_y == 0, bottom = tile_y == P.tiles_y - 1;
    const bool left = tile_x == 0, right = tile_x == P.tiles_x - 1;
    if (top || bottom || left || right)
        ring_of_tile<NA>(P, tA, S::region - RB, X0, Y0, vbuf, out, (top ? 1 : 0) | (bottom ? 2 : 0) | (left ? 4 : 0) | (right ? 8 : 0), tid);

    // ---- stage 1: level-1 output O_2 (global) and G_2 (smem); 4 chained DP4A per value.
    // Warp = one row segment, lane c = level-1 columns j0 .. j0+3, j0 = X0/2 - 2 + 4c, both phases (the tile
    // owns level-1 columns [X0/2 - 2, X0/2 + 126)).  The lane's pixels 2 j0 - 4 .. 2 j0 + 11 are 16 bytes at
    // tile byte 8c + 8: two aligned 8-byte loads, conflict free.
    const int t_first = Y0 / 2 - C::et;    // first emitted level-1 row of the tile
    {
        const int c = tid & (NCOL - 1), seg = tid / NCOL;
        const int j0 = X0 / 2 - 2 + 4 * c;
        // balanced row segments: NE = 33 -> 9, 8, 8, 8 (the slowest warp decides when the CTA reaches the barrier)
        constexpr int SEG_BASE = C::NE / NSEG, SEG_EXTRA = C::NE % NSEG;
        const int ts = t_first + seg * SEG_BASE + min(seg, SEG_EXTRA);
        const int te = ts + SEG_BASE + (seg < SEG_EXTRA ? 1 : 0);
        const bool live = (2 * j0 - 8) < w0 && (2 * ts - 8) < h0 && ts < te;  // partial tiles: skip dead work
        // level-1 rows this thread stores: its segment, the tile's O range, the level's interior
        const int o_lo = max(max(ts, Y0 / 2), 1);
        const int o_hi = (j0 < w1 && j0 + 3 >= 1) ? min(min(te, Y0 / 2 + TH / 2), h1 - 1) : o_lo;
        // which of the lane's four columns are interior columns of level 1 does not depend on the row: j0 = 2 (mod 4),
        // so (j0, j0+1) and (j0+2, j0+3) are aligned byte pairs; only column 1 and column w1 - 2 end up alone in theirs
        const bool ov0 = j0 + 0 >= 1 && j0 + 0 <= w1 - 2, ov1 = j0 + 1 >= 1 && j0 + 1 <= w1 - 2;
        const bool ov2 = j0 + 2 >= 1 && j0 + 2 <= w1 - 2, ov3 = j0 + 3 >= 1 && j0 + 3 <= w1 - 2;
        const unsigned lone = (ov0 && !ov1 ? 1u : 0u) | (ov1 && !ov0 ? 2u : 0u) | (ov2 && !ov3 ? 4u : 0u) | (ov3 && !ov2 ? 8u : 0u);
        // Row range and column validity of a store folded into ONE unsigned compare each: row `to` is stored by this
        // lane's low pair iff to - o_lo < span_lo (0 rows when the pair is not interior).  Opaque registers: the
        // compiler otherwise re-derives the column predicates from j0 and w1 on every row (17 ISETP per level-1 row).
        unsigned span = (unsigned)max(o_hi - o_lo, 0);
        unsigned span_lo = (ov0 && ov1) ? span : 0u, span_hi = (ov2 && ov3) ? span : 0u, span_lone = lone ? span : 0u;
        asm volatile("mov.u32 %0, %0;" : "+r"(span_lo));
        asm volatile("mov.u32 %0, %0;" : "+r"(span_hi));
        asm volatile("mov.u32 %0, %0;" : "+r"(span_lone));
        uint32_t aOo[4] = {0, 0, 0, 0}, aOn[4] = {0, 0, 0, 0}, aGo[4] = {0, 0, 0, 0}, aGm[4] = {0, 0, 0, 0},
                 aGn[4] = {0, 0, 0, 0};
        uint8_t *__restrict__ orow = out + P.pd.lv[1].off + (ptrdiff_t)(ts - 1) * w1 + j0;  // row tau - 1
        // the lane's 16 bytes on base row 2*tau - 1 (lanes 0..15 read region A, 16..31 region B)
        const uint2 *__restrict__ wr = reinterpret_cast<const uint2 *>(
            (c < 16 ? tA : tB + RB) + (2 * ts - 1 - (Y0 - C::ytop)) * RW + 8 * (c & 15) + 8);
        uint2 *__restrict__ grow = reinterpret_cast<uint2 *>(&g2[0][4 * c]) + (ptrdiff_t)(ts - 2 - t_first) * (G2P / 4);
        constexpr bool G = NA >= 2;
        constexpr int RWW = RW / 8;  // row pitch in 8-byte units
#if PYR_SPANSTORE
        auto store_o = [&](int to) {  // O_{to} complete -> level 1, round half up: (N + 32) >> 6
            constexpr uint32_t kR = 32u - kORound;  // what of the rounding constant is not in the accumulator yet
            const unsigned dr = (unsigned)(to - o_lo);
            // two aligned 16-bit stores, predicated per lane (no divergent second path for the tiles' edge lanes)
            if (dr < span_lo)
                reinterpret_cast<uint16_t *>(orow)[0] = (uint16_t)__byte_perm((aOo[0] + kR) >> 6, (aOo[1] + kR) >> 6, 0x0040);
            if (dr < span_hi)
                reinterpret_cast<uint16_t *>(orow)[1] = (uint16_t)__byte_perm((aOo[2] + kR) >> 6, (aOo[3] + kR) >> 6, 0x0040);
            if (dr < span_lone) {  // one lane of the first and of the last tile column
                if (lone & 1u) orow[0] = (uint8_t)((aOo[0] + kR) >> 6);
                if (lone & 2u) orow[1] = (uint8_t)((aOo[1] + kR) >> 6);
                if (lone & 4u) orow[2] = (uint8_t)((aOo[2] + kR) >> 6);
                if (lone & 8u) orow[3] = (uint8_t)((aOo[3] + kR) >> 6);
            }
        };
#else  // round-2 baseline: predicates re-derived per row
        auto store_o = [&](int to) {  // O_{to} complete -> level 1, round half up: (N + 32) >> 6
            if (to >= o_lo && to < o_hi) {
                constexpr uint32_t kR = 32u - kORound;
                const uint32_t b0 = (aOo[0] + kR) >> 6, b1 = (aOo[1] + kR) >> 6, b2 = (aOo[2] + kR) >> 6, b3 = (aOo[3] + kR) >> 6;
                const bool pair_lo = ov0 && ov1, pair_hi = ov2 && ov3;
                if (pair_lo) reinterpret_cast<uint16_t *>(orow)[0] = (uint16_t)(b0 | (b1 << 8));
                if (pair_hi) reinterpret_cast<uint16_t *>(orow)[1] = (uint16_t)(b2 | (b3 << 8));
                if (lone) {
                    if (lone & 1u) orow[0] = (uint8_t)b0;
                    if (lone & 2u) orow[1] = (uint8_t)b1;
                    if (lone & 4u) orow[2] = (uint8_t)b2;
                    if (lone & 8u) orow[3] = (uint8_t)b3;
                }
            }
        };
#endif
        auto rotate = [&]() {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                aOo[q] = aOn[q];
                if (G) {
                    aGo[q] = aGm[q];
                    aGm[q] = aGn[q];
                }
            }
            wr += 2 * RWW;
            orow += w1;
            grow += G2P / 4;
        };
        if (live) {
            // rows read: 2 ts - 1 .. 2 te + 1 (+ 1 more without G); G rows written: ts .. te - 1
            RSVIO_CHECK(2 * ts - 1 - (Y0 - C::ytop) >= 0 && 2 * te + 1 - (Y0 - C::ytop) < C::ROWS);
            RSVIO_CHECK(!G || (ts - t_first >= 0 && te - 1 - t_first < C::G2ROWS && 4 * c + 3 < G2P));
            // O_t spans rows 2t-1..2t+2 (complete in iteration t+1, second row), G_t rows 2t..2t+3
            // (complete in iteration t+2, first row).  Iterations ts and ts+1 and te, te+1 are peeled:
            // they only feed the accumulators of rows this segment emits.
            { const uint2 q0 = (wr)[0], q1 = (wr)[1]; stage1_words<true, true, false, false, false, false, true>(q0.x, make_uint2(q0.y, q1.x), q1.y, aOn, aOo, aGn, aGm, aGo); }
            { const uint2 q0 = (wr + RWW)[0], q1 = (wr + RWW)[1]; stage1_words<false, true, false, G, false, false, true>(q0.x, make_uint2(q0.y, q1.x), q1.y, aOn, aOo, aGn, aGm, aGo); }
            rotate();
            if (ts + 1 < te) {
                { const uint2 q0 = (wr)[0], q1 = (wr)[1]; stage1_words<true, true, true, false, G, false, true>(q0.x, make_uint2(q0.y, q1.x), q1.y, aOn, aOo, aGn, aGm, aGo); }
                { const uint2 q0 = (wr + RWW)[0], q1 = (wr + RWW)[1]; stage1_words<false, true, true, G, G, false, true>(q0.x, make_uint2(q0.y, q1.x), q1.y, aOn, aOo, aGn, aGm, aGo); }
                store_o(ts);
                rotate();
                for (int tau = ts + 2; tau < te; ++tau) {
                    { const uint2 q0 = (wr)[0], q1 = (wr)[1]; stage1_words<true, true, true, false, G, G, true>(q0.x, make_uint2(q0.y, q1.x), q1.y, aOn, aOo, aGn, aGm, aGo); }
                    if (G) *grow = make_uint2(aGo[0] | (aGo[1] << 16), aGo[2] | (aGo[3] << 16));  // G_{tau-2}
                    { const uint2 q0 = (wr + RWW)[0], q1 = (wr + RWW)[1]; stage1_words<false, true, true, G, G, false, true>(q0.x, make_uint2(q0.y, q1.x), q1.y, aOn, aOo, aGn, aGm, aGo); }
                    store_o(tau - 1);
                    rotate();
                }
                // tau = te: finish O_{te-1}, G_{te-2}; keep G_{te-1} going; nothing new starts
                { const uint2 q0 = (wr)[0], q1 = (wr)[1]; stage1_words<true, false, true, false, G, G, true>(q0.x, make_uint2(q0.y, q1.x), q1.y, aOn, aOo, aGn, aGm, aGo); }
                if (G) *grow = make_uint2(aGo[0] | (aGo[1] << 16), aGo[2] | (aGo[3] << 16));
                { const uint2 q0 = (wr + RWW)[0], q1 = (wr + RWW)[1]; stage1_words<false, false, true, false, G, false, true>(q0.x, make_uint2(q0.y, q1.x), q1.y, aOn, aOo, aGn, aGm, aGo); }
                store_o(te - 1);
                rotate();
            } else {  // single-row segment: tau = te = ts + 1
                { const uint2 q0 = (wr)[0], q1 = (wr)[1]; stage1_words<true, false, true, false, G, false, true>(q0.x, make_uint2(q0.y, q1.x), q1.y, aOn, aOo, aGn, aGm, aGo); }
                { const uint2 q0 = (wr + RWW)[0], q1 = (wr + RWW)[1]; stage1_words<false, false, true, false, G, false, true>(q0.x, make_uint2(q0.y, q1.x), q1.y, aOn, aOo, aGn, aGm, aGo); }
                store_o(ts);
                rotate();
            }
            if (G) {  // tau = te+1, first row only: finish G_{te-1}
                { const uint2 q0 = (wr)[0], q1 = (wr)[1]; stage1_words<true, false, false, false, false, true, true>(q0.x, make_uint2(q0.y, q1.x), q1.y, aOn, aOo, aGn, aGm, aGo); }
                *grow = make_uint2(aGo[0] | (aGo[1] << 16), aGo[2] | (aGo[3] << 16));
            }
        }
        // last tile column: level-1 column X0/2 + 126 would be lane 0 of a further tile that does not exist
        // (the image ends there): 4 chained DP4A per output straight from the tile
        if (tile_x == P.tiles_x - 1) {
            const int j = X0 / 2 + 126;
            if (j >= 1 && j <= w1 - 2) {
                for (int idx = tid; idx < TH / 2; idx += NTA) {
                    const int t = Y0 / 2 + idx;
                    if (t < 1 || t > h1 - 2) continue;
                    const int ry = 2 * t - 1 - (Y0 - C::ytop);
                    uint32_t acc = 0;
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {  // pixels 2j-1 .. 2j+2 = tile bytes 267 .. 270 (region B)
                        const uint32_t *__restrict__ wq = reinterpret_cast<const uint32_t *>(tB + (ry + rr) * RW) + 66;
                        acc = dp4a_u(__funnelshift_r(wq[0], wq[1], 24), (rr == 0 || rr == 3) ? 0x01030301u : 0x03090903u, acc);
                    }
                    out[P.pd.lv[1].off + (size_t)t * w1 + j] = (uint8_t)((acc + 32u) >> 6);
                }
            }
        }
        // the few G_2 columns right of the warp's 128 (needed by G_4 / O_4 at the tile's right edge)
        if (G) {
            constexpr int GXD = C::GX > 0 ? C::GX : 1;
            for (int idx = tid; idx < C::GX * C::NE; idx += NTA) {
                const int gx = idx % GXD, tr = idx / GXD;             // column X0/2 + TW/2 - 2 + gx, row t_first + tr
                const int px = TW - 4 + 2 * gx + 16;                 // tile byte column of pixel 2j (region B)
                const int ry = 2 * (t_first + tr) - (Y0 - C::ytop);  // tile row of base row 2t
                uint32_t acc = 0;
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const uint32_t *__restrict__ wq = reinterpret_cast<const uint32_t *>(tB + (ry + rr) * RW) + (px >> 2);
                    const uint32_t v = (px & 2) ? __funnelshift_r(wq[0], wq[1], 16) : wq[0];
                    acc = dp4a_u(v, (rr == 0 || rr == 3) ? 0x01030301u : 0x03090903u, acc);
                }
                g2[tr][TW / 2 + gx] = (uint16_t)acc;
            }
        }
    }
    if (NA >= 2) {
        __syncthreads();
        // the tile buffer is dead from here on: fetch the next tile while stages 2 and 3 run
        if (!dbl && boss) hand_out(my_next);
        // ---- stage 2: level-2 output O_4 (global) and G_4 (smem) from G_2: 3x3 binomials as chained
        // 2-way dot products on the u16 pairs.  g2[t - t_first][j - (X0/2 - 2)].
        const int w2 = P.pd.lv[2].w, h2 = P.pd.lv[2].h;
        uint8_t *__restrict__ o2 = out + P.pd.lv[2].off;
        constexpr int OQ = TW / 16;                       // groups of 4 outputs per O_4 row
        constexpr int NO = OQ * (TH / 4);
        constexpr int GQ = NA >= 3 ? (C::G4C + 1) / 2 : 1;  // pairs of G_4 columns
        constexpr int NG = NA >= 3 ? GQ * C::G4R : 0;
        // byte weights for dp2a_lo: bytes (b0, b1) multiply (lo16, hi16) of the word
        const uint32_t kA[3] = {0x0100u, 0x0200u, 0x0100u};   // (0, 1) * row weight: hi half only
        const uint32_t kB[3] = {0x0102u, 0x0204u, 0x0102u};   // (2, 1) * row weight
        const uint32_t kC[3] = {0x0201u, 0x0402u, 0x0201u};   // (1, 2) * row weight
        const uint32_t kD[3] = {0x0001u, 0x0002u, 0x0001u};   // (1, 0) * row weight: lo half only
        for (int idx = tid; idx < NO + NG; idx += NTA) {
            if (idx < NO) {
                // 4 outputs ox = X0/4 + 4q + {0..3}, row oy = Y0/4 + ly.  Window of output k starts at g2
                // column 8q + 2k + 1 (odd): hi(A_i) + 2 lo(A_{i+1}) + hi(A_{i+1}), A_i = word 4q + k.
                const int q = idx % OQ, ly = idx / OQ;
                const int tc = 2 * ly + C::et;                 // centre row in g2
                RSVIO_CHECK(tc - 1 >= 0 && tc + 1 < C::G2ROWS && 2 * (4 * q + 4) + 1 < G2P);
                uint32_t n0 = 0, n1 = 0, n2 = 0, n3 = 0;
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
                    const uint32_t *__restrict__ gw = reinterpret_cast<const uint32_t *>(&g2[tc - 1 + rr][0]) + 4 * q;
                    const uint4 a = *reinterpret_cast<const uint4 *>(gw);
                    const uint32_t a4 = gw[4];
                    n0 = __dp2a_lo(a.y, kB[rr], __dp2a_lo(a.x, kA[rr], n0));
                    n1 = __dp2a_lo(a.z, kB[rr], __dp2a_lo(a.y, kA[rr], n1));
                    n2 = __dp2a_lo(a.w, kB[rr], __dp2a_lo(a.z, kA[rr], n2));
                    n3 = __dp2a_lo(a4, kB[rr], __dp2a_lo(a.w, kA[rr], n3));
                }
                const int ox = X0 / 4 + 4 * q, oy = Y0 / 4 + ly;
                if (oy >= 1 && oy <= h2 - 2 && ox < w2) {
                    const uint32_t b0 = (n0 + 512u) >> 10, b1 = (n1 + 512u) >> 10;
                    const uint32_t b2 = (n2 + 512u) >> 10, b3 = (n3 + 512u) >> 10;
                    uint8_t *__restrict__ dp = o2 + (size_t)oy * w2 + ox;
                    if (ox >= 1 && ox + 3 <= w2 - 2) {
                        *reinterpret_cast<uint32_t *>(dp) = __byte_perm(b0 | (b1 << 8), b2 | (b3 << 8), 0x5410);
                    } else {
                        if (ox + 0 >= 1 && ox + 0 <= w2 - 2) dp[0] = (uint8_t)b0;
                        if (ox + 1 >= 1 && ox + 1 <= w2 - 2) dp[1] = (uint8_t)b1;
                        if (ox + 2 >= 1 && ox + 2 <= w2 - 2) dp[2] = (uint8_t)b2;
                        if (ox + 3 >= 1 && ox + 3 <= w2 - 2) dp[3] = (uint8_t)b3;
                    }
                }
            } else if (NA >= 3) {
                // 2 values G_4[X0/4 - 1 + gm + {0,1}][Y0/4 - 1 + gs]: window of gm starts at g2 column 2 gm
                // (even): lo(A_i) + 2 hi(A_i) + lo(A_{i+1}), A_i = word gm
                const int k = idx - NO;
                const int gq = k % GQ, gs = k / GQ;
                const int gm = 2 * gq;
                const int tc = 2 * gs + 1;                     // centre row = 2 (gs - 1) + et + 1 with et == 2
                RSVIO_CHECK(tc - 1 >= 0 && tc + 1 < C::G2ROWS && 2 * (gm + 2) + 1 < G2P && gs < C::G4R);
                uint32_t n0 = 0, n1 = 0;
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
                    const uint32_t *__restrict__ gw = reinterpret_cast<const uint32_t *>(&g2[tc - 1 + rr][0]) + gm;
                    const uint32_t a0 = gw[0], a1 = gw[1], a2 = gw[2];
                    n0 = __dp2a_lo(a1, kD[rr], __dp2a_lo(a0, kC[rr], n0));
                    n1 = __dp2a_lo(a2, kD[rr], __dp2a_lo(a1, kC[rr], n1));
                }
                g4[gs][gm] = n0;
                if (gm + 1 < C::G4C) g4[gs][gm + 1] = n1;
            }
        }
    }
    if (NA >= 3) {
        __syncthreads();
        // ---- stage 3: level-3 output O_8 (global) and G_8 (global scratch, if deeper levels follow)
        const int w3 = P.pd.lv[3].w, h3 = P.pd.lv[3].h;
        uint8_t *__restrict__ o3 = out + P.pd.lv[3].off;
        constexpr int NO = (TW / 8) * (TH / 8);
        const int ng = P.g8 ? NO : 0;
        const int w8 = w0 / 8, h8 = h0 / 8;
        for (int idx = tid; idx < NO + ng; idx += NTA) {
            const bool is_o = idx < NO;
            const int k = is_o ? idx : idx - NO;
            const int lx = k % (TW / 8), ly = k / (TW / 8);
            // g4[s - (Y0/4 - 1)][m - (X0/4 - 1)]:  O window centre m = 2*ox, G window centre m = 2*ox + 1
            const int mc = 2 * lx + 1 + (is_o ? 0 : 1), sc = 2 * ly + 1 + (is_o ? 0 : 1);
            RSVIO_CHECK(sc + 1 < C::G4R && mc + 1 < C::G4C);
            const uint32_t *r0 = &g4[sc - 1][mc - 1], *r1 = &g4[sc][mc - 1], *r2 = &g4[sc + 1][mc - 1];
            const uint32_t n = r0[0] + 2u * r0[1] + r0[2] + 2u * (r1[0] + 2u * r1[1] + r1[2]) + r2[0] + 2u * r2[1] + r2[2];
            const int ox = X0 / 8 + lx, oy = Y0 / 8 + ly;
            if (is_o) {
                if (ox >= 1 && ox <= w3 - 2 && oy >= 1 && oy <= h3 - 2)
                    o3[(size_t)oy * w3 + ox] = (uint8_t)((n + 8192u) >> 14);
            } else if (ox < w8 && oy < h8) {
                P.g8[((size_t)img * h8 + oy) * w8 + ox] = n;
            }
        }
    }

    __syncthreads();  // every read of this buffer, g2, g4 and vbuf is done before the next tile reuses them
    if (NA < 2 && !dbl) {
        if (boss) hand_out(my_next);
        __syncthreads();
    }
    }  // persistent tile loop
}

// ------------------------------------------------------------------------------------------
// deep dyadic levels (4, 5) from the exported G_8; near ties fall back to the faithful chains
// ------------------------------------------------------------------------------------------
constexpr int kDeepThreads = 256;

__global__ void __launch_bounds__(kDeepThreads)
pyr_deep_kernel(const __grid_constant__ PyrDesc pd, const uint32_t *__restrict__ g8, int n_images, int do5,
                TieEntry *__restrict__ tie_queue, unsigned *__restrict__ tie_count, uint8_t *__restrict__ dst) {
    const int w8 = pd.lv[0].w / 8, h8 = pd.lv[0].h / 8;
    const int w4 = pd.lv[4].w, h4 = pd.lv[4].h;
    const int n4 = max(w4 - 2, 0) * max(h4 - 2, 0);
    const int w5 = do5 ? pd.lv[5].w : 0, h5 = do5 ? pd.lv[5].h : 0;
    const int n5 = do5 ? max(w5 - 2, 0) * max(h5 - 2, 0) : 0;
    const long long total = (long long)(n4 + n5) * n_images;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int img = (int)(idx / (n4 + n5));
        int k = (int)(idx - (long long)img * (n4 + n5));
        const uint32_t *__restrict__ g = g8 + (size_t)img * w8 * h8;
        uint32_t n, half;
        int shift, lev, ox, oy;
        if (k < n4) {  // O_16[o] = sum (1,2,1)x(1,2,1) G_8[2o-1 .. 2o+1]
            lev = 4;
            ox = 1 + k % (w4 - 2);
            oy = 1 + k / (w4 - 2);
            RSVIO_CHECK(2 * oy + 1 < h8 && 2 * ox + 1 < w8);
            const uint32_t *r0 = g + (size_t)(2 * oy - 1) * w8 + (2 * ox - 1), *r1 = r0 + w8, *r2 = r1 + w8;
            n = r0[0] + 2u * r0[1] + r0[2] + 2u * (r1[0] + 2u * r1[1] + r1[2]) + r2[0] + 2u * r2[1] + r2[2];
            shift = 18;  // D = 4 * 16^4
        } else {       // O_32[o] = sum (1,2,3,4,3,2,1)^2 G_8[4o-2 .. 4o+4]
            k -= n4;
            lev = 5;
            ox = 1 + k % (w5 - 2);
            oy = 1 + k / (w5 - 2);
            n = 0;
            RSVIO_CHECK(4 * oy + 4 < h8 && 4 * ox + 4 < w8);
#pragma unroll
            for (int a = 0; a < 7; ++a) {
                const uint32_t *r = g + (size_t)(4 * oy - 2 + a) * w8 + (4 * ox - 2);
                const uint32_t row = r[0] + 2u * r[1] + 3u * r[2] + 4u * r[3] + 3u * r[4] + 2u * r[5] + r[6];
                n += (uint32_t)(a < 4 ? a + 1 : 7 - a) * row;
            }
            shift = 22;  // D = 4 * 32^4
        }
        half = 1u << (shift - 1);
        // The reference's f32 horizontal accumulation differs from N / D by at most 31 * 2^-17 = 2.4e-4 grey
        // levels at level 4 (products exact, 31 rounded adds below 256) and 63 * 2^-17 + 64 * 2^-22 = 5.0e-4
        // at level 5: the outputs within 2^-11 resp. 2^-10 of a tie could round the other way (~0.2 % of them).
        const uint32_t frac = n & ((1u << shift) - 1u);
        const uint32_t dist = frac > half ? frac - half : half - frac;
        // Those outputs are not written here: they go to a queue that the extra CTAs of the following
        // pyr_generic_kernel launch drain as 1x1 tiles with the reference's f32 chains from the base image.  (The first
        // version recomputed them on the spot, one warp per tie: a 3-5 us serial detour for the unlucky warps, which
        // decided the kernel's time -- 20-30 us for 3 MB of G_8.)
        if (dist <= (1u << (shift - (lev == 4 ? 11 : 10)))) {
            const unsigned slot = atomicAdd(tie_count, 1u);
            tie_queue[slot] = make_uint2((unsigned)img, ((unsigned)lev << 28) | ((unsigned)oy << 14) | (unsigned)ox);
        } else {
            dst[(size_t)img * pd.pyr_stride + pd.lv[lev].off + (size_t)oy * pd.lv[lev].w + ox] = (uint8_t)((n + half) >> shift);
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
        else
            cudaGetLastError();
    }
    return fn;
}

// u8 tensor (x: w, y: h, z: images) with byte strides (1, stride, pitch); box = RW x rows x 1
static bool make_tile_map(CUtensorMap *m, const uint8_t *base, int w, int h, int n, size_t stride, size_t pitch,
                          int rows, int box_w = RW) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    if (n <= 1 || pitch == 0) pitch = stride * (size_t)h;
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)(n > 0 ? n : 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)stride, (cuuint64_t)pitch};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)rows, 1u};
    const cuuint32_t es[3] = {1u, 1u, 1u};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t *>(base), dims, strides, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NA>
static cudaError_t fast_kernel_setup(int *grid, int nbuf, int *smem_bytes) {
    const int bytes = FastSmem<NA>::bytes - (2 - nbuf) * 2 * FastSmem<NA>::region;
    *smem_bytes = bytes;
    cudaError_t e = cudaFuncSetAttribute(pyr_fast_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         FastSmem<NA>::bytes);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pyr_fast_kernel<NA>, NTA, bytes);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = kNumSMs;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    *grid = sms * (occ > 0 ? occ : 1);
    return cudaSuccess;
}

struct PyrPlan {
    PyrDesc pd;
    PyrTaps taps;
    // tile descriptors of the generic kernel: [0, n_all) every level (fast path unusable),
    // [n_all, n_all + n_rest) what the fast path leaves: ring of the dyadic levels + non-dyadic levels
    TileDesc *d_tiles;
    int n_all, n_rest;
    // fast path
    int n_fast;      // deepest level on the dyadic path (0: none)
    int n_a;         // deepest level produced by pyr_fast_kernel (<= 3)
    int do4, do5;
    int tiles_x, tiles_y;
    int fast_rows;   // tile rows fetched per CTA (TMA box height)
    int fast_grid;   // persistent grid: SMs x resident CTAs per SM
    int fast_nbuf, fast_smem;  // tile buffers per CTA and the resulting dynamic shared memory
    unsigned *d_tile_counter;  // two counters used alternately
    int tile_parity;
    RingWeights ring[4];
    void *d_tables;
    // scratch for deep levels (grown on demand)
    uint32_t *d_g8;
    size_t g8_images;
    TieEntry *d_tie_queue;  // near-tie outputs of pyr_deep_kernel (capacity: every interior output of levels 4 / 5)
};

// tiles covering the output rectangle [x0, x0 + nx) x [y0, y0 + ny) of one level
static void add_rect_tiles(const PyrDesc &pd, int lev, int x0, int y0, int nx, int ny, std::vector<TileDesc> &out) {
    if (nx <= 0 || ny <= 0) return;
    const double ratio = (double)pd.lv[0].w / (double)(pd.lv[lev].w > 0 ? pd.lv[lev].w : 1);
    // intermediate columns per tile row: (tox + 1) * ratio + 8 must fit the smem budget with noy rows
    int tox = nx < 32 ? nx : 32;
    while (tox > 1 && (tox + 1) * ratio + 8 > 672) tox >>= 1;
    tox = (nx + (nx + tox - 1) / tox - 1) / ((nx + tox - 1) / tox);  // equal tiles: 23 -> 8, 8, 7 instead of 11, 11, 1
    const int cols = (int)((tox + 1) * ratio) + 8;
    // rows per tile: one round of the vertical pass (a thread = 4 columns x 1 output row) fills the CTA once, so
    // that tiles are short and numerous (no wave quantisation; rows share no work, only columns do)
    int toy = kGenThreads / ((cols + 3) / 4);
    if (toy > kGenSmemFloats / cols) toy = kGenSmemFloats / cols;
    if (toy > 64) toy = 64;
    if (toy < 1) toy = 1;
    toy = (ny + (ny + toy - 1) / toy - 1) / ((ny + toy - 1) / toy);
    for (int y = y0; y < y0 + ny; y += toy)
        for (int x = x0; x < x0 + nx; x += tox) {
            TileDesc t;
            t.lev = (int16_t)lev;
            t.ox0 = x;
            t.oy0 = y;
            t.nox = (int16_t)((x0 + nx - x) < tox ? (x0 + nx - x) : tox);
            t.noy = (int16_t)((y0 + ny - y) < toy ? (y0 + ny - y) : toy);
            t.pad = 0;
            out.push_back(t);
        }
}

static void add_level0_tiles(const PyrDesc &pd, std::vector<TileDesc> &out) {
    const int w = pd.lv[0].w, h = pd.lv[0].h;
    int rows = 16384 / (w > 0 ? w : 1);
    rows = rows < 1 ? 1 : rows;
    for (int y = 0; y < h; y += rows) {
        TileDesc t;
        t.lev = 0;
        t.ox0 = 0;
        t.oy0 = y;
        t.nox = 0;
        t.noy = (int16_t)((h - y) < rows ? (h - y) : rows);
        t.pad = 0;
        out.push_back(t);
    }
}

cudaError_t pyr_plan_create(int w, int h, int levels, cudaStream_t st, PyrPlan **out) {
    PyrPlan *p = new PyrPlan();
    p->pd = make_pyr_desc(w, h, levels);
    p->d_tables = nullptr;
    p->d_tile_counter = nullptr;
    p->d_tiles = nullptr;
    p->d_g8 = nullptr;
    p->g8_images = 0;
    p->d_tie_queue = nullptr;
    // one allocation for all tap tables
    size_t bytes = 0;
    struct Sz { int n_in, n_out, kmax; size_t off_left, off_cnt, off_w; };
    std::vector<Sz> sz;
    for (int i = 1; i < levels; ++i)
        for (int axis = 0; axis < 2; ++axis) {
            Sz s;
            s.n_in = axis == 0 ? w : h;
            s.n_out = axis == 0 ? (w >> i) : (h >> i);
            const double ratio = (double)s.n_in / (double)(s.n_out > 0 ? s.n_out : 1);
            s.kmax = (int)(2.0 * ratio) + 4;
            s.off_left = bytes;
            bytes += (size_t)s.n_out * sizeof(int);
            s.off_cnt = bytes;
            bytes += (size_t)s.n_out * sizeof(int);
            bytes = (bytes + 15) & ~(size_t)15;
            s.off_w = bytes;
            bytes += (size_t)s.n_out * s.kmax * sizeof(float);
            bytes = (bytes + 15) & ~(size_t)15;
            sz.push_back(s);
        }
    cudaError_t e = cudaSuccess;
    if (bytes) {
        e = cudaMalloc(&p->d_tables, bytes);
        if (e != cudaSuccess) {
            delete p;
            return e;
        }
    }
    int k = 0;
    for (int i = 1; i < levels; ++i)
        for (int axis = 0; axis < 2; ++axis, ++k) {
            const Sz &s = sz[k];
            TapTable t;
            char *base = (char *)p->d_tables;
            t.left = (const int *)(base + s.off_left);
            t.cnt = (const int *)(base + s.off_cnt);
            t.w = (const float *)(base + s.off_w);
            t.n_in = s.n_in;
            t.n_out = s.n_out;
            t.kmax = s.kmax;
            (axis == 0 ? p->taps.x[i] : p->taps.y[i]) = t;
            if (s.n_out > 0)
                taps_kernel<<<(s.n_out + 127) / 128, 128, 0, st>>>(s.n_in, s.n_out, s.kmax, (int *)t.left, (int *)t.cnt,
                                                                   (float *)t.w);
        }
    // dyadic fast path: level i qualifies while w and h are divisible by 2^i and the level is at
    // least 3x3 (it needs an interior); levels >= 6 overflow the u32 tent sums
    p->n_fast = 0;
    if ((w % 16) == 0 && levels >= 2) {
        for (int i = 1; i < levels && i <= 5; ++i) {
            if ((w % (1 << i)) || (h % (1 << i)) || (w >> i) < 3 || (h >> i) < 3) break;
            p->n_fast = i;
        }
    }
    p->n_a = p->n_fast < 3 ? p->n_fast : 3;
    p->do4 = p->n_fast >= 4;
    p->do5 = p->n_fast >= 5;
    p->tiles_x = (w + TW - 1) / TW;  // a tile's lanes own level-1 columns [128 k - 2, 128 k + 126); the last tile adds + 126
    p->tiles_y = (h + TH - 1) / TH;
    p->fast_rows = 0;
    p->fast_grid = kNumSMs;
    if (p->n_fast >= 1 && !encode_tiled_fn()) p->n_fast = 0;  // no TMA descriptor encoder: generic path
    p->n_a = p->n_fast < 3 ? p->n_fast : 3;
    p->do4 = p->n_fast >= 4;
    p->do5 = p->n_fast >= 5;
    if (p->n_a >= 1) {
        e = cudaMalloc((void **)&p->d_tile_counter, 4 * sizeof(unsigned));  // [0], [1]: tiles, [2]: tie queue length
        if (e == cudaSuccess) e = cudaMemsetAsync(p->d_tile_counter, 0, 4 * sizeof(unsigned), st);
        p->tile_parity = 0;
        if (e != cudaSuccess) {
            pyr_plan_destroy(p);
            return e;
        }
        const char *nb = getenv("RSVIO_PYR_NBUF");
        p->fast_nbuf = (nb && atoi(nb) == 2) ? 2 : 1;
        e = p->n_a == 1 ? fast_kernel_setup<1>(&p->fast_grid, p->fast_nbuf, &p->fast_smem)
                        : (p->n_a == 2 ? fast_kernel_setup<2>(&p->fast_grid, p->fast_nbuf, &p->fast_smem)
                                       : fast_kernel_setup<3>(&p->fast_grid, p->fast_nbuf, &p->fast_smem));
        if (e != cudaSuccess) {
            pyr_plan_destroy(p);
            return e;
        }
        p->fast_rows = p->n_a == 1 ? FastCfg<1>::ROWS : (p->n_a == 2 ? FastCfg<2>::ROWS : FastCfg<3>::ROWS);
    }
    // faithful border weights of the dyadic levels handled by pyr_fast_kernel: they depend on the
    // ratio only; take them from the x table (outputs 0, 1 and n-1) as computed on the device
    memset(p->ring, 0, sizeof(p->ring));
    for (int i = 1; i <= p->n_a; ++i) {
        const TapTable &t = p->taps.x[i];
        const int r = 1 << i, nb = r + r / 2;
        e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaMemcpy(p->ring[i].first, t.w, sizeof(float) * nb, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            e = cudaMemcpy(p->ring[i].inner, t.w + (size_t)1 * t.kmax, sizeof(float) * 2 * r, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            e = cudaMemcpy(p->ring[i].last, t.w + (size_t)(t.n_out - 1) * t.kmax, sizeof(float) * nb, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            pyr_plan_destroy(p);
            return e;
        }
    }
    {
        std::vector<TileDesc> all, rest;
        add_level0_tiles(p->pd, all);
        for (int i = 1; i < levels; ++i) {
            const int wl = p->pd.lv[i].w, hl = p->pd.lv[i].h;
            add_rect_tiles(p->pd, i, 0, 0, wl, hl, all);
            if (i > p->n_fast) {
                add_rect_tiles(p->pd, i, 0, 0, wl, hl, rest);
            } else if (i > p->n_a) {  // ring of the deep dyadic levels (levels <= n_a: inside pyr_fast_kernel)
                add_rect_tiles(p->pd, i, 0, 0, wl, 1, rest);
                add_rect_tiles(p->pd, i, 0, hl - 1, wl, 1, rest);
                add_rect_tiles(p->pd, i, 0, 1, 1, hl - 2, rest);
                add_rect_tiles(p->pd, i, wl - 1, 1, 1, hl - 2, rest);
            }
        }
        p->n_all = (int)all.size();
        p->n_rest = (int)rest.size();
        all.insert(all.end(), rest.begin(), rest.end());
        e = cudaMalloc((void **)&p->d_tiles, all.size() * sizeof(TileDesc));
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(p->d_tiles, all.data(), all.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // `all` is a local
        if (e != cudaSuccess) {
            pyr_plan_destroy(p);
            return e;
        }
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        pyr_plan_destroy(p);
        return e;
    }
    *out = p;
    return cudaSuccess;
}

void pyr_plan_destroy(PyrPlan *p) {
    if (!p) return;
    if (p->d_tables) cudaFree(p->d_tables);
    if (p->d_tile_counter) cudaFree(p->d_tile_counter);
    if (p->d_tiles) cudaFree(p->d_tiles);
    if (p->d_g8) cudaFree(p->d_g8);
    if (p->d_tie_queue) cudaFree(p->d_tie_queue);
    delete p;
}

const PyrDesc &pyr_plan_desc(const PyrPlan *p) { return p->pd; }

int pyr_plan_parity(const PyrPlan *p) { return p->tile_parity; }
void pyr_plan_set_parity(PyrPlan *p, int parity) { p->tile_parity = parity & 1; }

cudaError_t pyr_plan_reserve(PyrPlan *p, int n_images) {
    if (!p->do4 || (size_t)n_images <= p->g8_images) return cudaSuccess;
    if (p->d_g8) cudaFree(p->d_g8);
    if (p->d_tie_queue) cudaFree(p->d_tie_queue);
    p->d_g8 = nullptr;
    p->d_tie_queue = nullptr;
    const size_t per = (size_t)(p->pd.lv[0].w / 8) * (p->pd.lv[0].h / 8);
    cudaError_t e = cudaMalloc((void **)&p->d_g8, per * n_images * sizeof(uint32_t));
    if (e != cudaSuccess) return e;
    const size_t outs = (size_t)p->pd.lv[4].w * p->pd.lv[4].h + (p->do5 ? (size_t)p->pd.lv[5].w * p->pd.lv[5].h : 0);
    e = cudaMalloc((void **)&p->d_tie_queue, outs * n_images * sizeof(TieEntry));
    if (e != cudaSuccess) return e;
    p->g8_images = (size_t)n_images;
    return cudaSuccess;
}

cudaError_t pyr_build(const PyrPlan *p, const uint8_t *src0, const uint8_t *src1, int split, int n_images,
                      size_t stride, size_t pitch, uint8_t *dst, cudaStream_t st, int *n_launches) {
    if (n_images <= 0) return cudaSuccess;
    auto aligned16 = [](const void *q) { return (((size_t)q) & 15) == 0; };
    const bool fast = p->n_fast >= 1 && aligned16(src0) && aligned16(dst) && (split >= n_images || aligned16(src1)) && (stride % 16) == 0 &&
                      (n_images == 1 || (pitch % 16) == 0) && (!p->do4 || (size_t)n_images <= p->g8_images);
    // gridDim.y is limited to 65535 images per launch
    for (int first = 0; first < n_images; first += 65535) {
        const int n = n_images - first < 65535 ? n_images - first : 65535;
        const uint8_t *s0 = src0, *s1 = src1;
        int sps = split - first;
        if (sps < 0) {
            s1 = src1 + (size_t)(-sps) * pitch;
            sps = 0;
        } else {
            s0 = src0 + (size_t)first * pitch;
        }
        uint8_t *d = dst + (size_t)first * p->pd.pyr_stride;
        if (!fast) {
            dim3 grid((unsigned)p->n_all, (unsigned)n, 1);
            pyr_generic_kernel<<<grid, kGenThreads, 0, st>>>(p->pd, p->taps, p->d_tiles, p->n_all, nullptr, nullptr, s0, s1, sps,
                                                             stride, pitch, d);
            if (n_launches) ++*n_launches;
            continue;
        }
        {
        FastParams fp;
        fp.pd = p->pd;
        memcpy(fp.ring, p->ring, sizeof(fp.ring));
        fp.split = sps;
        fp.n_images = n;
        fp.dst = d;
        fp.g8 = p->do4 ? p->d_g8 : nullptr;
        fp.tiles_x = p->tiles_x;
        fp.tiles_y = p->tiles_y;
        {
            const unsigned tpi = (unsigned)(p->tiles_x * p->tiles_y);
            fp.magic_tpi = magic_for(tpi, (uint64_t)tpi * (uint64_t)n);
            fp.magic_tx = magic_for((unsigned)p->tiles_x, tpi);
        }
        fp.nbuf = p->fast_nbuf;
        fp.tie_count = p->do4 ? p->d_tile_counter + 2 : nullptr;
        {
            PyrPlan *mp = const_cast<PyrPlan *>(p);  // launches of one plan are stream ordered
            fp.tile_counter = p->d_tile_counter + mp->tile_parity;
            fp.tile_counter_next = p->d_tile_counter + (mp->tile_parity ^ 1);
            mp->tile_parity ^= 1;
        }
        CUtensorMap m0, m1;
        const int n0 = sps < n ? sps : n;               // images behind map 0
        const uint8_t *b1 = sps < n ? s1 : s0;
        const int n1 = sps < n ? n - sps : 1;
        if (!make_tile_map(&m0, n0 > 0 ? s0 : b1, p->pd.lv[0].w, p->pd.lv[0].h, n0 > 0 ? n0 : 1, stride, pitch, p->fast_rows) ||
            !make_tile_map(&m1, b1, p->pd.lv[0].w, p->pd.lv[0].h, n1, stride, pitch, p->fast_rows))
            return cudaErrorInvalidValue;
        const int total = p->tiles_x * p->tiles_y * n;
        const int blocks = total < p->fast_grid ? total : p->fast_grid;
        if (p->n_a == 1)
            pyr_fast_kernel<1><<<blocks, NTA, p->fast_smem, st>>>(fp, m0, m1);
        else if (p->n_a == 2)
            pyr_fast_kernel<2><<<blocks, NTA, p->fast_smem, st>>>(fp, m0, m1);
        else
            pyr_fast_kernel<3><<<blocks, NTA, p->fast_smem, st>>>(fp, m0, m1);
        if (n_launches) ++*n_launches;
        }
        if (p->do4) {
            const long long items = (long long)n * ((long long)p->pd.lv[4].w * p->pd.lv[4].h +
                                                    (p->do5 ? (long long)p->pd.lv[5].w * p->pd.lv[5].h : 0));
            long long blocks = (items + kDeepThreads - 1) / kDeepThreads;
            if (blocks > sm_count() * 8) blocks = sm_count() * 8;
            if (blocks < 1) blocks = 1;
            pyr_deep_kernel<<<(unsigned)blocks, kDeepThreads, 0, st>>>(p->pd, p->d_g8, n, p->do5, p->d_tie_queue,
                                                                        p->d_tile_counter + 2, d);
            if (n_launches) ++*n_launches;
        }
        if (p->n_rest > 0) {  // ring of the dyadic levels + whole non-dyadic levels
            // + the CTAs that drain the tie queue of pyr_deep_kernel: ~0.2 % of the level-4/5 outputs, ~2 us each
            const int tie_ctas = p->do4 ? (n >= 64 ? 1 : (n >= 8 ? 2 : 4)) : 0;  // x 8 warps, one tie per warp at a time
            dim3 ggrid((unsigned)(p->n_rest + tie_ctas), (unsigned)n, 1);
            pyr_generic_kernel<<<ggrid, kGenThreads, 0, st>>>(p->pd, p->taps, p->d_tiles + p->n_all, p->n_rest, p->d_tie_queue,
                                                             p->d_tile_counter + 2, s0, s1, sps, stride, pitch, d);
            if (n_launches) ++*n_launches;
        }
    }
    return cudaGetLastError();
}

}  // namespace rsvio

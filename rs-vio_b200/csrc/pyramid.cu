// pyramid.cu -- u8 image pyramid, every level resampled from the base image.
//
// Replaces build_image_pyramid (reference src/feature_tracker/feature_tracker.rs:209-220),
// i.e. image 0.25 `imageops::resize(img, w0 >> i, h0 >> i, FilterType::Triangle)` per level:
// vertical pass first into an un-rounded f32 intermediate, then the horizontal pass, clamp,
// round half away from zero.  Weights are tent(|i - c| / ratio) normalised by their f32 sum;
// accumulation is `t += pix * w` in tap order with separately rounded multiply and add
// (SURVEY.md Appendix A).  Bit-exactness is the contract, so every multiply/add that can round
// is an explicit __fmul_rn/__fadd_rn and the whole library is compiled with -fmad=false.
//
// pyr_generic_kernel: faithful path for any size ratio (also the non-power-of-two levels and
// the truncated windows at the image border).  One CTA = one tile of one level of one image;
// the vertical pass reads the base image straight from global memory (coalesced along x, each
// row re-used from L1/L2), leaves its f32 columns in shared memory and the horizontal pass
// finishes from there, so nothing but the u8 levels is ever written to HBM.
#include "common.cuh"

#include <vector>

namespace rsvio {

namespace {

constexpr int kTileRows = 8;     // output rows per CTA
constexpr int kMaxCols = 672;    // f32 intermediate columns per CTA row
constexpr int kThreads = 256;

struct TileTable {
    int first_block[kMaxLevels + 1];  // prefix of CTAs per level
    int tiles_x[kMaxLevels];
    int tox[kMaxLevels];              // output columns per tile (level >= 1); rows per block (level 0)
};

// image 0.25 sample.rs weight construction, one thread per output index
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

__device__ __forceinline__ const uint8_t *image_ptr(const uint8_t *src0, const uint8_t *src1, int split, int j,
                                                    size_t pitch) {
    return j < split ? src0 + (size_t)j * pitch : src1 + (size_t)(j - split) * pitch;
}

__global__ void __launch_bounds__(kThreads)
pyr_generic_kernel(PyrDesc pd, PyrTaps taps, TileTable tt, const uint8_t *__restrict__ src0,
                   const uint8_t *__restrict__ src1, int split, size_t stride, size_t pitch,
                   uint8_t *__restrict__ dst) {
    __shared__ float v[kTileRows][kMaxCols];
    const int img = blockIdx.y;
    const uint8_t *__restrict__ src = image_ptr(src0, src1, split, img, pitch);
    uint8_t *__restrict__ out = dst + (size_t)img * pd.pyr_stride;
    int lev = 0;
    while (lev + 1 < pd.levels && (int)blockIdx.x >= tt.first_block[lev + 1]) ++lev;
    const int blk = blockIdx.x - tt.first_block[lev];
    const int w0 = pd.lv[0].w, h0 = pd.lv[0].h;

    if (lev == 0) {  // same size: plain copy (image `resize` short-circuit)
        const int rows = tt.tox[0];
        const int y0 = blk * rows;
        const int ny = min(rows, h0 - y0);
        const bool vec = ((w0 & 15) == 0) && ((stride & 15) == 0) && ((((size_t)src) & 15) == 0);
        if (vec) {
            const int wv = w0 >> 4;
            for (int idx = threadIdx.x; idx < wv * ny; idx += kThreads) {
                const int y = idx / wv, x = idx - y * wv;
                const uint4 q = __ldg(reinterpret_cast<const uint4 *>(src + (size_t)(y0 + y) * stride) + x);
                reinterpret_cast<uint4 *>(out + (size_t)(y0 + y) * w0)[x] = q;
            }
        } else {
            for (int idx = threadIdx.x; idx < w0 * ny; idx += kThreads) {
                const int y = idx / w0, x = idx - y * w0;
                out[(size_t)(y0 + y) * w0 + x] = __ldg(src + (size_t)(y0 + y) * stride + x);
            }
        }
        return;
    }

    const int wl = pd.lv[lev].w, hl = pd.lv[lev].h;
    const TapTable tx = taps.x[lev], ty = taps.y[lev];
    const int tox = tt.tox[lev];
    const int tile_y = blk / tt.tiles_x[lev], tile_x = blk - tile_y * tt.tiles_x[lev];
    const int ox0 = tile_x * tox, oy0 = tile_y * kTileRows;
    const int nox = min(tox, wl - ox0), noy = min(kTileRows, hl - oy0);
    const int xl = tx.left[ox0];
    const int xr = tx.left[ox0 + nox - 1] + tx.cnt[ox0 + nox - 1];
    const int ncols = xr - xl;

    // vertical pass: f32, tap order, multiply and add rounded separately
    for (int idx = threadIdx.x; idx < ncols * noy; idx += kThreads) {
        const int cy = idx / ncols, cx = idx - cy * ncols;
        const int oy = oy0 + cy;
        const int top = ty.left[oy], n = ty.cnt[oy];
        const float *__restrict__ wy = ty.w + (size_t)oy * ty.kmax;
        const uint8_t *__restrict__ p = src + (size_t)top * stride + (xl + cx);
        float t = 0.0f;
        for (int k = 0; k < n; ++k) t = __fadd_rn(t, __fmul_rn((float)__ldg(p + (size_t)k * stride), wy[k]));
        v[cy][cx] = t;
    }
    __syncthreads();
    // horizontal pass, clamp, round half away from zero
    uint8_t *__restrict__ lo = out + pd.lv[lev].off;
    for (int idx = threadIdx.x; idx < nox * noy; idx += kThreads) {
        const int cy = idx / nox, cx = idx - cy * nox;
        const int ox = ox0 + cx;
        const int left = tx.left[ox] - xl, n = tx.cnt[ox];
        const float *__restrict__ wx = tx.w + (size_t)ox * tx.kmax;
        float t = 0.0f;
        for (int k = 0; k < n; ++k) t = __fadd_rn(t, __fmul_rn(v[cy][left + k], wx[k]));
        t = fminf(fmaxf(t, 0.0f), 255.0f);
        lo[(size_t)(oy0 + cy) * wl + ox] = (uint8_t)__float2uint_rz(roundf(t));
    }
}

}  // namespace

struct PyrPlan {
    PyrDesc pd;
    PyrTaps taps;
    TileTable tt;
    int total_blocks;
    void *d_tables;
};

cudaError_t pyr_plan_create(int w, int h, int levels, cudaStream_t st, PyrPlan **out) {
    PyrPlan *p = new PyrPlan();
    p->pd = make_pyr_desc(w, h, levels);
    p->d_tables = nullptr;
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
    // tiles
    int blocks = 0;
    for (int i = 0; i < levels; ++i) {
        p->tt.first_block[i] = blocks;
        const int wl = w >> i, hl = h >> i;
        if (i == 0) {
            int rows = 16384 / (w > 0 ? w : 1);
            rows = rows < 1 ? 1 : rows;
            p->tt.tox[0] = rows;
            p->tt.tiles_x[0] = 1;
            blocks += (h + rows - 1) / rows;
        } else {
            const double ratio = (double)w / (double)(wl > 0 ? wl : 1);
            int tox = 32;
            while (tox > 1 && (tox + 1) * ratio + 3 > kMaxCols) tox >>= 1;
            p->tt.tox[i] = tox;
            p->tt.tiles_x[i] = (wl + tox - 1) / tox;
            blocks += p->tt.tiles_x[i] * ((hl + kTileRows - 1) / kTileRows);
        }
    }
    for (int i = levels; i <= kMaxLevels; ++i) p->tt.first_block[i] = blocks;
    p->total_blocks = blocks;
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
    delete p;
}

const PyrDesc &pyr_plan_desc(const PyrPlan *p) { return p->pd; }

cudaError_t pyr_build(const PyrPlan *p, const uint8_t *src0, const uint8_t *src1, int split, int n_images,
                      size_t stride, size_t pitch, uint8_t *dst, cudaStream_t st, int *n_launches) {
    if (n_images <= 0) return cudaSuccess;
    // gridDim.y is limited to 65535 images per launch
    for (int first = 0; first < n_images; first += 65535) {
        const int n = n_images - first < 65535 ? n_images - first : 65535;
        // images [first, first+n): shift the two sources so that local index j maps correctly
        const uint8_t *s0 = src0, *s1 = src1;
        int sp = split - first;
        if (sp < 0) {
            s1 = src1 + (size_t)(-sp) * pitch;
            sp = 0;
        } else {
            s0 = src0 + (size_t)first * pitch;
        }
        dim3 grid((unsigned)p->total_blocks, (unsigned)n, 1);
        pyr_generic_kernel<<<grid, kThreads, 0, st>>>(p->pd, p->taps, p->tt, s0, s1, sp, stride, pitch,
                                                      dst + (size_t)first * p->pd.pyr_stride);
        if (n_launches) ++*n_launches;
    }
    return cudaGetLastError();
}

}  // namespace rsvio

// trackb.cu -- row f-2: the tracking core of the reference's second tracker (the `feature_tracker` crate,
// "crate B") and its chained f32 pyramid.  One warp per point, forward and backward pass.
//
// Replaces (paths relative to the reference repository):
//   track_points / track_point / track_point_at_level   feature_tracker/src/feature_tracker/feature_tracking.rs:16-192
//   exp_se2 (twist = [theta, vx, vy])                    feature_tracking.rs:195-219
//   Patch52::new, jacobian, residuals_ssd               feature_tracker/src/patch.rs:57-164, 244-255
//   interpolate_bicubic / d_interpolate_bicubic         feature_tracker/src/image_operations.rs:140-282
//   in_bounds                                           image_operations.rs:4-7
//   residuals_lssd, compute_intensities_and_jacobian_lssd  patch.rs:84-105, 167-216
// The kernel is the LK skeleton of tracker A with a 4x4 Catmull-Rom gather, absolute-coordinate Jacobians, an
// LM-damped Hessian and isometries instead of affine matrices.  Blur, detection and the stateful FeatureTracker are
// in detectb.cu / apib.cu.  Checked against oracle/crate_b.py (its tracking core is pinned by the reference's own
// unit tests).
#include "common.cuh"

namespace rsvio {

namespace {

constexpr int kNP = 52;

struct PyrB {
    const float *a[RSVIO_MAX_LEVELS], *b[RSVIO_MAX_LEVELS];
    int w[RSVIO_MAX_LEVELS], h[RSVIO_MAX_LEVELS];
    int levels;
};

struct Iso {  // na::Isometry2<f32>: unit complex (c, s) and translation
    float c, s, tx, ty;
};
// FMA policy of this file (as lk.cu): the library is compiled with -fmad=false, the multiply-add chains of the
// sampler and of the isometries are fused EXPLICITLY (one rounding instead of two per step).  The kernel is bound by
// the issue rate of a single warp (77 k warp instructions per track before, a third of them in the two Horner
// passes), the bar is 1e-3 px against an oracle that rounds every operation.
__device__ __forceinline__ void iso_apply(const Iso &T, float x, float y, float &ox, float &oy) {
    ox = fmaf(T.c, x, fmaf(-T.s, y, T.tx));
    oy = fmaf(T.s, x, fmaf(T.c, y, T.ty));
}

// offsets of patch.rs:259-277: rows y = 7, 5, .., -7 with 4, 6, 8, 8, 8, 8, 6, 4 samples at odd x (not halved)
__device__ __forceinline__ void pattern_b(int i, float &ox, float &oy) {
    const int start[9] = {0, 4, 10, 18, 26, 34, 42, 48, 52};
    const int half_w[8] = {3, 5, 7, 7, 7, 7, 5, 3};
    int r = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) r += (i >= start[k]) ? 1 : 0;
    int s0 = 0, hw = 3;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (r == k) {
            s0 = start[k];
            hw = half_w[k];
        }
    ox = (float)(-hw + 2 * (i - s0));
    oy = (float)(7 - 2 * r);
}

__device__ __forceinline__ void cubic_coeffs(const float f[4], float &a0, float &a1, float &a2, float &a3) {
    a0 = f[1];
    a1 = f[2] - f[0];
    a2 = fmaf(4.0f, f[2], fmaf(-5.0f, f[1], 2.0f * f[0])) - f[3];
    a3 = fmaf(3.0f, f[1] - f[2], f[3]) - f[0];
}
__device__ __forceinline__ float bicubic_1d(const float f[4], float t) {
    float a0, a1, a2, a3;
    cubic_coeffs(f, a0, a1, a2, a3);
    return fmaf(0.5f * t, fmaf(t, fmaf(t, a3, a2), a1), a0);
}
__device__ __forceinline__ float d_bicubic_1d(const float f[4], float t, float &dt) {
    float a0, a1, a2, a3;
    cubic_coeffs(f, a0, a1, a2, a3);
    dt = 0.5f * fmaf(t, fmaf(t, 3.0f * a3, 2.0f * a2), a1);
    return fmaf(0.5f * t, fmaf(t, fmaf(t, a3, a2), a1), a0);
}

// window test of interpolate_bicubic: floor(x) in [1, w-3], floor(y) in [1, h-3] (`as u32` saturates)
__device__ __forceinline__ bool window_b(int w, int h, float x, float y, int &xf, int &yf, float &tx, float &ty) {
    // floor(x) in [1, w - 3]  <=>  1 <= x < w - 2 (NaN and infinities fail the compares; w <= 3 leaves no window)
    if (!(x >= 1.0f && x < (float)(w - 2) && y >= 1.0f && y < (float)(h - 2))) return false;
    xf = (int)x;  // truncation == floor for x >= 1
    yf = (int)y;
    tx = x - (float)xf;
    ty = y - (float)yf;
    return true;
}

__device__ __forceinline__ bool interp_b(const float *__restrict__ img, int w, int h, float x, float y, float &out) {
    int xf, yf;
    float tx, ty;
    if (!window_b(w, h, x, y, xf, yf, tx, ty)) return false;
    float fy[4];
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
        const float *__restrict__ r = img + (size_t)(yf - 1 + dy) * w + (xf - 1);
        const float f[4] = {__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3)};
        fy[dy] = bicubic_1d(f, tx);
    }
    out = bicubic_1d(fy, ty);
    return true;
}

__device__ __forceinline__ bool d_interp_b(const float *__restrict__ img, int w, int h, float x, float y, float &out,
                                           float &gx, float &gy) {
    int xf, yf;
    float tx, ty;
    if (!window_b(w, h, x, y, xf, yf, tx, ty)) return false;
    float fy[4], dfy[4];
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
        const float *__restrict__ r = img + (size_t)(yf - 1 + dy) * w + (xf - 1);
        const float f[4] = {__ldg(r), __ldg(r + 1), __ldg(r + 2), __ldg(r + 3)};
        fy[dy] = d_bicubic_1d(f, tx, dfy[dy]);
    }
    // d_bicubic_1d_full (image_operations.rs:263-282)
    float dty;
    out = d_bicubic_1d(fy, ty, dty);
    const float t = ty;
    const float d0 = 0.5f * (t * (-1.0f + t * (2.0f + t * -1.0f)));
    const float d1 = 1.0f + 0.5f * (t * (t * (-5.0f + t * 3.0f)));
    const float d2 = 0.5f * (t * (1.0f + t * (4.0f + t * -3.0f)));
    const float d3 = 0.5f * (t * (t * (-1.0f + t)));
    gx = d0 * dfy[0] + d1 * dfy[1] + d2 * dfy[2] + d3 * dfy[3];
    gy = dty;
    return true;
}

// in_bounds (image_operations.rs:4-7): both non-negative (sign bit clear) and the rounded pixel inside
__device__ __forceinline__ bool in_bounds_b(int w, int h, float x, float y) {
    if (isnan(x) || isnan(y) || signbit(x) || signbit(y)) return false;
    const unsigned rx = sat_u32(roundf(x)), ry = sat_u32(roundf(y));
    return rx < (unsigned)w && ry < (unsigned)h;
}

__device__ __forceinline__ Iso exp_se2_b(float theta, float v0, float v1) {
    float sn, cs;
    if (fabsf(theta) < 0.5f) {
        // LM increments are small rotations: Taylor polynomials (truncation < 2e-11) instead of the general-range
        // sinf / cosf (the reference's libm values are not reproducible bit for bit on a GPU anyway)
        const float t2 = theta * theta;
        float ps = fmaf(t2, 2.7557319e-06f, -1.9841270e-04f);
        ps = fmaf(ps, t2, 8.3333333e-03f);
        ps = fmaf(ps, t2, -1.6666667e-01f);
        sn = fmaf(ps * t2, theta, theta);
        float pc = fmaf(t2, -2.7557319e-07f, 2.4801587e-05f);
        pc = fmaf(pc, t2, -1.3888889e-03f);
        pc = fmaf(pc, t2, 4.1666667e-02f);
        pc = fmaf(pc, t2, -0.5f);
        cs = fmaf(pc, t2, 1.0f);
    } else {
        sincosf(theta, &sn, &cs);
    }
    float diag, cross;
    if (fabsf(theta) > 1e-4f) {
        diag = sn / theta;
        cross = (1.0f - cs) / theta;
    } else {
        const float theta_sq = theta * theta;
        diag = 1.0f - theta * theta / 6.0f;
        cross = (0.5f - theta_sq / 24.0f) * theta;
    }
    Iso E;
    E.c = cs;
    E.s = sn;
    E.tx = diag * v0 - cross * v1;
    E.ty = cross * v0 + diag * v1;
    return E;
}

// track_point_at_level; false == OutOfBound.  All lanes hold the same T and take the same exits.
// cost: 0 SSD (patch.rs:69-83, 119-164), 1 LSSD (patch.rs:84-105, 167-216: target samples divided by their mean, the
// template intensities stay raw, quotient-rule Jacobian)
__device__ bool track_level_b(const float *__restrict__ img0, const float *__restrict__ img1, int w, int h, float cx,
                              float cy, Iso &T, int max_it, float lambda, int cost, int lane) {
    float sx[2], sy[2], inten[2], J[2][3];
    bool has[2], valid[2];
    float g_x[2], g_y[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = lane + 32 * k;
        has[k] = i < kNP;
        float ox, oy;
        pattern_b(has[k] ? i : 0, ox, oy);
        sx[k] = cx + ox;
        sy[k] = cy + oy;
        float v = 0.0f, gx = 0.0f, gy = 0.0f;
        valid[k] = has[k] && d_interp_b(img0, w, h, sx[k], sy[k], v, gx, gy);
        inten[k] = valid[k] ? v : 0.0f;
        g_x[k] = gx;
        g_y[k] = gy;
    }
    // an out-of-bounds sample keeps the gradient of the last valid sample BEFORE it in sample order (the
    // reference reuses its `dimg_dpixel` buffer, patch.rs:129-150); zero when there is none
    const unsigned long long vmask = (unsigned long long)__ballot_sync(0xffffffffu, valid[0]) |
                                     ((unsigned long long)__ballot_sync(0xffffffffu, valid[1]) << 32);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const int i = lane + 32 * k;
        const unsigned long long below = vmask & ((1ull << i) - 1ull);
        const int j = below ? 63 - __clzll((long long)below) : -1;   // uniform shuffles below: every lane takes part
        const float a0 = __shfl_sync(0xffffffffu, g_x[0], j < 0 ? 0 : (j & 31)), a1 = __shfl_sync(0xffffffffu, g_x[1], j < 0 ? 0 : (j & 31));
        const float b0 = __shfl_sync(0xffffffffu, g_y[0], j < 0 ? 0 : (j & 31)), b1 = __shfl_sync(0xffffffffu, g_y[1], j < 0 ? 0 : (j & 31));
        float gx = g_x[k], gy = g_y[k];
        if (has[k] && !valid[k]) {
            gx = j < 0 ? 0.0f : (j < 32 ? a0 : a1);
            gy = j < 0 ? 0.0f : (j < 32 ? b0 : b1);
        }
        g_x[k] = has[k] ? gx : 0.0f;
        g_y[k] = has[k] ? gy : 0.0f;
    }
    if (cost == 1) {  // dr/dpixel = (dI * mean - I * mean_dI) / mean^2
        const float mean_i = warp_sum(inten[0] + inten[1]) / (float)kNP;
        const float mean_gx = warp_sum(g_x[0] + g_x[1]) / (float)kNP, mean_gy = warp_sum(g_y[0] + g_y[1]) / (float)kNP;
        const float m2 = mean_i * mean_i;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            g_x[k] = has[k] ? (g_x[k] * mean_i - inten[k] * mean_gx) / m2 : 0.0f;
            g_y[k] = has[k] ? (g_y[k] * mean_i - inten[k] * mean_gy) / m2 : 0.0f;
        }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        // [gx gy] * [[-py, 1, 0], [px, 0, 1]]
        J[k][0] = has[k] ? g_x[k] * (-sy[k]) + g_y[k] * sx[k] : 0.0f;
        J[k][1] = g_x[k];
        J[k][2] = g_y[k];
    }
    // H = lambda I + J^T J, inverse by cofactors (nalgebra try_inverse of a 3x3)
    float h00 = 0, h01 = 0, h02 = 0, h11 = 0, h12 = 0, h22 = 0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        h00 += J[k][0] * J[k][0];
        h01 += J[k][0] * J[k][1];
        h02 += J[k][0] * J[k][2];
        h11 += J[k][1] * J[k][1];
        h12 += J[k][1] * J[k][2];
        h22 += J[k][2] * J[k][2];
    }
    float hv[8] = {h00, h01, h02, h11, h12, h22, 0.0f, 0.0f};
    warp_sum8(hv, lane);  // 9 shuffles for the six sums instead of 30
    const float m11 = lambda + hv[0], m12 = hv[1], m13 = hv[2];
    const float m22 = lambda + hv[3], m23 = hv[4], m33 = lambda + hv[5];
    const float m21 = m12, m31 = m13, m32 = m23;
    const float mn1 = m22 * m33 - m32 * m23, mn2 = m21 * m33 - m31 * m23, mn3 = m21 * m32 - m31 * m22;
    const float det = m11 * mn1 - m12 * mn2 + m13 * mn3;
    float Hi[3][3];
    Hi[0][0] = mn1 / det;
    Hi[0][1] = (m13 * m32 - m33 * m12) / det;
    Hi[0][2] = (m12 * m23 - m22 * m13) / det;
    Hi[1][0] = -mn2 / det;
    Hi[1][1] = (m11 * m33 - m31 * m13) / det;
    Hi[1][2] = (m13 * m21 - m23 * m11) / det;
    Hi[2][0] = mn3 / det;
    Hi[2][1] = (m12 * m31 - m32 * m11) / det;
    Hi[2][2] = (m11 * m22 - m21 * m12) / det;

    for (int it = 0; it < max_it; ++it) {
        float b0 = 0.0f, b1 = 0.0f, b2 = 0.0f;
        float tv[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            float px, py, v = 0.0f;
            iso_apply(T, sx[k], sy[k], px, py);
            const bool ok = has[k] && interp_b(img1, w, h, px, py, v);
            tv[k] = ok ? v : 0.0f;
        }
        if (cost == 1) {  // residuals_lssd: target samples divided by their mean (warp-uniform branch)
            const float tmean = warp_sum(tv[0] + tv[1]) / (float)kNP;
            tv[0] = tv[0] / tmean;
            tv[1] = tv[1] / tmean;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float res = has[k] ? tv[k] - inten[k] : 0.0f;
            b0 += J[k][0] * res;
            b1 += J[k][1] * res;
            b2 += J[k][2] * res;
        }
        float b3 = 0.0f;
        warp_sum4(b0, b1, b2, b3, lane);  // 9 shuffles for the three sums instead of 15
        const float t0 = Hi[0][0] * b0 + Hi[0][1] * b1 + Hi[0][2] * b2;
        const float t1 = Hi[1][0] * b0 + Hi[1][1] * b1 + Hi[1][2] * b2;
        const float t2 = Hi[2][0] * b0 + Hi[2][1] * b1 + Hi[2][2] * b2;
        const Iso E = exp_se2_b(-t0, -t1, -t2);
        Iso N;  // T * E
        N.c = T.c * E.c - T.s * E.s;
        N.s = T.s * E.c + T.c * E.s;
        iso_apply(T, E.tx, E.ty, N.tx, N.ty);
        T = N;
        float ix, iy;
        iso_apply(T, cx, cy, ix, iy);
        if (!in_bounds_b(w, h, ix, iy)) return false;
        if (sqrtf(t0 * t0 + t1 * t1 + t2 * t2) < 1e-3f) break;
    }
    return true;
}

__device__ bool track_point_b(const PyrB &P, bool forward, float px, float py, Iso &T, int max_it, float lambda, int cost,
                              int lane) {
    const float w0 = (float)P.w[0], h0 = (float)P.h[0];
    T.c = 1.0f;
    T.s = 0.0f;
    T.tx = T.ty = 0.0f;
    for (int lev = P.levels - 1; lev >= 0; --lev) {
        const int wl = P.w[lev], hl = P.h[lev];
        const float scx = (float)wl / w0, scy = (float)hl / h0;
        const float cx = scx * (px + 0.5f) - 0.5f, cy = scy * (py + 0.5f) - 0.5f;
        const float *img0 = forward ? P.a[lev] : P.b[lev], *img1 = forward ? P.b[lev] : P.a[lev];
        if (!track_level_b(img0, img1, wl, hl, cx, cy, T, max_it, lambda, cost, lane)) return false;
        if (lev > 0) {
            T.tx = T.tx * ((float)P.w[lev - 1] / (float)wl);
            T.ty = T.ty * ((float)P.h[lev - 1] / (float)hl);
        }
    }
    return true;
}

__global__ void __launch_bounds__(128)
trackb_kernel(const __grid_constant__ PyrB P, const float2 *__restrict__ pts, int n, int max_it, float lambda, int cost,
              float4 *__restrict__ out_T, int32_t *__restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= n) return;
    const float2 p = pts[warp];
    Iso T0, T1;
    int st = 0;
    if (!track_point_b(P, true, p.x, p.y, T0, max_it, lambda, cost, lane)) {
        st = 1;  // forward pass left the image
    } else {
        float qx, qy;
        iso_apply(T0, p.x, p.y, qx, qy);
        if (!track_point_b(P, false, qx, qy, T1, max_it, lambda, cost, lane)) {
            st = 2;  // backward pass left the image
        } else {
            float rx, ry;
            iso_apply(T1, qx, qy, rx, ry);
            const float dx = p.x - rx, dy = p.y - ry;
            if (!(sqrtf(dx * dx + dy * dy) < 2.0f)) st = 3;  // forward-backward test (feature_tracking.rs:52-56)
        }
    }
    if (lane == 0) {
        out_T[warp] = st == 1 ? make_float4(1.0f, 0.0f, 0.0f, 0.0f) : make_float4(T0.c, T0.s, T0.tx, T0.ty);
        status[warp] = st;
    }
}

}  // namespace

cudaError_t trackb_points(const float *const *d_a, const float *const *d_b, const int *w, const int *h, int levels,
                          const float *d_pts, int n, int max_it, float lambda, int cost, float *d_out_T4,
                          int32_t *d_status, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    PyrB P{};
    P.levels = levels;
    for (int i = 0; i < levels; ++i) {
        P.a[i] = d_a[i];
        P.b[i] = d_b[i];
        P.w[i] = w[i];
        P.h[i] = h[i];
    }
    const int warps_per_block = 4;
    trackb_kernel<<<(n + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, st>>>(
        P, reinterpret_cast<const float2 *>(d_pts), n, max_it, lambda, cost, reinterpret_cast<float4 *>(d_out_T4), d_status);
    return cudaGetLastError();
}

}  // namespace rsvio

// ---- crate B's pyramid: level l = resize(level l-1, round(w0 / ratio^l), round(h0 / ratio^l), Triangle) on f32 ----
// (feature_tracker/src/image_operations.rs:47-78; `image` 0.25.9 resize on Luma<f32>, not vendored: restated as
// SURVEY.md Appendix A -- vertical pass first, weights tent((i - c') / sratio) normalised by their f32 sum,
// `t += pix * w` in tap order, no rounding, clamp to [0, 1] after the horizontal pass.  PARITY UNPINNED.)
// Correctness-first: one thread per output sample, the window and its weights recomputed per thread.
namespace rsvio {

namespace {

struct Taps {
    int left, right;
    float cc, sratio, sum;
};
__device__ __forceinline__ Taps taps_for(int o, int n_in, int n_out) {
    Taps t;
    const float ratio = (float)n_in / (float)n_out;
    t.sratio = fmaxf(ratio, 1.0f);
    const float support = 1.0f * t.sratio;
    const float c = ((float)o + 0.5f) * ratio;
    const long long l0 = (long long)floorf(c - support), r0 = (long long)ceilf(c + support);
    t.left = (int)min(max(l0, 0ll), (long long)n_in - 1);
    t.right = (int)min(max(r0, (long long)t.left + 1), (long long)n_in);
    t.cc = c - 0.5f;
    float s = 0.0f;
    for (int i = t.left; i < t.right; ++i) {
        const float xx = fabsf(((float)i - t.cc) / t.sratio);
        s = s + (xx < 1.0f ? 1.0f - xx : 0.0f);
    }
    t.sum = s;
    return t;
}
__device__ __forceinline__ float tap_weight(const Taps &t, int i) {
    const float xx = fabsf(((float)i - t.cc) / t.sratio);
    return (xx < 1.0f ? 1.0f - xx : 0.0f) / t.sum;
}

// src: h_in x w, dst: h_out x w
__global__ void resize_f32_vertical(const float *__restrict__ src, int w, int h_in, int h_out, float *__restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, o = blockIdx.y;
    if (x >= w) return;
    const Taps t = taps_for(o, h_in, h_out);
    float acc = 0.0f;
    for (int i = t.left; i < t.right; ++i) acc = acc + src[(size_t)i * w + x] * tap_weight(t, i);
    dst[(size_t)o * w + x] = acc;
}
// src: h x w_in, dst: h x w_out, clamped to [0, 1]
__global__ void resize_f32_horizontal(const float *__restrict__ src, int w_in, int w_out, int h, float *__restrict__ dst) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (o >= w_out) return;
    const Taps t = taps_for(o, w_in, w_out);
    float acc = 0.0f;
    for (int i = t.left; i < t.right; ++i) acc = acc + src[(size_t)y * w_in + i] * tap_weight(t, i);
    dst[(size_t)y * w_out + o] = fminf(fmaxf(acc, 0.0f), 1.0f);
}

}  // namespace

cudaError_t resize_triangle_f32(const float *d_src, int w_in, int h_in, float *d_tmp, float *d_dst, int w_out, int h_out,
                                cudaStream_t st) {
    if (w_in < 1 || h_in < 1 || w_out < 1 || h_out < 1) return cudaErrorInvalidValue;
    resize_f32_vertical<<<dim3((unsigned)((w_in + 127) / 128), (unsigned)h_out, 1), 128, 0, st>>>(d_src, w_in, h_in, h_out, d_tmp);
    resize_f32_horizontal<<<dim3((unsigned)((w_out + 127) / 128), (unsigned)h_out, 1), 128, 0, st>>>(d_tmp, w_in, w_out, h_out, d_dst);
    return cudaGetLastError();
}

}  // namespace rsvio

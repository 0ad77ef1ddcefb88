// lk.cu -- 52-point-pattern inverse-compositional SE(2) Lucas-Kanade, one warp per track.
//
// Replaces, fused into one launch per track_points call:
//   track_points          reference src/feature_tracker/feature_tracker.rs:252-291
//   track_one_point       feature_tracker.rs:292-342
//   track_point_at_level  feature_tracker.rs:344-395
//   Pattern52::new / set_data_jac_se2 / residual      src/feature_tracker/patch.rs:75-232
//   image_grad / inbound / se2_exp_matrix             src/feature_tracker/image_utilities.rs:5-106
//
// Mapping: a warp owns one (point, forward+backward) chain.  Lane l holds pattern samples l and
// l + 32 (the second one only for l < 20): template intensities, the three rows of H^-1 J^T and
// the bilinear gathers live in registers; sums over the 52 samples are xor-butterfly shuffles
// (every lane ends with the same value, so the 3x3 Cholesky solve, the SE(2) exponential and all
// accept/reject decisions are computed redundantly per lane and stay warp-uniform).
// Per-sample arithmetic follows the reference operation by operation (f32, no FMA contraction:
// the library is built with -fmad=false); only the order of the 52-term sums differs (tree
// instead of sequential), which is why positions agree with the CPU oracle to ~1e-5 px rather
// than bit for bit.  Tensor cores are not used: nothing here is a dense contraction.
#include <cstdlib>

#include "common.cuh"

namespace rsvio {

namespace {

constexpr float kEps = 1.1920929e-07f;  // f32::EPSILON

// FMA policy of this file.  The library is compiled with -fmad=false (the pyramid must be bit-exact), so nothing is
// contracted behind our back; the LK bar is 1e-3 px, not bit-exactness, and the kernel is issue-bound, so the
// multiply-add chains below are fused EXPLICITLY (one rounding instead of two per step: never less accurate than
// the reference's separately rounded chain).  -DRSVIO_LK_FMA=0 restores the reference's operation-by-operation
// rounding for A/B runs.
#ifndef RSVIO_LK_FMA
#define RSVIO_LK_FMA 1
#endif
__device__ __forceinline__ float mad(float a, float b, float c) {  // a * b + c
#if RSVIO_LK_FMA
    return fmaf(a, b, c);
#else
    return a * b + c;
#endif
}
// w00 p00 + w01 p01 + w10 p10 + w11 p11 in the reference's order (image_utilities.rs:30, patch.rs:203)
__device__ __forceinline__ float bilin(float w00, float w01, float w10, float w11, float p00, float p01, float p10,
                                       float p11) {
    return mad(w11, p11, mad(w10, p10, mad(w01, p01, w00 * p00)));
}
// u8 pixel -> f32.  Left to itself the compiler narrows the conversion to I2F.U16, which issues on the quarter-rate XU
// pipe (the busiest pipe of this kernel: 45 % with F2I, MUFU and POPC); behind an opaque 32-bit value it is
// I2FP.F32.U32 on the FMA pipe.  Exact either way.
#ifndef RSVIO_LK_I2FP
#define RSVIO_LK_I2FP 1
#endif
__device__ __forceinline__ float px_f32(const uint8_t *__restrict__ p) {
#if RSVIO_LK_I2FP
    uint32_t v = __ldg(p);
    asm volatile("mov.b32 %0, %0;" : "+r"(v));
    return __uint2float_rn(v);
#else
    return (float)__ldg(p);
#endif
}
// floor(v) for 0 <= v < 2^23 without F2I / I2F: adding 2^23 with round-down leaves floor(v) in the low mantissa bits
// (bit pattern 0x4B000000 + floor(v)), and subtracting 2^23 again gives floor(v) as a float, both exactly.
// A/B on the C5 step (three interleaved runs each, one box): I2F.U16 baseline 368.7 us, pixels through I2FP 356.0 us,
// I2FP + this floor 362.8 us -- taking F2I off the XU pipe buys nothing and the sequence is 2 instructions longer
// per sample, so it is off; the code stays for the record (bit-identical results either way).
#ifndef RSVIO_LK_FLOOR
#define RSVIO_LK_FLOOR 0
#endif
constexpr float kTwo23 = 8388608.0f;
[[maybe_unused]] __device__ __forceinline__ float floor_bits(float v) { return __fadd_rd(v, kTwo23); }
constexpr int kWarpsPerBlock = 4;
constexpr int kMaxGroupsSmem = 2048;  // groups whose live-point prefix fits the shared-memory table

struct Img {
    const uint8_t *__restrict__ p;
    int w, h;
};

// pattern sample i -> offset in pixels (PATTERN_RAW / pattern_scale_down, patch.rs:19-72,126-131)
__device__ __forceinline__ void pattern_offset(int i, float &ox, float &oy) {
    const int start[9] = {0, 4, 10, 18, 26, 34, 42, 48, 52};
    const int half_w[8] = {3, 5, 7, 7, 7, 7, 5, 3};
    int r = 0;
#pragma unroll
    for (int k = 1; k < 8; ++k) r += (i >= start[k]) ? 1 : 0;
    int s = 0, hw = 3;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (r == k) {
            s = start[k];
            hw = half_w[k];
        }
    const int x = -hw + 2 * (i - s);
    const int y = 7 - 2 * r;
    ox = (float)x / 2.0f;
    oy = (float)y / 2.0f;
}

// hardware reciprocal / reciprocal square root with one Newton step (<= 1 ulp for normal arguments): the IEEE
// sequences behind __frcp_rn / sqrtf cost 10-25 instructions each with their range handling
[[maybe_unused]] __device__ __forceinline__ float rcp_nr(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
[[maybe_unused]] __device__ __forceinline__ float rsqrt_nr(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(0.5f * r, fmaf(-x * r, r, 1.0f), r);
}

// image_grad -- image_utilities.rs:5-66 (same products and sum order).  `ok` == the sample passed the
// inbound test; otherwise the 12 gathers are redirected to clamped in-image addresses (results unused),
// which keeps the warp free of divergent branches.
__device__ __forceinline__ void image_grad(const Img &im, float x, float y, bool ok, float &v, float &gx, float &gy) {
    // floor; a sample that passed inbound(.., 2) has 1 <= ix <= w - 3 and 1 <= iy <= h - 3, so the 4x4
    // neighbourhood needs no clamping; rejected samples read the neighbourhood of (1, 1) (the caller
    // guarantees w, h >= 5)
    const uint32_t W = (uint32_t)im.w;
#if RSVIO_LK_FLOOR
    // floor without the XU pipe (see floor_bits): the low mantissa bits of yx, yy are the pixel indices
    const float yx = ok ? floor_bits(x) : kTwo23 + 1.0f, yy = ok ? floor_bits(y) : kTwo23 + 1.0f;
    const float dx = x - (yx - kTwo23), dy = y - (yy - kTwo23);
    const uint32_t ix = __float_as_uint(yx) & 0x7fffffu, iy = __float_as_uint(yy) & 0x7fffffu;
#else
    const uint32_t ix = ok ? __float2uint_rd(x) : 1u, iy = ok ? __float2uint_rd(y) : 1u;
    const float dx = x - (float)ix, dy = y - (float)iy;
#endif
    const float ddx = 1.0f - dx, ddy = 1.0f - dy;
    const float w00 = ddx * ddy, w01 = ddx * dy, w10 = dx * ddy, w11 = dx * dy;
    RSVIO_CHECK(ix >= 1u && ix + 2u < W && iy >= 1u && iy + 2u < (uint32_t)im.h);  // the 4x4 neighbourhood minus corners
    const uint8_t *__restrict__ r0 = im.p + (iy * W + ix);  // 32-bit offset inside one level
    const uint8_t *__restrict__ rm = r0 - W;
    const uint8_t *__restrict__ r1 = r0 + W;
    const uint8_t *__restrict__ r2 = r1 + W;
    const float p00 = px_f32(r0), p10 = px_f32(r0 + 1), p01 = px_f32(r1), p11 = px_f32(r1 + 1);
    const float pm0 = px_f32(r0 - 1), pm1 = px_f32(r1 - 1);
    const float p20 = px_f32(r0 + 2), p21 = px_f32(r1 + 2);
    const float p0m = px_f32(rm), p1m = px_f32(rm + 1);
    const float p02 = px_f32(r2), p12 = px_f32(r2 + 1);
    v = bilin(w00, w01, w10, w11, p00, p01, p10, p11);
    const float res_mx = bilin(w00, w01, w10, w11, pm0, pm1, p00, p01);
    const float res_px = bilin(w00, w01, w10, w11, p10, p11, p20, p21);
    gx = 0.5f * (res_px - res_mx);
    const float res_my = bilin(w00, w01, w10, w11, p0m, p00, p1m, p10);
    const float res_py = bilin(w00, w01, w10, w11, p01, p02, p11, p12);
    gy = 0.5f * (res_py - res_my);
}

// nalgebra Cholesky::new + solve_mut(identity) for a 3x3 -- patch.rs:143-147
__device__ __forceinline__ bool chol3_inverse(const float H[3][3], float Hinv[3][3]) {
    float m[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) m[i][j] = H[i][j];
    bool ok = true;
    float inv[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int k = 0; k < j; ++k) {
            const float factor = -m[j][k];
#pragma unroll
            for (int i = j; i < 3; ++i) m[i][j] = factor * m[i][k] + m[i][j];
        }
        const float diag = m[j][j];
        ok = ok && (diag != 0.0f) && (diag >= 0.0f);
#if RSVIO_LK_FMA
        inv[j] = rsqrt_nr(diag);  // only 1 / L_jj is used below (the divisions by the pivot are multiplications)
#else
        const float denom = sqrtf(diag);
        m[j][j] = denom;
        inv[j] = __frcp_rn(denom);  // one IEEE reciprocal per pivot, multiplications below
#endif
#pragma unroll
        for (int i = j + 1; i < 3; ++i) m[i][j] = m[i][j] * inv[j];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float b[3] = {c == 0 ? 1.0f : 0.0f, c == 1 ? 1.0f : 0.0f, c == 2 ? 1.0f : 0.0f};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float coeff = b[i] * inv[i];
            b[i] = coeff;
#pragma unroll
            for (int r = i + 1; r < 3; ++r) b[r] = (-coeff) * m[r][i] + b[r];
        }
        b[2] = (b[2] - 0.0f) * inv[2];
        b[1] = (b[1] - m[2][1] * b[2]) * inv[1];
        b[0] = (b[0] - (m[1][0] * b[1] + m[2][0] * b[2])) * inv[0];
#pragma unroll
        for (int i = 0; i < 3; ++i) Hinv[i][c] = b[i];
    }
    return ok;
}

// se2_exp_matrix -- image_utilities.rs:82-106
__device__ __forceinline__ void se2_exp(const float a[3], float m[6]) {
    const float theta = a[2];
    float sn, cs;
    if (fabsf(theta) < 0.5f) {
        // LK increments are small rotations: Taylor polynomials (truncation < 2e-11, i.e. below half an ulp of the
        // result) instead of the general-range sincosf; explicitly fused, the reference's libm sinf/cosf are not
        // reproducible bit for bit anyway
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
    float s_by_t, omc_by_t;
    if (fabsf(theta) < kEps) {
        const float t2 = theta * theta;
        s_by_t = 1.0f - 1.0f / 6.0f * t2;
        omc_by_t = 0.5f * theta - 1.0f / 24.0f * theta * t2;
    } else {
        // The reference's `(1 - cos t) / t` (image_utilities.rs:96) cancels: for t = 1e-3 the numerator keeps 3
        // digits, and that rounding noise (up to 6e-8 / t, times the translation increment) is part of what the
        // reference computes.  It is reproduced here from the correctly rounded f32 cosine -- a cancellation-free
        // series would be more accurate and 1e-4 px away from the reference (measured: 28 of 9645 points off).
#if RSVIO_LK_FMA
        const float inv_t = rcp_nr(theta);     // |t| >= EPSILON: normal range; <= 1 ulp
#else
        const float inv_t = __frcp_rn(theta);  // one IEEE reciprocal instead of two divisions
#endif
        s_by_t = sn * inv_t;
        omc_by_t = (1.0f - cs) * inv_t;
    }
    m[0] = cs;
    m[1] = -sn;
    m[3] = sn;
    m[4] = cs;
    m[2] = mad(s_by_t, a[0], -(omc_by_t * a[1]));
    m[5] = mad(omc_by_t, a[0], s_by_t * a[1]);
}

__device__ __forceinline__ bool finite3(const float a[3]) {
    return isfinite(a[0]) && isfinite(a[1]) && isfinite(a[2]);
}

// Per-lane state of the lane's two pattern samples
// The lane's two pattern offsets live in a shared-memory table (float2 per sample): the compiler
// otherwise rematerialises them from the lane id inside the iteration loop (~20 instructions per sample).
struct Lane {
    uint32_t pat;  // shared-window address of table[lane]; sample k is 256 k bytes further
    bool has[2];
    // reloaded at every use (volatile): one live register instead of four, and no address rematerialisation
    __device__ __forceinline__ float2 offset(int k) const {
        float2 v;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(pat + 256u * (uint32_t)k));
        return v;
    }
};

// track_one_point (feature_tracker.rs:292-342): coarse-to-fine, template from pyramid A at the
// un-rotated source position, running transform keeps its rotation across levels.
__device__ int track_one_point(const PyrDesc &pd, const uint8_t *__restrict__ A, const uint8_t *__restrict__ B,
                               const float T0[6], float T1[6], int max_it, float thr, uint32_t floor_bias, const Lane &ln,
                               int lane, unsigned &n_iters, unsigned &n_templates) {
    T1[0] = 1.0f; T1[1] = 0.0f; T1[2] = T0[2];
    T1[3] = 0.0f; T1[4] = 1.0f; T1[5] = T0[5];
    for (int lev = pd.levels - 1; lev >= 0; --lev) {
        // the reference divides by 2^lev (feature_tracker.rs:305-313): a power of two, so multiplying by its
        // exact reciprocal gives the same bits without four IEEE division sequences per level
        const float s = (float)(1 << lev), inv_s = __int_as_float((127 - lev) << 23);
        T1[2] = T1[2] * inv_s;
        T1[5] = T1[5] * inv_s;
        const Img ia{A + pd.lv[lev].off, pd.lv[lev].w, pd.lv[lev].h};
        const Img ib{B + pd.lv[lev].off, pd.lv[lev].w, pd.lv[lev].h};
        const float pxx = T0[2] * inv_s, pyy = T0[5] * inv_s;
        // a level smaller than 5x5 has no sample that passes inbound(.., 2): the template is invalid
        // (patch.rs:95,146: n = 0, mean = NaN), and the unclamped gathers below need w, h >= 5
        if (ia.w < 5 || ia.h < 5) {
            ++n_templates;
            return RSVIO_FAIL_TEMPLATE;
        }

        // ---- Pattern52::new (patch.rs:124-162) --------------------------------------------
        float d[2], J[2][3];
        float sum = 0.0f, g0 = 0.0f, g1 = 0.0f, g2 = 0.0f;
        int n_valid = 0;
        unsigned tmask[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            d[k] = -1.0f;
            J[k][0] = J[k][1] = J[k][2] = 0.0f;
            const float2 o = ln.offset(k);
            const float x = pxx + o.x, y = pyy + o.y;
            const bool ok = ln.has[k] && inbound(ia.w, ia.h, x, y, 2);
            float v, gx, gy;
            image_grad(ia, x, y, ok, v, gx, gy);
            if (ok) {
                d[k] = v;
                // grad^T [[1, 0, -o_y], [0, 1, o_x]] (patch.rs:100-101); the products with 1 and 0 of the first two
                // columns are exact for finite gradients and are not spelled out
                J[k][0] = gx;
                J[k][1] = gy;
                J[k][2] = mad(gy, o.x, gx * (-o.y));
                sum += v;
                g0 += J[k][0];
                g1 += J[k][1];
                g2 += J[k][2];
            }
            tmask[k] = __ballot_sync(0xffffffffu, ok);  // samples valid in the template (data >= 0)
            n_valid += __popc(tmask[k]);
        }
        ++n_templates;
        warp_sum4(sum, g0, g1, g2, lane);
#if RSVIO_LK_FMA
        const float inv_sum = rcp_nr(sum);  // sum == 0 (black patch): inf -> NaN below -> template invalid, as the reference's mean test
        const float mean_inv = (float)n_valid * inv_sum;
#else
        const float mean_inv = (float)n_valid / sum;
        const float inv_sum = __frcp_rn(sum);
#endif
        float h00 = 0.f, h01 = 0.f, h02 = 0.f, h11 = 0.f, h12 = 0.f, h22 = 0.f;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const bool first = k == 0;  // every lane holds sample 0: its products start the sums (no 0 + x)
            if (d[k] >= 0.0f) {
#if RSVIO_LK_FMA
                const float dn = -(d[k] * inv_sum);
                J[k][0] = fmaf(g0, dn, J[k][0]);
                J[k][1] = fmaf(g1, dn, J[k][1]);
                J[k][2] = fmaf(g2, dn, J[k][2]);
#else
                J[k][0] += -(g0 * d[k] * inv_sum);
                J[k][1] += -(g1 * d[k] * inv_sum);
                J[k][2] += -(g2 * d[k] * inv_sum);
#endif
                d[k] *= mean_inv;
            } else {
                J[k][0] = J[k][1] = J[k][2] = 0.0f;
            }
            J[k][0] *= mean_inv;
            J[k][1] *= mean_inv;
            J[k][2] *= mean_inv;
            if (first) {
                h00 = J[k][0] * J[k][0];
                h01 = J[k][0] * J[k][1];
                h02 = J[k][0] * J[k][2];
                h11 = J[k][1] * J[k][1];
                h12 = J[k][1] * J[k][2];
                h22 = J[k][2] * J[k][2];
            } else if (ln.has[k]) {
                h00 = mad(J[k][0], J[k][0], h00);
                h01 = mad(J[k][0], J[k][1], h01);
                h02 = mad(J[k][0], J[k][2], h02);
                h11 = mad(J[k][1], J[k][1], h11);
                h12 = mad(J[k][1], J[k][2], h12);
                h22 = mad(J[k][2], J[k][2], h22);
            }
        }
        float H[3][3], Hinv[3][3];
        {
            float hv[8] = {h00, h01, h02, h11, h12, h22, 0.0f, 0.0f};
            warp_sum8(hv, lane);
            H[0][0] = hv[0];
            H[0][1] = H[1][0] = hv[1];
            H[0][2] = H[2][0] = hv[2];
            H[1][1] = hv[3];
            H[1][2] = H[2][1] = hv[4];
            H[2][2] = hv[5];
        }
        bool valid = chol3_inverse(H, Hinv);
        float M[2][3];
        bool fin = true;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                M[k][r] = mad(Hinv[r][2], J[k][2], mad(Hinv[r][1], J[k][1], Hinv[r][0] * J[k][0]));
                fin = fin && (!ln.has[k] || isfinite(M[k][r]));
            }
            fin = fin && (!ln.has[k] || isfinite(d[k]));
        }
        // mean = sum / n > EPSILON (patch.rs:146,156) without the division: n = 0 gives 0 > 0, invalid as the
        // reference's NaN > EPSILON; the two forms differ only for a patch mean within an ulp of 1.2e-7 grey levels
        valid = valid && (sum > kEps * (float)n_valid) && __all_sync(0xffffffffu, fin);
        if (!valid) return RSVIO_FAIL_TEMPLATE;

        // ---- track_point_at_level (feature_tracker.rs:344-395) ----------------------------
        float wlim = (float)((uint32_t)ib.w - 2u), hlim = (float)((uint32_t)ib.h - 2u);
        asm volatile("mov.b32 %0, %0;" : "+f"(wlim));  // kept in registers across the iterations, as the base and pitch below
        asm volatile("mov.b32 %0, %0;" : "+f"(hlim));
        // `inc.norm() < thr` (feature_tracker.rs:374) is never true for thr <= 0
        float thr2 = thr > 0.0f ? thr * thr : 0.0f;
        asm volatile("mov.b32 %0, %0;" : "+f"(thr2));  // in a register across the iterations, not re-derived from the constant bank
        // the level base and row pitch stay in registers across the iterations (opaque to rematerialisation)
        const uint8_t *ibp = ib.p;
        uint32_t ibw = (uint32_t)ib.w;
#if RSVIO_LK_FLOOR
        const uint8_t *ibq = ibp - (uint64_t)floor_bias * ((uint64_t)ibw + 1u);  // see the gather below
        asm volatile("mov.u64 %0, %0;" : "+l"(ibq));
#else
        asm volatile("mov.u64 %0, %0;" : "+l"(ibp));
#endif
        asm volatile("mov.u32 %0, %0;" : "+r"(ibw));
        for (int it = 0; it < max_it; ++it) {
            ++n_iters;
            float r[2];
            float rsum = 0.0f;
            int nv = 0;
            unsigned cmask[2];
            bool okk[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float2 o = ln.offset(k);
                // (R o) + t in the reference's association (feature_tracker.rs:356-360): the rotated offset (a few px)
                // is summed first and then added to the translation (hundreds of px).  Folding t into the chain
                // instead rounds twice at the magnitude of t: measured, half of all tracks then differ from the
                // reference in the last bit and six times as many convergence tests flip.
                const float x = mad(T1[0], o.x, T1[1] * o.y) + T1[2];
                const float y = mad(T1[3], o.x, T1[4] * o.y) + T1[5];
                const bool ok = ln.has[k] && x >= 2.0f && y >= 2.0f && x < wlim && y < hlim;
                // branch-free gather: rejected samples read pixels (0,0)..(1,1), unused; accepted ones have
                // ix + 1 <= w - 2 and iy + 1 <= h - 2, so nothing needs clamping
#if RSVIO_LK_FLOOR
                // floor_bits(v) = 2^23 + floor(v) as a float: its bit pattern is 0x4B000000 + floor(v).  The two
                // constants are folded into the level base (ibq), so the pixel address is one wide multiply-add and one
                // 64-bit add of the raw bit patterns: no F2I / I2F (XU pipe), no masks.
                const float yx = ok ? floor_bits(x) : kTwo23, yy = ok ? floor_bits(y) : kTwo23;
                const float dx = x - (yx - kTwo23), dy = y - (yy - kTwo23);
                const float ddx = 1.0f - dx, ddy = 1.0f - dy;
                RSVIO_CHECK((__float_as_uint(yx) & 0x7fffffu) + 1u < ibw && (__float_as_uint(yy) & 0x7fffffu) + 1u < (uint32_t)ib.h);
                const uint8_t *__restrict__ q0 = ibq + ((uint64_t)__float_as_uint(yy) * ibw + __float_as_uint(yx));
#else
                const uint32_t ix = ok ? __float2uint_rd(x) : 0u, iy = ok ? __float2uint_rd(y) : 0u;
                const float dx = x - (float)ix, dy = y - (float)iy;
                const float ddx = 1.0f - dx, ddy = 1.0f - dy;
                RSVIO_CHECK(ix + 1u < ibw && iy + 1u < (uint32_t)ib.h);
                const uint8_t *__restrict__ q0 = ibp + (iy * ibw + ix);
#endif
                const uint8_t *__restrict__ q1 = q0 + ibw;
                const float p00 = px_f32(q0), p10 = px_f32(q0 + 1);
                const float p01 = px_f32(q1), p11 = px_f32(q1 + 1);
                r[k] = bilin(ddx * ddy, ddx * dy, dx * ddy, dx * dy, p00, p01, p10, p11);
                okk[k] = ok;
                rsum = k == 0 ? (ok ? r[k] : 0.0f) : rsum + (ok ? r[k] : 0.0f);
                cmask[k] = __ballot_sync(0xffffffffu, ok);
                nv += __popc(cmask[k]);
            }
            rsum = warp_sum(rsum);
            if (rsum < kEps) return RSVIO_FAIL_RESIDUAL;
            // residuals exist where the sample is valid in the template and in the current image
            const int nres = __popc(cmask[0] & tmask[0]) + __popc(cmask[1] & tmask[1]);
            float i0 = 0.0f, i1 = 0.0f, i2 = 0.0f;
#if RSVIO_LK_FMA
            // n / sum once per iteration (warp-uniform).  eps <= sum <= 52 * 255: always a normal number, so the
            // hardware reciprocal with one Newton step (<= 1 ulp) needs none of __frcp_rn's range handling
            const float rscale = (float)nv * rcp_nr(rsum);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                // a sample that is invalid in the template has a zero row in H^-1 J^T (J row = 0, M finite), so only
                // the current image's validity needs a select
                const float res = okk[k] ? fmaf(r[k], rscale, -d[k]) : 0.0f;
                i0 = k == 0 ? (-M[k][0]) * res : fmaf(-M[k][0], res, i0);
                i1 = k == 0 ? (-M[k][1]) * res : fmaf(-M[k][1], res, i1);
                i2 = k == 0 ? (-M[k][2]) * res : fmaf(-M[k][2], res, i2);
            }
#else
            const float inv_rsum = __frcp_rn(rsum);
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const bool both = okk[k] && d[k] >= 0.0f;
                const float res = both ? (float)nv * r[k] * inv_rsum - d[k] : 0.0f;
                i0 = k == 0 ? (-M[k][0]) * res : i0 + (-M[k][0]) * res;
                i1 = k == 0 ? (-M[k][1]) * res : i1 + (-M[k][1]) * res;
                i2 = k == 0 ? (-M[k][2]) * res : i2 + (-M[k][2]) * res;
            }
#endif
            if (nres <= kPattern / 2) return RSVIO_FAIL_RESIDUAL;
            float inc[3], i3 = 0.0f;
            warp_sum4(i0, i1, i2, i3, lane);
            inc[0] = i0;
            inc[1] = i1;
            inc[2] = i2;
            // |inc| > 1e6 and |inc| < thr (feature_tracker.rs:369-376) on the squared norm: no square root on the
            // per-iteration path; the two forms can only disagree within an ulp of the threshold itself.  A
            // non-finite increment (feature_tracker.rs:366) makes the squared norm NaN or +inf: one compare
            // catches both exits, the (rare) failure path tells them apart.
            const float norm2 = mad(inc[2], inc[2], mad(inc[1], inc[1], inc[0] * inc[0]));
            if (!(norm2 <= 1e12f)) return finite3(inc) ? RSVIO_FAIL_DIVERGED : RSVIO_FAIL_NONFINITE;
            if (norm2 < thr2) break;  // converged: increment not applied (feature_tracker.rs:374-376)
            float E[6], N[6];
            se2_exp(inc, E);
            N[0] = mad(T1[1], E[3], T1[0] * E[0]);
            N[1] = mad(T1[1], E[4], T1[0] * E[1]);
            N[2] = mad(T1[0], E[2], T1[1] * E[5]) + T1[2];  // same association as the reference's 3x3 product
            N[3] = mad(T1[4], E[3], T1[3] * E[0]);
            N[4] = mad(T1[4], E[4], T1[3] * E[1]);
            N[5] = mad(T1[3], E[2], T1[4] * E[5]) + T1[5];
#pragma unroll
            for (int k = 0; k < 6; ++k) T1[k] = N[k];
            // inbound(.., 2) (feature_tracker.rs:380-388) in the float domain: round-half-away(x) >= 2 <=> x >= 1.5 and
            // round-half-away(x) <= w - 3 <=> x < w - 2.5 (exact for every float, NaN fails both forms)
            if (!(T1[2] >= 1.5f && T1[5] >= 1.5f && T1[2] < wlim - 0.5f && T1[5] < hlim - 0.5f)) return RSVIO_FAIL_OOB;
        }
        T1[2] = T1[2] * s;
        T1[5] = T1[5] * s;
    }
    // rotation block of T0 * T1 (feature_tracker.rs:335-340)
    const float r00 = T0[0] * T1[0] + T0[1] * T1[3];
    const float r01 = T0[0] * T1[1] + T0[1] * T1[4];
    const float r10 = T0[3] * T1[0] + T0[4] * T1[3];
    const float r11 = T0[3] * T1[1] + T0[4] * T1[4];
    T1[0] = r00; T1[1] = r01; T1[3] = r10; T1[4] = r11;
    return RSVIO_TRACK_OK;
}

#ifndef RSVIO_LK_MINB
#define RSVIO_LK_MINB 8  // resident CTAs per SM the register budget is sized for (A/B: -DRSVIO_LK_MINB=10)
#endif
__global__ void __launch_bounds__(kWarpsPerBlock * 32, RSVIO_LK_MINB)
lk_track_fb_kernel(const __grid_constant__ PyrDesc pd, const __grid_constant__ LkBatch b) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    __shared__ float2 s_pat[64];
    if (threadIdx.x < 64) {
        float ox, oy;
        pattern_offset(threadIdx.x < kPattern ? (int)threadIdx.x : 0, ox, oy);
        s_pat[threadIdx.x] = make_float2(ox, oy);
    }
    __syncthreads();
    Lane ln;
    ln.pat = (uint32_t)__cvta_generic_to_shared(s_pat + lane);
    asm volatile("mov.u32 %0, %0;" : "+r"(ln.pat));  // opaque: kept in a register, not rematerialised from the thread id
    ln.has[0] = true;
    ln.has[1] = lane + 32 < kPattern;

    // Work distribution: the live points of all groups form one compacted list (exclusive prefix of the
    // per-group counts in shared memory); warps fetch the next point from a global counter, so a warp that
    // drew long tracks (many iterations) does not hold up the launch while others idle.
    extern __shared__ int prefix[];  // groups + 1 entries when the launch schedules dynamically (sized at launch:
                                     // a small table leaves the shared memory to kernels running beside this one)
    __shared__ int warp_tot[kWarpsPerBlock];
    const bool dynamic = b.work_counter != nullptr && b.groups <= kMaxGroupsSmem;
    // two counters used alternately: this launch draws from [work_parity] and clears the other one for the next
    unsigned *const wc = dynamic ? b.work_counter + b.work_parity : nullptr;
    if (dynamic && blockIdx.x == 0 && threadIdx.x == 0) b.work_counter[b.work_parity ^ 1] = 0u;
    if (dynamic) {
        int base = 0;
        for (int g0 = 0; g0 < b.groups; g0 += blockDim.x) {
            const int g = g0 + threadIdx.x;
            int c = g < b.groups ? min(__ldg(b.counts + g), b.cap) : 0;
            if (b.first && g < b.groups) c = max(c - __ldg(b.first + g), 0);  // rows [first, count)
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) warp_tot[threadIdx.x >> 5] = incl;
            __syncthreads();
            int off = base;
            for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) off += warp_tot[w];
            if (g < b.groups) prefix[g] = off + incl - c;
            int tot = 0;
            for (int w = 0; w < kWarpsPerBlock; ++w) tot += warp_tot[w];
            base += tot;
            __syncthreads();
        }
        if (threadIdx.x == 0) prefix[b.groups] = base;
        __syncthreads();
    }
    unsigned n_iters = 0, n_templates = 0, n_tracks = 0;
    const int total = dynamic ? prefix[b.groups] : b.groups * b.cap;
    for (int it0 = warp;; it0 += n_warps) {
        int g, k, item;
        if (dynamic) {
            int w = 0;
            if (lane == 0) w = (int)atomicAdd(wc, 1u);
            w = __shfl_sync(0xffffffffu, w, 0);
            if (w >= total) break;
            int lo = 0, hi = b.groups;  // largest g with prefix[g] <= w
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (prefix[mid] <= w) lo = mid; else hi = mid;
            }
            g = lo;
            k = w - prefix[g] + (b.first ? __ldg(b.first + g) : 0);
        } else {
            if (it0 >= total) break;
            g = it0 / b.cap;
            k = it0 - g * b.cap;
            if (k >= __ldg(b.counts + g) || (b.first && k < __ldg(b.first + g))) continue;
        }
        item = g * b.cap + k;
        const uint8_t *A = b.pyr_a + (size_t)g * pd.pyr_stride;
        const uint8_t *B = b.pyr_b + (size_t)g * pd.pyr_stride;
        float T0[6], Tf[6], Tb[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) T0[q] = __ldg(b.in_T6 + (size_t)item * 6 + q);
        unsigned it_f = 0, it_b = 0;
        float Tin[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) Tin[q] = T0[q];
        int st = RSVIO_TRACK_OK;
        // forward (A -> B) then backward (B -> A) through ONE copy of the level loop: the fused kernel
        // stays within the instruction cache
#pragma unroll 1
        for (int dir = 0; dir < 2; ++dir) {
            unsigned iters = 0;
            st = track_one_point(pd, dir == 0 ? A : B, dir == 0 ? B : A, Tin, Tb, b.max_it, b.thr, b.floor_bias, ln, lane, iters,
                                 n_templates);
            if (dir == 0) it_f = iters; else it_b = iters;
            if (dir == 0) {
#pragma unroll
                for (int q = 0; q < 6; ++q) Tf[q] = Tin[q] = Tb[q];
                if (st != RSVIO_TRACK_OK) break;
            } else if (st != RSVIO_TRACK_OK) {
                st += RSVIO_FAIL_BACKWARD;
            } else {  // forward-backward consistency (feature_tracker.rs:277-281)
                const float dx = T0[2] - Tb[2], dy = T0[5] - Tb[5];
                if (!(dx * dx + dy * dy < 0.4f)) st = RSVIO_FAIL_FB;
            }
        }
        n_iters += it_f + it_b;
        ++n_tracks;
        if (lane < 6) b.out_T6[(size_t)item * 6 + lane] = Tf[lane];
        if (lane == 0) {
            b.status[item] = st;
            if (b.iters) {
                b.iters[(size_t)item * 2] = (int)it_f;
                b.iters[(size_t)item * 2 + 1] = (int)it_b;
            }
        }
    }
    if (b.counters && lane == 0 && n_tracks) {
        atomicAdd(b.counters + CNT_TRACKS, (unsigned long long)n_tracks);
        atomicAdd(b.counters + CNT_ITERS, (unsigned long long)n_iters);
        atomicAdd(b.counters + CNT_TEMPLATES, (unsigned long long)n_templates);
    }
}

__global__ void se2_exp_kernel(const float *__restrict__ tw, int n, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a[3] = {tw[3 * i], tw[3 * i + 1], tw[3 * i + 2]};
    float m[6];
    se2_exp(a, m);
    float *o = out + (size_t)9 * i;
    o[0] = m[0]; o[1] = m[1]; o[2] = m[2];
    o[3] = m[3]; o[4] = m[4]; o[5] = m[5];
    o[6] = 0.0f; o[7] = 0.0f; o[8] = 1.0f;
}

}  // namespace

cudaError_t lk_track(const PyrDesc &pd, const LkBatch &b, cudaStream_t st) {
    const long long total = (long long)b.groups * b.cap;
    if (total <= 0) return cudaSuccess;
    long long blocks = (total + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const bool dynamic = b.work_counter != nullptr && b.groups <= kMaxGroupsSmem;
    // dynamic scheduling: one resident wave (64 registers -> 8 CTAs per SM); static striding: up to 16 per SM
    // RSVIO_LK_BLOCKS_PER_SM (experiments): fewer persistent CTAs than the register file holds leave room for the
    // kernels of another handle to run beside this launch
    const char *env = getenv("RSVIO_LK_BLOCKS_PER_SM");
    const int env_v = env ? atoi(env) : 0;
    const int resident = (env_v >= 1 && env_v <= RSVIO_LK_MINB) ? env_v : RSVIO_LK_MINB;
    const long long max_blocks = (long long)sm_count() * (dynamic ? (b.blocks_per_sm >= 1 && b.blocks_per_sm < resident ? b.blocks_per_sm : resident) : 16);
    if (blocks > max_blocks) blocks = max_blocks;
    const size_t smem = dynamic ? sizeof(int) * (size_t)(b.groups + 1) : 0;
    LkBatch bb = b;
    bb.floor_bias = 0x4B000000u;
    lk_track_fb_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, smem, st>>>(pd, bb);
    return cudaGetLastError();
}

cudaError_t se2_exp_batch(const float *d_twists, int n, float *d_out9, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    se2_exp_kernel<<<(n + 127) / 128, 128, 0, st>>>(d_twists, n, d_out9);
    return cudaGetLastError();
}

}  // namespace rsvio

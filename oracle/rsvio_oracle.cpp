// rsvio_oracle.cpp -- CPU parity oracle for RS-VIO's stereo patch tracker (tracker A).
//
// TEST INFRASTRUCTURE ONLY (see rsvio_oracle.h).  PARITY UNPINNED: the reference has
// no golden vectors for this path and cannot be built here; see the header comment.
//
// Every function cites the reference file:line it restates (paths relative to the
// reference repo root).  Build with -O2 -ffp-contract=off: Rust never contracts
// a*b+c into an FMA, and the reference computes everything on this path in f32.
// `.round()` in Rust is half-away-from-zero == roundf(); float -> u32 `as` casts
// saturate (negative / NaN -> 0).

#include "rsvio_oracle.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <vector>

#include <atomic>
#include <thread>

namespace {

constexpr int NP = ORC_PATTERN_SIZE;
constexpr float F32_EPS = std::numeric_limits<float>::epsilon();
constexpr float BIG = 3.4e38f;

// Dynamic-chunk parallel for over [0, n): stands in for rayon's into_par_iter / par_iter
// (feature_tracker.rs:213, :260).  Results do not depend on the thread count.
template <typename F>
void parallel_for(int n, int chunk, int nthreads, F f) {
    if (nthreads <= 1 || n <= chunk) {
        for (int i = 0; i < n; ++i) f(i);
        return;
    }
    std::atomic<int> next(0);
    auto worker = [&]() {
        for (;;) {
            const int b = next.fetch_add(chunk);
            if (b >= n) return;
            const int e = std::min(n, b + chunk);
            for (int i = b; i < e; ++i) f(i);
        }
    };
    const int nt = std::min(nthreads, (n + chunk - 1) / chunk);
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    for (int t = 1; t < nt; ++t) th.emplace_back(worker);
    worker();
    for (auto &t : th) t.join();
}

// Pattern52::PATTERN_RAW -- patch.rs:19-72.  Odd-integer lattice inside a radius-7.6
// disc: rows y = 7,5,3,1,-1,-3,-5,-7 with |x| <= 3,5,7,7,7,7,5,3, x ascending.
struct PatternTable {
    float raw[NP][2];
    PatternTable() {
        static const int half_w[8] = {3, 5, 7, 7, 7, 7, 5, 3};
        int k = 0;
        for (int r = 0; r < 8; ++r) {
            int y = 7 - 2 * r;
            for (int x = -half_w[r]; x <= half_w[r]; x += 2) {
                raw[k][0] = (float)x;
                raw[k][1] = (float)y;
                ++k;
            }
        }
    }
};
const PatternTable PATTERN;

inline uint32_t sat_u32(float v) {  // Rust `f32 as u32`
    if (!(v > 0.0f)) return 0u;     // NaN, negatives, -0
    if (v >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)v;
}

struct Level {
    const uint8_t *p;
    int w, h;
};

// ---------------------------------------------------------------------------------
// image::imageops::resize(.., FilterType::Triangle) for Luma<u8>  (image 0.25, sample.rs:
// `resize` -> `vertical_sample` then `horizontal_sample`; not vendored in the reference;
// call sites feature_tracker.rs:210,217).  SURVEY.md Appendix A.
// ---------------------------------------------------------------------------------
struct Taps {
    std::vector<int> left, cnt, off;
    std::vector<float> w;
};

inline float triangle_kernel(float x) {
    float a = std::fabs(x);
    return a < 1.0f ? 1.0f - a : 0.0f;
}

Taps make_taps(int in_dim, int out_dim) {
    Taps t;
    t.left.resize(out_dim);
    t.cnt.resize(out_dim);
    t.off.resize(out_dim);
    const float ratio = (float)in_dim / (float)out_dim;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float support = 1.0f * sratio;
    for (int o = 0; o < out_dim; ++o) {
        float c = ((float)o + 0.5f) * ratio;
        long long l = (long long)std::floor(c - support);
        l = l < 0 ? 0 : (l > (long long)in_dim - 1 ? (long long)in_dim - 1 : l);
        long long r = (long long)std::ceil(c + support);
        r = r < l + 1 ? l + 1 : (r > (long long)in_dim ? (long long)in_dim : r);
        c = c - 0.5f;
        t.left[o] = (int)l;
        t.cnt[o] = (int)(r - l);
        t.off[o] = (int)t.w.size();
        float sum = 0.0f;
        for (long long i = l; i < r; ++i) {
            float w = triangle_kernel(((float)i - c) / sratio);
            t.w.push_back(w);
            sum += w;
        }
        for (int k = 0; k < t.cnt[o]; ++k) t.w[t.off[o] + k] /= sum;
    }
    return t;
}

void resize_triangle_u8(const uint8_t *src, int w, int h, uint8_t *dst, int nw, int nh) {
    if (w == 0 || h == 0) {
        std::memset(dst, 0, (size_t)nw * nh);
        return;
    }
    if (nw == w && nh == h) {  // same size: plain copy
        std::memcpy(dst, src, (size_t)w * h);
        return;
    }
    // vertical pass first: f32 image w x nh, neither clamped nor rounded
    Taps ty = make_taps(h, nh);
    std::vector<float> tmp((size_t)w * nh);
    for (int oy = 0; oy < nh; ++oy) {
        const float *wy = &ty.w[ty.off[oy]];
        const int top = ty.left[oy], n = ty.cnt[oy];
        float *out = &tmp[(size_t)oy * w];
        for (int x = 0; x < w; ++x) {
            float t = 0.0f;
            for (int k = 0; k < n; ++k) t += (float)src[(size_t)(top + k) * w + x] * wy[k];
            out[x] = t;
        }
    }
    // horizontal pass, clamp to [0,255], round half away from zero
    Taps tx = make_taps(w, nw);
    for (int ox = 0; ox < nw; ++ox) {
        const float *wx = &tx.w[tx.off[ox]];
        const int left = tx.left[ox], n = tx.cnt[ox];
        for (int y = 0; y < nh; ++y) {
            const float *row = &tmp[(size_t)y * w + left];
            float t = 0.0f;
            for (int k = 0; k < n; ++k) t += row[k] * wx[k];
            t = t < 0.0f ? 0.0f : (t > 255.0f ? 255.0f : t);
            dst[(size_t)y * nw + ox] = (uint8_t)std::round(t);
        }
    }
}

std::vector<Level> split_pyramid(const uint8_t *buf, int w0, int h0, int levels) {
    std::vector<Level> v(levels);
    size_t off = 0;
    for (int i = 0; i < levels; ++i) {
        v[i].p = buf + off;
        v[i].w = w0 >> i;
        v[i].h = h0 >> i;
        off += (size_t)v[i].w * v[i].h;
    }
    return v;
}

// ---------------------------------------------------------------------------------
// image_utilities.rs
// ---------------------------------------------------------------------------------

// inbound -- image_utilities.rs:75-80
inline bool inbound(int w, int h, float x, float y, uint32_t radius) {
    uint32_t rx = sat_u32(std::round(x));
    uint32_t ry = sat_u32(std::round(y));
    return rx >= radius && ry >= radius && rx < (uint32_t)w - radius && ry < (uint32_t)h - radius;
}

// distance of (x,y) to the nearest decision boundary of inbound(.., radius)
inline float inbound_margin(int w, int h, float x, float y, uint32_t radius) {
    float lo = (float)radius - 0.5f;
    float hx = (float)((uint32_t)w - radius) - 0.5f, hy = (float)((uint32_t)h - radius) - 0.5f;
    float m = std::fabs(x - lo);
    m = std::min(m, std::fabs(y - lo));
    m = std::min(m, std::fabs(x - hx));
    m = std::min(m, std::fabs(y - hy));
    return m;
}

// image_grad -- image_utilities.rs:5-66.  Caller guarantees the 4x4 neighbourhood exists.
inline void image_grad(const Level &im, float x, float y, float out[3]) {
    const uint32_t ix = sat_u32(std::floor(x)), iy = sat_u32(std::floor(y));
    const float dx = x - (float)ix, dy = y - (float)iy;
    const float ddx = 1.0f - dx, ddy = 1.0f - dy;
    const uint8_t *p = im.p;
    const size_t W = (size_t)im.w;
    auto px = [&](int ox, int oy) { return (float)p[(size_t)(iy + oy) * W + (size_t)(ix + ox)]; };
    const float p00 = px(0, 0), p10 = px(1, 0), p01 = px(0, 1), p11 = px(1, 1);
    const float res0 = ddx * ddy * p00 + ddx * dy * p01 + dx * ddy * p10 + dx * dy * p11;
    const float pm0 = px(-1, 0), pm1 = px(-1, 1);
    const float res_mx = ddx * ddy * pm0 + ddx * dy * pm1 + dx * ddy * p00 + dx * dy * p01;
    const float p20 = px(2, 0), p21 = px(2, 1);
    const float res_px = ddx * ddy * p10 + ddx * dy * p11 + dx * ddy * p20 + dx * dy * p21;
    const float res1 = 0.5f * (res_px - res_mx);
    const float p0m = px(0, -1), p1m = px(1, -1);
    const float res_my = ddx * ddy * p0m + ddx * dy * p00 + dx * ddy * p1m + dx * dy * p10;
    const float p02 = px(0, 2), p12 = px(1, 2);
    const float res_py = ddx * ddy * p01 + ddx * dy * p02 + dx * ddy * p11 + dx * dy * p12;
    const float res2 = 0.5f * (res_py - res_my);
    out[0] = res0;
    out[1] = res1;
    out[2] = res2;
}

// se2_exp_matrix -- image_utilities.rs:82-106.  Rotation2::new(t) = [[cos,-sin],[sin,cos]].
// Returns the 6 free entries [m11,m12,m13, m21,m22,m23]; third row is (0,0,1).
inline void se2_exp(const float a[3], float m[6]) {
    const float theta = a[2];
    const float sn = std::sin(theta), cs = std::cos(theta);
    float s_by_t, omc_by_t;
    if (std::fabs(theta) < F32_EPS) {
        const float t2 = theta * theta;
        s_by_t = 1.0f - 1.0f / 6.0f * t2;
        omc_by_t = 0.5f * theta - 1.0f / 24.0f * theta * t2;
    } else {
        s_by_t = sn / theta;
        omc_by_t = (1.0f - cs) / theta;
    }
    m[0] = cs;
    m[1] = -sn;
    m[3] = sn;
    m[4] = cs;
    m[2] = s_by_t * a[0] - omc_by_t * a[1];
    m[5] = omc_by_t * a[0] + s_by_t * a[1];
}

// ---------------------------------------------------------------------------------
// patch.rs
// ---------------------------------------------------------------------------------
struct Margins {
    float conv = BIG, geom = BIG, fb = BIG;
};

// nalgebra 0.33 Cholesky::new for a 3x3 (column-by-column, lower triangle), then
// solve_mut(I): forward substitution (column axpy form) and adjoint back substitution
// (dot form).  patch.rs:143-147.  Returns false when a pivot is zero, negative or NaN.
bool chol3_inverse(const float H[3][3], float Hinv[3][3]) {
    float m[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m[i][j] = H[i][j];
    for (int j = 0; j < 3; ++j) {
        for (int k = 0; k < j; ++k) {
            const float factor = -m[j][k];
            for (int i = j; i < 3; ++i) m[i][j] = factor * m[i][k] + m[i][j];
        }
        const float diag = m[j][j];
        if (diag == 0.0f) return false;
        if (!(diag >= 0.0f)) return false;  // try_sqrt: None for negatives and NaN
        const float denom = std::sqrt(diag);
        m[j][j] = denom;
        for (int i = j + 1; i < 3; ++i) m[i][j] /= denom;
    }
    for (int c = 0; c < 3; ++c) {
        float b[3] = {c == 0 ? 1.0f : 0.0f, c == 1 ? 1.0f : 0.0f, c == 2 ? 1.0f : 0.0f};
        for (int i = 0; i < 3; ++i) {  // L y = b
            const float coeff = b[i] / m[i][i];
            b[i] = coeff;
            for (int r = i + 1; r < 3; ++r) b[r] = (-coeff) * m[r][i] + b[r];
        }
        for (int i = 2; i >= 0; --i) {  // L^T x = y
            float dot = 0.0f;
            if (i == 1) dot = m[2][1] * b[2];
            if (i == 0) dot = m[1][0] * b[1] + m[2][0] * b[2];
            b[i] = (b[i] - dot) / m[i][i];
        }
        for (int i = 0; i < 3; ++i) Hinv[i][c] = b[i];
    }
    return true;
}

// Pattern52::new + set_data_jac_se2 -- patch.rs:75-162
void pattern52_new(const Level &im, float px, float py, orc_pattern52 *out, Margins *mg) {
    float J[NP][3];
    std::memset(J, 0, sizeof(J));
    std::memset(out, 0, sizeof(*out));
    out->valid = 0;
    out->mean = 1.0f;
    int n_valid = 0;
    float sum = 0.0f;
    float gsum[3] = {0.0f, 0.0f, 0.0f};
    for (int i = 0; i < NP; ++i) {
        const float ox = PATTERN.raw[i][0] / 2.0f, oy = PATTERN.raw[i][1] / 2.0f;
        const float x = px + ox, y = py + oy;
        const float jw02 = -PATTERN.raw[i][1] / 2.0f, jw12 = PATTERN.raw[i][0] / 2.0f;
        if (mg) mg->geom = std::min(mg->geom, inbound_margin(im.w, im.h, x, y, 2));
        if (inbound(im.w, im.h, x, y, 2)) {
            float vg[3];
            image_grad(im, x, y, vg);
            out->data[i] = vg[0];
            sum += vg[0];
            // (1x2) * (2x3): column j = gx*jw[0][j] then += gy*jw[1][j]
            J[i][0] = vg[1] * 1.0f + vg[2] * 0.0f;
            J[i][1] = vg[1] * 0.0f + vg[2] * 1.0f;
            J[i][2] = vg[1] * jw02 + vg[2] * jw12;
            for (int k = 0; k < 3; ++k) gsum[k] += J[i][k];
            ++n_valid;
        } else {
            out->data[i] = -1.0f;
        }
    }
    out->mean = sum / (float)n_valid;
    const float mean_inv = (float)n_valid / sum;
    for (int i = 0; i < NP; ++i) {
        if (out->data[i] >= 0.0f) {
            for (int k = 0; k < 3; ++k) {
                const float rhs = gsum[k] * out->data[i] / sum;
                J[i][k] += -rhs;
            }
            out->data[i] *= mean_inv;
        } else {
            J[i][0] = J[i][1] = J[i][2] = 0.0f;
        }
    }
    for (int i = 0; i < NP; ++i)
        for (int k = 0; k < 3; ++k) J[i][k] *= mean_inv;

    // H = J^T J: nalgebra gemm = per result column a sequence of column axpys over i
    float H[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            float acc = J[0][a] * J[0][b];
            for (int i = 1; i < NP; ++i) acc = acc + J[i][a] * J[i][b];
            H[a][b] = acc;
        }
    float Hinv[3][3];
    if (!chol3_inverse(H, Hinv)) return;  // valid stays false, hinv_jt stays zero
    bool finite = true;
    for (int i = 0; i < NP; ++i)
        for (int k = 0; k < 3; ++k) {
            const float v = Hinv[k][0] * J[i][0] + Hinv[k][1] * J[i][1] + Hinv[k][2] * J[i][2];
            out->hinv_jt[k * NP + i] = v;
            finite = finite && std::isfinite(v);
        }
    for (int i = 0; i < NP; ++i) finite = finite && std::isfinite(out->data[i]);
    out->valid = (out->mean > F32_EPS && finite) ? 1 : 0;
}

// Pattern52::residual -- patch.rs:163-232
bool pattern52_residual(const orc_pattern52 &pat, const Level &im, const float *xs,
                        const float *ys, float *res, Margins *mg) {
    float sum = 0.0f;
    int n_valid = 0;
    const float wlim = (float)((uint32_t)im.w - 2u), hlim = (float)((uint32_t)im.h - 2u);
    for (int i = 0; i < NP; ++i) {
        const float x = xs[i], y = ys[i];
        if (mg) {
            float m = std::min(std::min(std::fabs(x - 2.0f), std::fabs(y - 2.0f)),
                               std::min(std::fabs(x - wlim), std::fabs(y - hlim)));
            mg->geom = std::min(mg->geom, m);
        }
        if (x >= 2.0f && y >= 2.0f && x < wlim && y < hlim) {
            const uint32_t ix = sat_u32(std::floor(x)), iy = sat_u32(std::floor(y));
            const float dx = x - (float)ix, dy = y - (float)iy;
            const float ddx = 1.0f - dx, ddy = 1.0f - dy;
            const uint8_t *p = im.p + (size_t)iy * im.w + ix;
            const float p00 = (float)p[0], p10 = (float)p[1];
            const float p01 = (float)p[im.w], p11 = (float)p[im.w + 1];
            res[i] = ddx * ddy * p00 + ddx * dy * p01 + dx * ddy * p10 + dx * dy * p11;
            sum += res[i];
            ++n_valid;
        } else {
            res[i] = -1.0f;
        }
    }
    if (sum < F32_EPS) return false;
    int n_res = 0;
    for (int i = 0; i < NP; ++i) {
        if (res[i] >= 0.0f && pat.data[i] >= 0.0f) {
            const float val = res[i];
            res[i] = (float)n_valid * val / sum - pat.data[i];
            ++n_res;
        } else {
            res[i] = 0.0f;
        }
    }
    return n_res > NP / 2;
}

// ---------------------------------------------------------------------------------
// feature_tracker.rs
// ---------------------------------------------------------------------------------

// track_point_at_level -- feature_tracker.rs:344-395.  T = [m11,m12,m13,m21,m22,m23].
int track_point_at_level(const Level &im, const orc_pattern52 &pat, float T[6], int max_it,
                         float thr, int *iters, Margins *mg) {
    float xs[NP], ys[NP], res[NP];
    for (int it = 0; it < max_it; ++it) {
        for (int i = 0; i < NP; ++i) {  // R * pattern + t   (:356-360)
            const float ox = PATTERN.raw[i][0] / 2.0f, oy = PATTERN.raw[i][1] / 2.0f;
            xs[i] = (T[0] * ox + T[1] * oy) + T[2];
            ys[i] = (T[3] * ox + T[4] * oy) + T[5];
        }
        if (iters) ++*iters;
        if (!pattern52_residual(pat, im, xs, ys, res, mg)) return ORC_FAIL_RESIDUAL;
        float inc[3];
        for (int k = 0; k < 3; ++k) {  // inc = (-H^-1 J^T) * res, sequential over samples
            const float *row = &pat.hinv_jt[k * NP];
            float acc = (-row[0]) * res[0];
            for (int i = 1; i < NP; ++i) acc = acc + (-row[i]) * res[i];
            inc[k] = acc;
        }
        if (!(std::isfinite(inc[0]) && std::isfinite(inc[1]) && std::isfinite(inc[2])))
            return ORC_FAIL_NONFINITE;
        const float norm = std::sqrt(inc[0] * inc[0] + inc[1] * inc[1] + inc[2] * inc[2]);
        if (norm > 1e6f) return ORC_FAIL_DIVERGED;
        if (mg) mg->conv = std::min(mg->conv, std::fabs(norm - thr));
        if (norm < thr) break;  // converged: the increment is NOT applied (:374-376)
        float E[6];
        se2_exp(inc, E);
        // T <- T * exp(inc): 3x3 gemm, k = 0,1,2 in order; third rows are (0,0,1)
        float N[6];
        N[0] = T[0] * E[0] + T[1] * E[3];
        N[1] = T[0] * E[1] + T[1] * E[4];
        N[2] = T[0] * E[2] + T[1] * E[5] + T[2];
        N[3] = T[3] * E[0] + T[4] * E[3];
        N[4] = T[3] * E[1] + T[4] * E[4];
        N[5] = T[3] * E[2] + T[4] * E[5] + T[5];
        for (int k = 0; k < 6; ++k) T[k] = N[k];
        if (mg) mg->geom = std::min(mg->geom, inbound_margin(im.w, im.h, T[2], T[5], 2));
        if (!inbound(im.w, im.h, T[2], T[5], 2)) return ORC_FAIL_OOB;
    }
    return ORC_TRACK_OK;
}

// track_one_point -- feature_tracker.rs:292-342
int track_one_point(const std::vector<Level> &pa, const std::vector<Level> &pb, const float T0[6],
                    int max_it, float thr, float T1[6], int *iters, Margins *mg) {
    const int levels = (int)pa.size();
    T1[0] = 1.0f; T1[1] = 0.0f; T1[2] = T0[2];
    T1[3] = 0.0f; T1[4] = 1.0f; T1[5] = T0[5];
    for (int i = levels - 1; i >= 0; --i) {
        const float s = (float)(1 << i);
        T1[2] /= s;
        T1[5] /= s;
        orc_pattern52 pat;
        pattern52_new(pa[i], T0[2] / s, T0[5] / s, &pat, mg);
        if (!pat.valid) return ORC_FAIL_TEMPLATE;
        int st = track_point_at_level(pb[i], pat, T1, max_it, thr, iters, mg);
        if (st != ORC_TRACK_OK) return st;
        T1[2] *= s;
        T1[5] *= s;
    }
    // rotation block of T0 * T1 (:335-340); translation stays T1's
    const float r00 = T0[0] * T1[0] + T0[1] * T1[3];
    const float r01 = T0[0] * T1[1] + T0[1] * T1[4];
    const float r10 = T0[3] * T1[0] + T0[4] * T1[3];
    const float r11 = T0[3] * T1[1] + T0[4] * T1[4];
    T1[0] = r00; T1[1] = r01; T1[3] = r10; T1[4] = r11;
    return ORC_TRACK_OK;
}

// one element of track_points -- feature_tracker.rs:259-287
int track_fb(const std::vector<Level> &pa, const std::vector<Level> &pb, const float T0[6],
             int max_it, float thr, float Tout[6], int iters[2], Margins *mg) {
    float Tf[6], Tb[6];
    iters[0] = iters[1] = 0;
    int st = track_one_point(pa, pb, T0, max_it, thr, Tf, &iters[0], mg);
    for (int k = 0; k < 6; ++k) Tout[k] = Tf[k];
    if (st != ORC_TRACK_OK) return st;
    st = track_one_point(pb, pa, Tf, max_it, thr, Tb, &iters[1], mg);
    if (st != ORC_TRACK_OK) return st + ORC_FAIL_BACKWARD;
    const float dx = T0[2] - Tb[2], dy = T0[5] - Tb[5];
    const float d2 = dx * dx + dy * dy;
    if (mg) mg->fb = std::fabs(d2 - 0.4f);
    return d2 < 0.4f ? ORC_TRACK_OK : ORC_FAIL_FB;
}

// ---------------------------------------------------------------------------------
// imageproc::corners (0.25; not vendored; call site image_utilities.rs:156).  Appendix B.
// ---------------------------------------------------------------------------------
static const int CIRCLE[16][2] = {{0, -3}, {1, -3}, {2, -2}, {3, -1}, {3, 0},  {3, 1},  {2, 2},   {1, 3},
                                  {0, 3},  {-1, 3}, {-2, 2}, {-3, 1}, {-3, 0}, {-3, -1}, {-2, -2}, {-1, -3}};

template <typename F>
bool search_span(const int16_t c[16], int length, F f) {
    int nb_ok = 0, nb_ok_start = -1;
    for (int i = 0; i < 16; ++i) {
        if (f(c[i])) {
            if (++nb_ok == length) return true;
        } else {
            if (nb_ok_start < 0) nb_ok_start = nb_ok;
            nb_ok = 0;
        }
    }
    return nb_ok + nb_ok_start >= length;
}

bool is_corner_fast9(const uint8_t *img, int w, int h, int threshold, int x, int y) {
    if (x < 3 || y < 3 || w <= x + 3 || h <= y + 3) return false;
    const int16_t c = img[(size_t)y * w + x];
    const int16_t lo = (int16_t)(c - threshold), hi = (int16_t)(c + threshold);
    int16_t px[16];
    for (int k = 0; k < 16; ++k) px[k] = img[(size_t)(y + CIRCLE[k][1]) * w + (x + CIRCLE[k][0])];
    const int16_t p0 = px[0], p4 = px[4], p8 = px[8], p12 = px[12];
    const bool above = (p0 > hi && p4 > hi) || (p4 > hi && p8 > hi) || (p8 > hi && p12 > hi) || (p12 > hi && p0 > hi);
    const bool below = (p0 < lo && p4 < lo) || (p4 < lo && p8 < lo) || (p8 < lo && p12 < lo) || (p12 < lo && p0 < lo);
    if (!above && !below) return false;
    return (above && search_span(px, 9, [hi](int16_t v) { return v > hi; })) ||
           (below && search_span(px, 9, [lo](int16_t v) { return v < lo; }));
}

int fast_corner_score(const uint8_t *img, int w, int h, int threshold, int x, int y) {
    int mx = 255, mn = threshold;
    for (;;) {
        if (mx == mn) return mx;
        const int mean = (mx + mn) / 2;
        const int probe = (mx == mn + 1) ? mx : mean;
        if (is_corner_fast9(img, w, h, probe, x, y))
            mn = probe;
        else
            mx = probe - 1;
    }
}

struct Corner {
    uint32_t x, y;
    float score;
};

void corners_fast9(const uint8_t *img, int w, int h, int threshold, std::vector<Corner> &out) {
    out.clear();
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
            if (is_corner_fast9(img, w, h, threshold, x, y))
                out.push_back({(uint32_t)x, (uint32_t)y, (float)fast_corner_score(img, w, h, threshold, x, y)});
}

// point_in_bound -- image_utilities.rs:68-73
inline bool point_in_bound(const Corner &k, uint32_t h, uint32_t w, uint32_t r) {
    return k.x >= r && k.x <= w - r && k.y >= r && k.y <= h - r;
}

// detect_key_points -- image_utilities.rs:108-175 (num_points_in_cell fixed to 1, feature_tracker.rs:227)
void detect_key_points(const uint8_t *img, uint32_t w, uint32_t h, uint32_t g, const uint32_t *cur, int n_cur,
                       std::vector<Corner> &all) {
    const uint32_t EDGE = 19;
    all.clear();
    if (g == 0 || w < g || h < g) return;
    const uint32_t rows = h / g + 1, cols = w / g + 1;
    std::vector<int32_t> grid((size_t)rows * cols, 0);
    const uint32_t x_start = (w % g) / 2, x_stop = x_start + g * (w / g - 1) + 1;
    const uint32_t y_start = (h % g) / 2, y_stop = y_start + g * (h / g - 1) + 1;
    for (int i = 0; i < n_cur; ++i) {
        const uint32_t cx = cur[2 * i], cy = cur[2 * i + 1];
        if (cx >= x_start && cy >= y_start && cx < x_stop + g && cy < y_stop + g)
            grid[(size_t)((cy - y_start) / g) * cols + (cx - x_start) / g] += 1;
    }
    std::vector<uint8_t> cell((size_t)g * g);
    std::vector<Corner> fc;
    for (uint32_t x = x_start; x < x_stop; x += g)
        for (uint32_t y = y_start; y < y_stop; y += g) {
            if (grid[(size_t)((y - y_start) / g) * cols + (x - x_start) / g] > 0) continue;
            for (uint32_t r = 0; r < g; ++r) std::memcpy(&cell[(size_t)r * g], img + (size_t)(y + r) * w + x, g);
            int added = 0;
            int threshold = 40;
            while (added < 1 && threshold >= 10) {
                corners_fast9(cell.data(), (int)g, (int)g, threshold, fc);
                std::stable_sort(fc.begin(), fc.end(), [](const Corner &a, const Corner &b) { return a.score < b.score; });
                for (Corner pt : fc) {
                    if (added >= 1) break;
                    pt.x += x;
                    pt.y += y;
                    if (point_in_bound(pt, h, w, EDGE)) {
                        all.push_back(pt);
                        ++added;
                    }
                }
                threshold -= 5;
            }
        }
}

void build_pyramid(const uint8_t *src, int w0, int h0, int levels, uint8_t *dst, int nthreads) {
    std::vector<size_t> off(levels);
    size_t o = 0;
    for (int i = 0; i < levels; ++i) {
        off[i] = o;
        o += (size_t)(w0 >> i) * (h0 >> i);
    }
    parallel_for(levels, 1, nthreads,
                 [&](int i) { resize_triangle_u8(src, w0, h0, dst + off[i], w0 >> i, h0 >> i); });
}

void track_points(const uint8_t *pyr_a, const uint8_t *pyr_b, int w0, int h0, int levels, int n,
                  const float *in_T, int max_it, float thr, float *out_T, int32_t *status,
                  int32_t *iters, float *margins, int nthreads) {
    const std::vector<Level> pa = split_pyramid(pyr_a, w0, h0, levels);
    const std::vector<Level> pb = split_pyramid(pyr_b, w0, h0, levels);
    parallel_for(n, 8, nthreads, [&](int i) {
        Margins mg;
        int it[2];
        float T[6];
        const int st = track_fb(pa, pb, in_T + 6 * i, max_it, thr, T, it, margins ? &mg : nullptr);
        for (int k = 0; k < 6; ++k) out_T[6 * i + k] = T[k];
        status[i] = st;
        if (iters) {
            iters[2 * i] = it[0];
            iters[2 * i + 1] = it[1];
        }
        if (margins) {
            margins[3 * i] = mg.conv;
            margins[3 * i + 1] = mg.geom;
            margins[3 * i + 2] = mg.fb;
        }
    });
}

}  // namespace

// ---------------------------------------------------------------------------------
// Stateful tracker -- feature_tracker.rs:36-207
// ---------------------------------------------------------------------------------
struct orc_tracker {
    int levels;
    uint32_t grid;
    int max_it;
    float thr;
    int nthreads;
    uint64_t last_id = 0;
    int w = 0, h = 0;
    bool have_prev = false;
    std::vector<uint8_t> prev[2], cur[2];
    std::map<uint64_t, std::array<float, 6>> maps[2];  // ordered: deterministic iteration
    uint64_t n_tracks = 0, n_frames = 0;
};

namespace {

using TrackMap = std::map<uint64_t, std::array<float, 6>>;

TrackMap track_map(orc_tracker *t, const std::vector<uint8_t> &pa, const std::vector<uint8_t> &pb,
                   const TrackMap &in) {
    const int n = (int)in.size();
    std::vector<uint64_t> ids(n);
    std::vector<float> Tin((size_t)n * 6), Tout((size_t)n * 6);
    std::vector<int32_t> st(n);
    int k = 0;
    for (const auto &kv : in) {
        ids[k] = kv.first;
        std::copy(kv.second.begin(), kv.second.end(), &Tin[(size_t)6 * k]);
        ++k;
    }
    track_points(pa.data(), pb.data(), t->w, t->h, t->levels, n, Tin.data(), t->max_it, t->thr,
                 Tout.data(), st.data(), nullptr, nullptr, t->nthreads);
    t->n_tracks += (uint64_t)n;
    TrackMap out;
    for (int i = 0; i < n; ++i)
        if (st[i] == ORC_TRACK_OK) {
            std::array<float, 6> a;
            std::copy(&Tout[(size_t)6 * i], &Tout[(size_t)6 * i + 6], a.begin());
            out.emplace(ids[i], a);
        }
    return out;
}

// add_points -- feature_tracker.rs:222-251
void add_points(const TrackMap &m, const uint8_t *img, int w, int h, uint32_t grid, std::vector<Corner> &out) {
    std::vector<uint32_t> cur;
    cur.reserve(m.size() * 2);
    for (const auto &kv : m) {
        cur.push_back(sat_u32(std::round(kv.second[2])));
        cur.push_back(sat_u32(std::round(kv.second[5])));
    }
    detect_key_points(img, (uint32_t)w, (uint32_t)h, grid, cur.data(), (int)m.size(), out);
}

void ensure_dims(orc_tracker *t, int w, int h) {
    if (t->w != w || t->h != h) {
        t->w = w;
        t->h = h;
        t->have_prev = false;
        const size_t nb = orc_pyramid_bytes(w, h, t->levels);
        for (int c = 0; c < 2; ++c) {
            t->prev[c].assign(nb, 0);
            t->cur[c].assign(nb, 0);
        }
    }
}

}  // namespace

extern "C" {

void orc_level_dims(int w0, int h0, int level, int *w, int *h) {
    *w = w0 >> level;
    *h = h0 >> level;
}

size_t orc_pyramid_bytes(int w0, int h0, int levels) {
    size_t o = 0;
    for (int i = 0; i < levels; ++i) o += (size_t)(w0 >> i) * (h0 >> i);
    return o;
}

void orc_resize_triangle_u8(const uint8_t *src, int w, int h, uint8_t *dst, int nw, int nh) {
    resize_triangle_u8(src, w, h, dst, nw, nh);
}

void orc_build_pyramid(const uint8_t *src, int w0, int h0, int levels, uint8_t *dst, int nthreads) {
    build_pyramid(src, w0, h0, levels, dst, nthreads);
}

void orc_se2_exp(const float a[3], float out[9]) {
    float m[6];
    se2_exp(a, m);
    for (int k = 0; k < 6; ++k) out[k] = m[k];
    out[6] = 0.0f;
    out[7] = 0.0f;
    out[8] = 1.0f;
}

void orc_to_luma8(const uint8_t *src, int w, int h, size_t stride_bytes, int fmt, uint8_t *out) {
    for (int y = 0; y < h; ++y) {
        const uint8_t *row = src + (size_t)y * stride_bytes;
        for (int x = 0; x < w; ++x) {
            uint32_t v;
            if (fmt == 1) {
                uint16_t c;
                std::memcpy(&c, row + 2 * x, 2);  // native-endian u16, as ImageBuffer<Luma<u16>> stores it
                v = ((uint32_t)c + 128u) / 257u;
            } else if (fmt == 2) {
                v = (2126u * row[3 * x] + 7152u * row[3 * x + 1] + 722u * row[3 * x + 2]) / 10000u;
            } else {
                v = row[x];
            }
            out[(size_t)y * w + x] = (uint8_t)v;
        }
    }
}

int orc_project(int model, const double p[9], double x, double y, double out_uv[2]) {
    double xd, yd;
    if (model == 1) {
        const double k1 = p[4], k2 = p[5], p1 = p[6], p2 = p[7], k3 = p[8];
        const double r2 = x * x + y * y;
        const double rad = 1.0 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2;
        xd = x * rad + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
        yd = y * rad + p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
    } else if (model == 2) {  // EUCM: pi(X) = (x, y) / (alpha d + (1 - alpha) z), d = sqrt(beta (x^2 + y^2) + z^2), z = 1
        const double alpha = p[4], beta = p[5];
        const double d = std::sqrt(beta * (x * x + y * y) + 1.0);
        const double den = alpha * d + (1.0 - alpha);
        if (!(den > 0.0)) return 0;
        xd = x / den;
        yd = y / den;
    } else {
        return 0;
    }
    out_uv[0] = p[0] * xd + p[2];
    out_uv[1] = p[1] * yd + p[3];
    return 1;
}

int orc_unproject(int model, const double p[9], double u, double v, double out_xy[2]) {
    const double xd = (u - p[2]) / p[0], yd = (v - p[3]) / p[1];
    out_xy[0] = out_xy[1] = std::nan("");
    if (model == 1) {
        const double k1 = p[4], k2 = p[5], p1 = p[6], p2 = p[7], k3 = p[8];
        double x = xd, y = yd;
        for (int it = 0; it < 2000; ++it) {
            const double r2 = x * x + y * y;
            const double rad = 1.0 + k1 * r2 + k2 * r2 * r2 + k3 * r2 * r2 * r2;
            const double dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
            const double dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
            const double nx = (xd - dx) / rad, ny = (yd - dy) / rad;
            const double step = std::fabs(nx - x) + std::fabs(ny - y);
            x = nx;
            y = ny;
            if (step <= 1e-14 * (1.0 + std::fabs(x) + std::fabs(y))) {  // a few ulps: the iteration may oscillate in the last bit
                out_xy[0] = x;
                out_xy[1] = y;
                return 1;
            }
        }
        return 0;
    }
    if (model == 2) {
        const double alpha = p[4], beta = p[5];
        const double r2 = xd * xd + yd * yd;
        const double disc = 1.0 - (2.0 * alpha - 1.0) * beta * r2;
        if (disc < 0.0) return 0;
        const double mz = (1.0 - beta * alpha * alpha * r2) / (alpha * std::sqrt(disc) + (1.0 - alpha));
        if (!(mz > 0.0)) return 0;
        out_xy[0] = xd / mz;
        out_xy[1] = yd / mz;
        return 1;
    }
    return 0;
}

void orc_image_grad(const uint8_t *img, int w, int h, float x, float y, float out[3]) {
    Level l{img, w, h};
    image_grad(l, x, y, out);
}

int orc_inbound(int w, int h, float x, float y, uint32_t radius) { return inbound(w, h, x, y, radius) ? 1 : 0; }

void orc_pattern52_new(const uint8_t *img, int w, int h, float px, float py, orc_pattern52 *out) {
    Level l{img, w, h};
    pattern52_new(l, px, py, out, nullptr);
}

int orc_pattern52_residual(const orc_pattern52 *pat, const uint8_t *img, int w, int h, const float *xy, float *res) {
    Level l{img, w, h};
    return pattern52_residual(*pat, l, xy, xy + NP, res, nullptr) ? 1 : 0;
}

int orc_track_point_at_level(const uint8_t *img, int w, int h, const orc_pattern52 *pat, float T[6], int max_it,
                             float thr, int *iters) {
    Level l{img, w, h};
    int it = 0;
    const int st = track_point_at_level(l, *pat, T, max_it, thr, &it, nullptr);
    if (iters) *iters = it;
    return st;
}

void orc_track_points(const uint8_t *pyr_a, const uint8_t *pyr_b, int w0, int h0, int levels, int n,
                      const float *in_T, int max_it, float thr, float *out_T, int32_t *status, int32_t *iters,
                      float *margins, int nthreads) {
    track_points(pyr_a, pyr_b, w0, h0, levels, n, in_T, max_it, thr, out_T, status, iters, margins, nthreads);
}

int orc_corners_fast9(const uint8_t *img, int w, int h, int threshold, uint32_t *xs, uint32_t *ys, float *scores,
                      int cap) {
    std::vector<Corner> c;
    corners_fast9(img, w, h, threshold, c);
    for (int i = 0; i < (int)c.size() && i < cap; ++i) {
        xs[i] = c[i].x;
        ys[i] = c[i].y;
        scores[i] = c[i].score;
    }
    return (int)c.size();
}

int orc_detect_key_points(const uint8_t *img, int w, int h, uint32_t grid, const uint32_t *cur_xy, int n_cur,
                          uint32_t *xs, uint32_t *ys, float *scores, int cap) {
    std::vector<Corner> c;
    detect_key_points(img, (uint32_t)w, (uint32_t)h, grid, cur_xy, n_cur, c);
    for (int i = 0; i < (int)c.size() && i < cap; ++i) {
        xs[i] = c[i].x;
        ys[i] = c[i].y;
        scores[i] = c[i].score;
    }
    return (int)c.size();
}

orc_tracker *orc_tracker_create(int levels, uint32_t grid, int max_it, float thr, int nthreads) {
    orc_tracker *t = new orc_tracker();
    t->levels = levels;
    t->grid = grid;
    t->max_it = max_it;
    t->thr = thr;
    t->nthreads = nthreads > 0 ? nthreads : 1;
    return t;
}

void orc_tracker_destroy(orc_tracker *t) { delete t; }

// StereoPatchTracker::process_frame -- feature_tracker.rs:116-187
void orc_tracker_process_stereo(orc_tracker *t, const uint8_t *left, const uint8_t *right, int w, int h) {
    ensure_dims(t, w, h);
    const uint8_t *imgs[2] = {left, right};
    for (int c = 0; c < 2; ++c) build_pyramid(imgs[c], w, h, t->levels, t->cur[c].data(), t->nthreads);
    if (t->have_prev) {
        t->maps[0] = track_map(t, t->prev[0], t->cur[0], t->maps[0]);
        t->maps[1] = track_map(t, t->prev[1], t->cur[1], t->maps[1]);
    }
    std::vector<Corner> fresh;
    add_points(t->maps[0], left, w, h, t->grid, fresh);
    TrackMap tmp0;
    for (size_t i = 0; i < fresh.size(); ++i)
        tmp0.emplace((uint64_t)i, std::array<float, 6>{1.0f, 0.0f, (float)fresh[i].x, 0.0f, 1.0f, (float)fresh[i].y});
    TrackMap tmp1 = track_map(t, t->cur[0], t->cur[1], tmp0);
    for (const auto &kv : tmp0) {
        auto it = tmp1.find(kv.first);
        if (it != tmp1.end()) {
            t->maps[0].emplace(t->last_id, kv.second);
            t->maps[1].emplace(t->last_id, it->second);
            t->last_id += 1;
        }
    }
    for (int c = 0; c < 2; ++c) t->prev[c].swap(t->cur[c]);
    t->have_prev = true;
    t->n_frames += 1;
}

// PatchTracker::process_frame -- feature_tracker.rs:43-77 (uses the handle's grid/max_it/thr;
// the reference hard-codes 30 / 30 / 0.005 for the mono tracker)
void orc_tracker_process_mono(orc_tracker *t, const uint8_t *img, int w, int h) {
    ensure_dims(t, w, h);
    build_pyramid(img, w, h, t->levels, t->cur[0].data(), t->nthreads);
    if (t->have_prev) t->maps[0] = track_map(t, t->prev[0], t->cur[0], t->maps[0]);
    std::vector<Corner> fresh;
    add_points(t->maps[0], img, w, h, t->grid, fresh);
    for (const Corner &c : fresh) {
        t->maps[0].emplace(t->last_id, std::array<float, 6>{1.0f, 0.0f, (float)c.x, 0.0f, 1.0f, (float)c.y});
        t->last_id += 1;
    }
    t->prev[0].swap(t->cur[0]);
    t->have_prev = true;
    t->n_frames += 1;
}

int orc_tracker_num_tracks(const orc_tracker *t, int cam) { return (int)t->maps[cam & 1].size(); }

int orc_tracker_get_tracks(const orc_tracker *t, int cam, uint64_t *ids, float *T, int cap) {
    int k = 0;
    for (const auto &kv : t->maps[cam & 1]) {
        if (k >= cap) break;
        ids[k] = kv.first;
        std::copy(kv.second.begin(), kv.second.end(), T + (size_t)6 * k);
        ++k;
    }
    return (int)t->maps[cam & 1].size();
}

void orc_tracker_remove_ids(orc_tracker *t, const uint64_t *ids, int n) {
    for (int i = 0; i < n; ++i) {
        t->maps[0].erase(ids[i]);
        t->maps[1].erase(ids[i]);
    }
}

void orc_tracker_counters(const orc_tracker *t, uint64_t out[2]) {
    out[0] = t->n_tracks;
    out[1] = t->n_frames;
}

}  // extern "C"

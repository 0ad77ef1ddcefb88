// microbench.cu -- in-run FP32 issue-rate microbenchmark: the denominator of the LK kernel's roofline.
//
// The LK kernel (lk.cu) is bound by the FP32 / issue pipes, not by HBM, and the nominal
// "148 SMs x 128 lanes x 2 flop x clock" figure assumes every instruction is an FFMA at the boost clock.  This
// measures, on the device and at the clocks of the run, (a) the FFMA rate and (b) the rate of separately rounded
// FMUL + FADD pairs (what un-contracted reference arithmetic costs: two issue slots per multiply-add).
#include "common.cuh"

namespace rsvio {

namespace {

constexpr int kChains = 8;      // independent dependency chains per thread (covers the 4-cycle FFMA latency)
constexpr int kInner = 512;     // chain steps per loop trip (unrolled)

template <bool FUSED>
__global__ void __launch_bounds__(256) fp32_rate_kernel(float *out, int trips, float a, float b) {
    float v[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) v[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int i = 0; i < kInner / kChains; ++i)
#pragma unroll
            for (int k = 0; k < kChains; ++k)
                v[k] = FUSED ? fmaf(v[k], a, b) : __fadd_rn(__fmul_rn(v[k], a), b);
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += v[k];
    if (s == 123.456f) out[0] = s;  // keeps the chains alive; practically never taken
}

}  // namespace

// TFLOP/s (2 flop per multiply-add) of the fused and the un-fused form on the current device
cudaError_t fp32_rate(float *ffma_tflops, float *fmul_fadd_tflops, cudaStream_t st) {
    float *d_out = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_out, sizeof(float));
    if (e != cudaSuccess) return e;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    const int blocks = sm_count() * 8, trips = 64;
    const double flop = 2.0 * (double)blocks * 256.0 * (double)trips * (double)kInner;
    for (int mode = 0; mode < 2 && e == cudaSuccess; ++mode) {
        float best = 0.0f;
        for (int rep = 0; rep < 4 && e == cudaSuccess; ++rep) {  // first repetition warms up
            cudaEventRecord(e0, st);
            if (mode == 0)
                fp32_rate_kernel<true><<<blocks, 256, 0, st>>>(d_out, trips, 0.999f, 1e-3f);
            else
                fp32_rate_kernel<false><<<blocks, 256, 0, st>>>(d_out, trips, 0.999f, 1e-3f);
            cudaEventRecord(e1, st);
            e = cudaEventSynchronize(e1);
            float ms = 0.0f;
            if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms > 0.0f) {
                const float tf = (float)(flop / (ms * 1e-3) / 1e12);
                best = tf > best ? tf : best;
            }
        }
        *(mode == 0 ? ffma_tflops : fmul_fadd_tflops) = best;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaFree(d_out);
    return e == cudaSuccess ? cudaGetLastError() : e;
}

}  // namespace rsvio

// lpc.cuh -- warp-level LPC analysis / synthesis (kernels K1, K2, K3, K6 of SURVEY.md 2).
//
// One warp processes one 2048-sample signal.  The double-precision front end
// reproduces the reference's operation order and rounding exactly
// (src/lpc/residue_generator.cpp:12-96, src/lpc/linear_predictor.cpp:16-61):
// parallelism is taken ACROSS lags / coefficients, never across the terms of a sum.
#pragma once

#include "common.cuh"
#include "lpc_tables.cuh"

namespace selab200 {

// Per-warp shared-memory workspace: predictor (shared by encoder and decoder) ...
struct CoefSmem {
    double  t[104];      // step-up scratch
    long long c[104];    // Q35 coefficients c[0..order]
    int32_t q[104];      // quantised reflection coefficients
};
// ... and the analysis path on top of it.
struct LpcSmem {
    double  ring[512];   // x tile for the mean pass, then the d = x - mean ring (swizzled)
    double  ac[128];     // autocorrelation (raw, then normalised)
    double  kk[104];     // reflection coefficients k[0..99]
    CoefSmem cf;
};

// ---------------------------------------------------------------------------
// ring addressing: 512 doubles = 256 chunks of 16 B.  Odd 128-byte rows have
// their chunk pairs swapped so that the "own window" LDS.128 of the
// autocorrelation (lanes 32 B apart) is bank-conflict free.
__device__ __forceinline__ int ring_chunk(int chunk)
{
    chunk &= 255;
    return chunk ^ ((chunk >> 3) & 1);
}
__device__ __forceinline__ int ring_index(int p) // logical sample index (may be negative)
{
    return (ring_chunk(p >> 1) << 1) | (p & 1);
}

// x[j] = (double)s[j] / 32767  (quantizeSamples, residue_generator.cpp:12-18)
__device__ __forceinline__ double sample_to_x(int s) { return ddiv((double)s, 32767.0); }

// ---------------------------------------------------------------------------
// K1: mean-removed autocorrelation, lags 0..100, + normalisation.
// generateAutoCorrelation (residue_generator.cpp:20-45).  Result in sm.ac[0..100].
//
//  - mean: ONE sequential chain  sum = sum + x[j]  over j (all lanes compute it
//    redundantly from a broadcast tile so no final broadcast is needed);
//  - lane l owns lags 4l..4l+3 (lanes 0..25 useful).  For step j the four products
//    are d[j]*d[j-4l-m]; each accumulator is a sequential chain over j, exactly
//    `ac[i] += d[j]*d[j-i]` with the multiply rounded before the add;
//  - j < i terms are fed as d[negative] = +0.0: acc + (+-0) leaves a +0.0
//    accumulator unchanged, so starting the chain at j = 0 instead of j = i is
//    bit-identical.
template <typename Sig>
__device__ void warp_autocorrelation(const Sig &sig, LpcSmem &sm)
{
    const int lane = lane_id();

    // ---- mean ----
    double sum = 0.0;
    for (int tile = 0; tile < kFrame / 256; tile++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
            int j = tile * 256 + r * 32 + lane;
            sm.ring[r * 32 + lane] = sample_to_x(sig.at(j));
        }
        __syncwarp();
        const double2 *x2 = reinterpret_cast<const double2 *>(sm.ring);
#pragma unroll 8
        for (int t = 0; t < 128; t++) {
            double2 v = x2[t];
            sum = dadd(sum, v.x);
            sum = dadd(sum, v.y);
        }
        __syncwarp();
    }
    const double mean = ddiv(sum, (double)kFrame);

    // ---- autocorrelation ----
    // logical d[-128..-1] = 0  -> chunks 192..255
    {
        double2 *r2 = reinterpret_cast<double2 *>(sm.ring);
        r2[192 + lane] = make_double2(0.0, 0.0);
        r2[224 + lane] = make_double2(0.0, 0.0);
    }
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    double p1 = 0.0, p2 = 0.0, p3 = 0.0; // d[4G-1], d[4G-2], d[4G-3]
    const double2 *ring2 = reinterpret_cast<const double2 *>(sm.ring);

    for (int tile = 0; tile < kFrame / 256; tile++) {
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; r++) {
            int j = tile * 256 + r * 32 + lane;
            sm.ring[ring_index(j)] = dsub(sample_to_x(sig.at(j)), mean);
        }
        __syncwarp();
#pragma unroll 2
        for (int g = tile * 64; g < tile * 64 + 64; g++) {
            const int G = g - lane;
            const double2 o0 = ring2[ring_chunk(2 * G)];
            const double2 o1 = ring2[ring_chunk(2 * G + 1)];
            const double2 b0 = ring2[ring_chunk(2 * g)];
            const double2 b1 = ring2[ring_chunk(2 * g + 1)];
            const double c0 = o0.x, c1 = o0.y, c2 = o1.x, c3 = o1.y; // d[4G..4G+3]
            // u = 0  (j = 4g)
            acc0 = dadd(acc0, dmul(b0.x, c0));
            acc1 = dadd(acc1, dmul(b0.x, p1));
            acc2 = dadd(acc2, dmul(b0.x, p2));
            acc3 = dadd(acc3, dmul(b0.x, p3));
            // u = 1
            acc0 = dadd(acc0, dmul(b0.y, c1));
            acc1 = dadd(acc1, dmul(b0.y, c0));
            acc2 = dadd(acc2, dmul(b0.y, p1));
            acc3 = dadd(acc3, dmul(b0.y, p2));
            // u = 2
            acc0 = dadd(acc0, dmul(b1.x, c2));
            acc1 = dadd(acc1, dmul(b1.x, c1));
            acc2 = dadd(acc2, dmul(b1.x, c0));
            acc3 = dadd(acc3, dmul(b1.x, p1));
            // u = 3
            acc0 = dadd(acc0, dmul(b1.y, c3));
            acc1 = dadd(acc1, dmul(b1.y, c2));
            acc2 = dadd(acc2, dmul(b1.y, c1));
            acc3 = dadd(acc3, dmul(b1.y, c0));
            p1 = c3;
            p2 = c2;
            p3 = c1;
        }
    }
    __syncwarp();
    sm.ac[4 * lane + 0] = acc0;
    sm.ac[4 * lane + 1] = acc1;
    sm.ac[4 * lane + 2] = acc2;
    sm.ac[4 * lane + 3] = acc3;
    __syncwarp();
    // normalise (residue_generator.cpp:41-44): ac[i] /= ac[0] for i >= 1, then ac[0] = 1
    const double ac0 = sm.ac[0];
    __syncwarp();
#pragma unroll
    for (int t = 0; t < 4; t++) {
        int i = lane + 32 * t;
        if (i >= 1 && i <= kMaxOrder)
            sm.ac[i] = ddiv(sm.ac[i], ac0);
    }
    if (lane == 0)
        sm.ac[0] = 1.0;
    __syncwarp();
}

// ---------------------------------------------------------------------------
// K2a: Schur recursion -> 100 reflection coefficients in sm.kk
// generateReflectionCoefficients (residue_generator.cpp:47-68).
// Lane l holds generator elements j = 4l..4l+3 in registers; every per-element
// update is independent, the only serial piece is k[i] = -g1[0]/err.
__device__ void warp_schur(LpcSmem &sm)
{
    const int lane = lane_id();
    double g0[4], g1[4];
#pragma unroll
    for (int m = 0; m < 4; m++) {
        int j = 4 * lane + m;
        double v = (j < kMaxOrder) ? sm.ac[j + 1] : 0.0;
        g0[m] = v;
        g1[m] = v;
    }
    double err = sm.ac[0];
    double head = shfl_d(g1[0], 0);
    double k = ddiv(-head, err);
    err = dadd(err, dmul(head, k));
    if (lane == 0)
        sm.kk[0] = k;
    for (int i = 1; i < kMaxOrder; i++) {
        const double kp = k;
        const double nxt = shfl_down_d(g1[0], 1); // g1[4(l+1)] of the previous sweep
        double up[4] = {g1[1], g1[2], g1[3], nxt};
#pragma unroll
        for (int m = 0; m < 4; m++) {
            double n1 = dadd(up[m], dmul(kp, g0[m]));
            double n0 = dadd(dmul(up[m], kp), g0[m]);
            g1[m] = n1;
            g0[m] = n0;
        }
        head = shfl_d(g1[0], 0);
        k = ddiv(-head, err);
        err = dadd(err, dmul(head, k));
        if (lane == 0)
            sm.kk[i] = k;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// K2b: order selection + 7-bit quantisation
// generateoptimalLpcOrder / quantizeReflectionCoefficients (residue_generator.cpp:70-96).
__device__ int warp_order_and_quantise(LpcSmem &sm)
{
    const int lane = lane_id();
    int best = -1;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        int i = lane + 32 * t;
        if (i < kMaxOrder && fabs(sm.kk[i]) > 0.05)
            best = i;
    }
    best = __reduce_max_sync(kFull, best);
    const int order = best < 0 ? 1 : best + 1; // default 1 (src/include/lpc.hpp:76)

    const double sqrt2 = 1.4142135623730950488016887242096; // src/include/lpc.hpp:9
#pragma unroll
    for (int t = 0; t < 4; t++) {
        int i = lane + 32 * t;
        if (i < order) {
            double kv = sm.kk[i];
            double v;
            if (i == 0)
                v = floor(dmul(64.0, dadd(-1.0, dmul(sqrt2, dsqrt(dadd(kv, 1.0))))));
            else if (i == 1)
                v = floor(dmul(64.0, dadd(-1.0, dmul(sqrt2, dsqrt(dadd(-kv, 1.0))))));
            else
                v = floor(dmul(64.0, kv));
            sm.cf.q[i] = isnan(v) ? 0 : __double2int_rz(v);
        } else if (i < 104) {
            sm.cf.q[i] = 0;
        }
    }
    __syncwarp();
    return order;
}

// ---------------------------------------------------------------------------
// K2c: de-quantise + step-up -> Q35 integer predictor in sm.c[0..order]
// LinearPredictor::dequantizeReflectionCoefficients / generatelinearPredictionCoefficients
// (src/lpc/linear_predictor.cpp:16-61).  Shared by encoder and decoder.
// Table indices are clamped to [0,127] (the reference reads out of bounds there).
__device__ __forceinline__ double dequantise(int i, int q)
{
    int idx = q + 64;
    idx = idx < 0 ? 0 : (idx > 127 ? 127 : idx);
    if (i == 0)
        return __longlong_as_double((long long)selab200_FIRST_BITS[idx]);
    if (i == 1)
        return idx == 0 ? __longlong_as_double((long long)SELAB200_SECOND0_BITS)
                        : -__longlong_as_double((long long)selab200_FIRST_BITS[idx]);
    return (double)(idx - 64) / 64.0; // exact
}

__device__ void warp_coefficients(CoefSmem &sm, int order)
{
    const int lane = lane_id();
    if (order <= 1) {
        // a single zero reflection coefficient (linear_predictor.cpp:19-22)
        if (lane == 0) {
            sm.c[0] = 0;
            sm.c[1] = 0;
        }
        __syncwarp();
        return;
    }
    for (int i = 0; i < order; i++) {
        const double ki = dequantise(i, sm.q[i]);
        const int half = i >> 1;
        for (int j = lane; j < half; j += 32) {
            double a = sm.t[j];
            double b = sm.t[i - 1 - j];
            sm.t[j] = dadd(a, dmul(ki, b));
            sm.t[i - 1 - j] = dadd(b, dmul(ki, a));
        }
        if (lane == 0) {
            if (i & 1) {
                double mid = sm.t[half];
                sm.t[half] = dadd(mid, dmul(mid, ki));
            }
            sm.t[i] = ki;
        }
        __syncwarp();
    }
    const double scale = 34359738368.0; // 2^35
    for (int m = lane; m < order; m += 32)
        sm.c[1 + m] = __double2ll_rz(dmul(scale, -sm.t[m]));
    if (lane == 0)
        sm.c[0] = 0;
    __syncwarp();
}

// ---------------------------------------------------------------------------
// K3: integer FIR residual.  generateResidues (residue_generator.cpp:98-119):
//   r[i] = s[i] - (int32)((2^34 + sum_{j=1..order} c[j]*s[i-j]) >> 35),  s[<0] = 0
// (the warm-up loop of the reference is the same sum with the missing history
// read as zero).  Integer adds wrap identically in any order, so the taps are
// free to be evaluated in any arrangement.
template <typename Sig>
__device__ void warp_fir_residual(const Sig &sig, const CoefSmem &sm, int order, int32_t *res)
{
    const int lane = lane_id();
    const long long half = 1ll << (kQ - 1);
    for (int base = 0; base < kFrame; base += 32) {
        const int i = base + lane;
        long long acc = half;
        const int taps = order < i ? order : i;
        for (int j = 1; j <= taps; j++)
            acc += sm.c[j] * (long long)sig.at(i - j);
        res[i] = sig.at(i) - (int32_t)(acc >> kQ);
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// K6: integer IIR synthesis.  SampleGenerator::generateSamples
// (src/lpc/sample_generator.cpp:11-30):
//   s[i] = r[i] - (int)((2^34 - sum_{j=1..order} c[j]*s[i-j]) >> 35),  s[<0] = 0
// A true recurrence: the warp splits the TAPS (lane l owns taps l+1, l+33, ...),
// reduces the 64-bit partial sums with shuffles, and lane 0 finishes the sample.
// `buf` holds r on entry and s on exit (in place).
__device__ void warp_iir_synthesis(const CoefSmem &sm, int order, int32_t *buf, int n)
{
    const int lane = lane_id();
    const long long half = 1ll << (kQ - 1);
    long long cj[4];
#pragma unroll
    for (int t = 0; t < 4; t++) {
        int j = lane + 1 + 32 * t;
        cj[t] = (j <= order) ? sm.c[j] : 0;
    }
    for (int i = 1; i < n; i++) {
        long long part = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            int j = lane + 1 + 32 * t;
            if (j <= order && j <= i)
                part += cj[t] * (long long)buf[i - j];
        }
        unsigned long long tot = warp_sum_u64((unsigned long long)part);
        if (lane == 0) {
            long long acc = half - (long long)tot;
            buf[i] = buf[i] - (int32_t)(acc >> kQ);
        }
        __syncwarp();
    }
}

} // namespace selab200

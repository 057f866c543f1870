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

// Per-warp shared memory.  The predictor (shared by encoder and decoder):
struct CoefSmem {
    uint32_t clo[112];           // low / high words of the Q35 coefficients c[1..], tap j at index
    int32_t  chi[112];           //   j-1, zero padded: the FIR / IIR read them in blocks of 8
    int32_t  q[104];             // quantised reflection coefficients
};
// Decoder only: warm-up bias table of the IIR, 2^34 + 2^17 * sum_{j<=t} c[j]  (t = 0..order)
struct IirSmem {
    unsigned long long pre[104];
};
// Analysis scratch (3 KB).  The ring is dead once the autocorrelation is done; the
// reflection coefficients (kk) and the step-up row (t) then live in its bytes.
struct AnalysisScratch {
    double ring[256]; // x tile for the mean pass, then the d = x - mean ring (swizzled)
    double ac[128];   // autocorrelation (raw, then normalised)
    __device__ __forceinline__ double *kk() { return ring; }        // k[0..99]
    __device__ __forceinline__ double *t() { return ring + 104; }   // step-up scratch [0..99]
};
using LpcSmem = AnalysisScratch;
// k[0..99] and the step-up row t[0..99] occupy doubles [0, 204) of the ring; from here to the end of ac[] the
// scratch is free after warp_schur() (which reads ac[] into registers first): the encoder puts its CoefSmem there.
constexpr size_t kCoefAlias = 208 * sizeof(double);

__device__ __forceinline__ long long coef_at(const CoefSmem &cf, int j) // c[j], j >= 1
{
    return (long long)(((unsigned long long)(uint32_t)cf.chi[j - 1] << 32) | cf.clo[j - 1]);
}

// ---------------------------------------------------------------------------
// ring addressing: 256 doubles = 128 chunks of 16 B (a 128-sample tile plus the 127
// samples of history the furthest lane still needs).  Odd 128-byte rows have their
// chunk pairs swapped so that the "own window" LDS.128 of the autocorrelation (lanes
// 32 B apart) is bank-conflict free.
__device__ __forceinline__ int ring_chunk(int chunk)
{
    chunk &= 127;
    return chunk ^ ((chunk >> 3) & 1);
}
__device__ __forceinline__ int ring_index(int p) // logical sample index (may be negative)
{
    return (ring_chunk(p >> 1) << 1) | (p & 1);
}

// x[j] = (double)s[j] / 32767  (quantizeSamples, residue_generator.cpp:12-18).
// Correctly rounded quotient without the division subroutine: q0 = s*rcp, one exact
// FMA residual, one FMA correction (Markstein).  Equality with IEEE division is
// verified EXHAUSTIVELY over the whole input domain |s| <= 65535 -- on the CPU in
// tests/test_host_logic.py and on the device by selab200_selftest().
__device__ __forceinline__ double sample_to_x(int s)
{
    const double rcp = 1.0 / 32767.0;
    const double a = (double)s;
    const double q0 = __dmul_rn(a, rcp);
    const double r = __fma_rn(-q0, 32767.0, a);
    return __fma_rn(r, rcp, q0);
}
__device__ __forceinline__ double sample_to_x_div(int s) { return ddiv((double)s, 32767.0); }

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
    const double mean = ddiv(sum, (double)kFrame); // exact: power of two

    // ---- autocorrelation ----
    // logical d[-128..-1] = 0  -> chunks 64..127 (the upper half of the ring)
    {
        double2 *r2 = reinterpret_cast<double2 *>(sm.ring);
        r2[64 + lane] = make_double2(0.0, 0.0);
        r2[96 + lane] = make_double2(0.0, 0.0);
    }
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    double p1 = 0.0, p2 = 0.0, p3 = 0.0; // d[4G-1], d[4G-2], d[4G-3]
    const char *ringb = reinterpret_cast<const char *>(sm.ring);
    // Groups are processed 8 at a time (g = g0 + i, g0 a multiple of 8), which makes the
    // swizzle bit of every access loop-invariant: for the broadcast group it is (i>>2)&1, a
    // compile-time constant; for the lane's own group G = g - lane it is ((i - lane)>>2)&1.
    // The physical byte offset of chunk 2G is then (32*g0 + own_off[i]) & 2047.
    int own_off[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int rel = i - lane;
        own_off[i] = 16 * (2 * rel + ((rel >> 2) & 1));
    }

    for (int tile = 0; tile < kFrame / 128; tile++) {
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 4; r++) {
            int j = tile * 128 + r * 32 + lane;
            sm.ring[ring_index(j)] = dsub(sample_to_x(sig.at(j)), mean);
        }
        __syncwarp();
        for (int blk = 0; blk < 4; blk++) {
            const int gb = (tile * 32 + blk * 8) * 32; // byte offset of chunk 2*g0 before wrapping
            const char *bbase = ringb + (gb & 2047);    // 256-byte aligned: the 8 broadcast groups never wrap
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int oo = (gb + own_off[i]) & 2047;
                const double2 o0 = *reinterpret_cast<const double2 *>(ringb + oo);
                const double2 o1 = *reinterpret_cast<const double2 *>(ringb + (oo ^ 16));
                constexpr int kSw[8] = {0, 0, 0, 0, 16, 16, 16, 16};
                const double2 b0 = *reinterpret_cast<const double2 *>(bbase + ((32 * i) ^ kSw[i]));
                const double2 b1 = *reinterpret_cast<const double2 *>(bbase + ((32 * i + 16) ^ kSw[i]));
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
    double *kk = sm.kk(); // overlays the ring, which the autocorrelation no longer needs
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
        kk[0] = k;
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
            kk[i] = k;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// K2b: order selection + 7-bit quantisation
// generateoptimalLpcOrder / quantizeReflectionCoefficients (residue_generator.cpp:70-96).
__device__ int warp_order_and_quantise(LpcSmem &sm, CoefSmem &cf)
{
    const int lane = lane_id();
    const double *kk = sm.kk();
    int best = -1;
#pragma unroll
    for (int t = 0; t < 4; t++) {
        int i = lane + 32 * t;
        if (i < kMaxOrder && fabs(kk[i]) > 0.05)
            best = i;
    }
    best = __reduce_max_sync(kFull, best);
    const int order = best < 0 ? 1 : best + 1; // default 1 (src/include/lpc.hpp:76)

    const double sqrt2 = 1.4142135623730950488016887242096; // src/include/lpc.hpp:9
#pragma unroll
    for (int t = 0; t < 4; t++) {
        int i = lane + 32 * t;
        if (i < order) {
            double kv = kk[i];
            double v;
            if (i == 0)
                v = floor(dmul(64.0, dadd(-1.0, dmul(sqrt2, dsqrt(dadd(kv, 1.0))))));
            else if (i == 1)
                v = floor(dmul(64.0, dadd(-1.0, dmul(sqrt2, dsqrt(dadd(-kv, 1.0))))));
            else
                v = floor(dmul(64.0, kv));
            cf.q[i] = isnan(v) ? 0 : __double2int_rz(v);
        } else if (i < 104) {
            cf.q[i] = 0;
        }
    }
    __syncwarp();
    return order;
}

// ---------------------------------------------------------------------------
// K2c: de-quantise + step-up -> Q35 integer predictor (cf.clo/chi, tap j at index j-1)
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

__device__ void warp_coefficients(CoefSmem &cf, double *t, int order)
{
    const int lane = lane_id();
    for (int i = lane; i < 112; i += 32) {
        cf.clo[i] = 0;
        cf.chi[i] = 0;
    }
    __syncwarp();
    if (order <= 1)
        return; // a single zero reflection coefficient -> c[1] = 0 (linear_predictor.cpp:19-22)
    for (int i = 0; i < order; i++) {
        const double ki = dequantise(i, cf.q[i]);
        const int half = i >> 1;
        for (int j = lane; j < half; j += 32) {
            double a = t[j];
            double b = t[i - 1 - j];
            t[j] = dadd(a, dmul(ki, b));
            t[i - 1 - j] = dadd(b, dmul(ki, a));
        }
        if (lane == 0) {
            if (i & 1) {
                double mid = t[half];
                t[half] = dadd(mid, dmul(mid, ki));
            }
            t[i] = ki;
        }
        __syncwarp();
    }
    const double scale = 34359738368.0; // 2^35
    for (int m = lane; m < order; m += 32) {
        const long long v = __double2ll_rz(dmul(scale, -t[m]));
        cf.clo[m] = (uint32_t)v;
        cf.chi[m] = (int32_t)(v >> 32);
    }
    __syncwarp();
}

// Signals hand the filters 8 consecutive samples at a time, biased by 2^17 so that
// they are non-negative 18-bit numbers: c*s' then needs one IMAD.WIDE.U32 (low word)
// and one IMAD (high word), and the bias is taken out once per output:
//   sum_j c[j]*s[i-j] = sum_j c[j]*s'[i-j] - 2^17 * sum_j c[j].
// Group g covers samples [8g, 8g+8); negative groups (history before the frame) read
// the zero padding in front of every channel, i.e. s = 0, as the reference's warm-up
// loop implies.
// ---------------------------------------------------------------------------
// K3: integer FIR residual.  generateResidues (residue_generator.cpp:98-119):
//   r[i] = s[i] - (int32)((2^34 + sum_{j=1..order} c[j]*s[i-j]) >> 35),  s[<0] = 0
// Integer adds wrap identically in any order.  Lane l computes 8 consecutive outputs
// per pass with a 16-sample register window that slides 8 taps per iteration: per 64
// multiply-accumulates the lane issues 128 IMADs, one 16-byte sample load and four
// broadcast coefficient loads.
template <typename Sig>
__device__ void warp_fir_residual(const Sig &sig, const CoefSmem &cf, int order, int32_t *res)
{
    const int lane = lane_id();
    const int nblk = (order + 7) >> 3;
    long long csum = 0;
    for (int j = lane + 1; j <= order; j += 32)
        csum += coef_at(cf, j);
    csum = (long long)warp_sum_u64((unsigned long long)csum);
    const unsigned long long corr = (1ull << (kQ - 1)) - ((unsigned long long)csum << 17);
    const uint4 *clo4 = reinterpret_cast<const uint4 *>(cf.clo);
    const int4 *chi4 = reinterpret_cast<const int4 *>(cf.chi);

    for (int pass = 0; pass < kFrame / 256; pass++) {
        const int g0 = pass * 32 + lane;
        uint32_t hi[8], own[8];
        sig.load8(g0, hi);
#pragma unroll
        for (int r = 0; r < 8; r++)
            own[r] = hi[r];
        unsigned long long alo[8];
        uint32_t ahi[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            alo[r] = 0;
            ahi[r] = 0;
        }
        for (int t = 0; t < nblk; t++) {
            uint32_t lo[8];
            sig.load8(g0 - t - 1, lo);
            const uint4 l0 = clo4[2 * t], l1 = clo4[2 * t + 1];
            const int4 h0 = chi4[2 * t], h1 = chi4[2 * t + 1];
            const uint32_t cl[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            const int32_t ch[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
            for (int tau = 0; tau < 8; tau++) {
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const int idx = 7 + r - tau; // window position of s[i0 + r - (8t + 1 + tau)]
                    const uint32_t w = idx < 8 ? lo[idx] : hi[idx - 8];
                    alo[r] = mad_wide_u32(cl[tau], w, alo[r]);
                    ahi[r] += (uint32_t)ch[tau] * w;
                }
            }
#pragma unroll
            for (int r = 0; r < 8; r++)
                hi[r] = lo[r];
        }
        int32_t out[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const unsigned long long p = alo[r] + ((unsigned long long)ahi[r] << 32) + corr;
            out[r] = ((int)own[r] - kSampleBias) - (int32_t)((long long)p >> kQ);
        }
        int4 *dst = reinterpret_cast<int4 *>(res + 8 * g0);
        dst[0] = make_int4(out[0], out[1], out[2], out[3]);
        dst[1] = make_int4(out[4], out[5], out[6], out[7]);
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// K6: integer IIR synthesis.  SampleGenerator::generateSamples
// (src/lpc/sample_generator.cpp:11-30):
//   s[i] = r[i] - (int)((2^34 - sum_{j=1..order} c[j]*s[i-j]) >> 35),  s[<0] = 0
// A true recurrence, evaluated in TRANSPOSED form: lane l owns taps TPL*l+1..TPL*l+TPL
// and the partial sums of the outputs those taps will feed next.  When s[i] becomes
// known every lane adds c[j]*s'[i] to the accumulator of output i+j; the accumulator
// of output i+1 (tap 1, lane 0) is then complete, lane 0 finishes the sample and
// broadcasts it, and every accumulator moves one tap down (one 64-bit shuffle per
// lane, register renaming inside a lane).  Critical path per sample: one IMAD, a
// 64-bit subtract, a shift and ONE shuffle -- instead of a five-level reduction.
__device__ void warp_iir_prepare(const CoefSmem &cf, IirSmem &ii, int order)
{
    // pre[t] = 2^34 + 2^17 * sum_{j=1..t} c[j]: removes the sample bias for output t
    // (during warm-up only taps j <= t have seen a real sample)
    const int lane = lane_id();
    long long v[4], run = 0;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int j = 4 * lane + m + 1;
        run += (j <= order && j <= 112) ? coef_at(cf, j) : 0;
        v[m] = run;
    }
    unsigned long long incl = (unsigned long long)run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long tmp = __shfl_up_sync(kFull, incl, o);
        if (lane >= o)
            incl += tmp;
    }
    const unsigned long long excl = incl - (unsigned long long)run;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const int j = 4 * lane + m + 1;
        if (j < 104)
            ii.pre[j] = (1ull << (kQ - 1)) + ((excl + (unsigned long long)v[m]) << 17);
    }
    if (lane == 0)
        ii.pre[0] = 1ull << (kQ - 1);
    __syncwarp();
}

// Two subframes share a warp: lanes 0-15 run subframe A, lanes 16-31 subframe B, each
// lane owning TPL taps (TPL*15 >= order, so the last lane of a half only ever holds
// zero coefficients and its accumulators stay zero -- shfl_down past the half's edge
// returns the lane's own, zero, value: no special case).  Per step the warp issues
// 2*TPL IMADs + ~14 bookkeeping instructions for TWO samples.
template <int TPL>
__device__ void warp_iir_pair(const CoefSmem &cf, const IirSmem &ii, int order, int32_t *buf, bool active, int n, int order_max)
{
    const int hl = lane_id() & 15;
    uint32_t cl[TPL];
    int32_t ch[TPL];
#pragma unroll
    for (int m = 0; m < TPL; m++) {
        const int j = TPL * hl + m; // tap j+1
        cl[m] = j < 112 ? cf.clo[j] : 0u;
        ch[m] = j < 112 ? cf.chi[j] : 0;
    }
    // accumulators: 64-bit low-word products + a separate 32-bit column for the high-word
    // products (one IMAD.WIDE.U32 + one IMAD per tap); combined only when handed on
    unsigned long long alo[TPL];
    uint32_t ahi[TPL];
#pragma unroll
    for (int m = 0; m < TPL; m++) {
        alo[m] = 0;
        ahi[m] = 0;
    }
    const unsigned long long steady = ii.pre[order];
    uint32_t sp = (uint32_t)(buf[0] + kSampleBias); // s[0] = r[0]
    const bool writer = active && hl == 0;
    // step(u, i, base): consumes s'[i] in `sp`, produces s[i+1].  Slot m of this step
    // lives in physical register (m + u) % TPL, so the per-step slot shift is free.
#define SELAB200_IIR_STEP(u, i, base)                                                            \
    {                                                                                            \
        _Pragma("unroll") for (int m = 0; m < TPL; m++)                                          \
        {                                                                                        \
            alo[(m + (u)) % TPL] = mad_wide_u32(cl[m], sp, alo[(m + (u)) % TPL]);               \
            ahi[(m + (u)) % TPL] += (uint32_t)ch[m] * sp;                                        \
        }                                                                                        \
        const unsigned long long full0 = alo[(u) % TPL] + ((unsigned long long)ahi[(u) % TPL] << 32); \
        const unsigned long long incoming = __shfl_down_sync(kFull, full0, 1, 16);               \
        const unsigned long long tt = (base) - full0;                                            \
        int vnext = buf[(i) + 1] - (int32_t)((long long)tt >> kQ);                               \
        vnext = __shfl_sync(kFull, vnext, 0, 16);                                                \
        if (writer)                                                                              \
            buf[(i) + 1] = vnext;                                                                \
        alo[(u) % TPL] = incoming; /* becomes the top slot of the next step */                   \
        ahi[(u) % TPL] = 0;                                                                      \
        sp = (uint32_t)(vnext + kSampleBias);                                                    \
    }
    int i = 0;
    const int last = n - 1; // steps i = 0 .. last-1
    // warm-up: outputs 1..order use the prefix table (the longer of the two orders decides)
    const int warm = order_max < last ? order_max : last;
    const int warm_groups = (warm + TPL - 1) / TPL;
    for (int gq = 0; gq < warm_groups; gq++) {
#pragma unroll
        for (int u = 0; u < TPL; u++) {
            if (i < last) {
                const int t = i + 1;
                const unsigned long long base = ii.pre[t < order ? t : order];
                SELAB200_IIR_STEP(u, i, base);
                i++;
            }
        }
    }
    // steady state
    for (; i + TPL <= last;) {
#pragma unroll
        for (int u = 0; u < TPL; u++) {
            SELAB200_IIR_STEP(u, i, steady);
            i++;
        }
    }
#pragma unroll
    for (int u = 0; u < TPL; u++) {
        if (i < last) {
            SELAB200_IIR_STEP(u, i, steady);
            i++;
        }
    }
#undef SELAB200_IIR_STEP
    __syncwarp();
}

// Synthesis of two subframes in one warp.  cf/buf/order/active are PER HALF (lanes
// 0-15: A, lanes 16-31: B); cf.pre must be ready (warp_iir_prepare).  buf: r on entry,
// s on exit (in place), n samples.  An inactive half computes but never stores.
__device__ void warp_iir_synthesis_pair(const CoefSmem &cf, const IirSmem &ii, int order, int32_t *buf, bool active, int n)
{
    const int other = __shfl_xor_sync(kFull, order, 16);
    const int order_max = order > other ? order : other;
    if (order_max <= 30)
        warp_iir_pair<2>(cf, ii, order, buf, active, n, order_max);
    else if (order_max <= 60)
        warp_iir_pair<4>(cf, ii, order, buf, active, n, order_max);
    else
        warp_iir_pair<8>(cf, ii, order, buf, active, n, order_max);
}

// ---------------------------------------------------------------------------
// K6, batch form: FOUR subframes per warp (quarter q = lane>>3 owns one), eight lanes x
// TPL taps each (TPL*7 >= order, so a quarter's last lane only ever holds zero
// coefficients and shfl_down past the quarter's edge returns its own, zero, value).
// No shared-memory sample planes: residues arrive from global memory 16 at a time per
// quarter (one coalesced 64-byte load, one block ahead) and are handed to lane 0 of the
// quarter by shuffle; finished samples are parked in a 64-byte staging row per quarter
// and leave 16 at a time.  Per 16 outputs a quarter therefore touches shared memory
// 16 + 1 times and global memory twice.
struct QuadIo {
    const int32_t *res;   // this quarter's residues (global), 2048 ints, 16-byte aligned
    int32_t *stage;       // this quarter's staging rows in shared memory: [2][16]
};

template <int TPL>
struct QuadState {
    uint32_t cl[TPL];
    int32_t ch[TPL];
    unsigned long long alo[TPL];
    uint32_t ahi[TPL];
    uint32_t sp;
};

// One block of 16 outputs t = 16*B + e.  WARM: bias term from the prefix table (t <= order).
template <int TPL, bool WARM>
__device__ __forceinline__ void quad_block(QuadState<TPL> &st, const IirSmem &ii, int order, unsigned long long steady,
                                           int B, int2 rcur, int32_t *stage_row, bool writer)
{
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int t = 16 * B + e;
        const int rt = __shfl_sync(kFull, (e & 1) ? rcur.y : rcur.x, e >> 1, 8);
        int vnext;
        if (WARM && e == 0 && B == 0) { // uniform branch, only compiled into the warm-up variant
            vnext = rt; // s[0] = r[0]
        } else {
            const int u = (e + 16 * TPL - 1) % TPL; // == (t - 1) % TPL, static
#pragma unroll
            for (int m = 0; m < TPL; m++) {
                st.alo[(m + u) % TPL] = mad_wide_u32(st.cl[m], st.sp, st.alo[(m + u) % TPL]);
                st.ahi[(m + u) % TPL] += (uint32_t)st.ch[m] * st.sp;
            }
            const unsigned long long full0 = st.alo[u] + ((unsigned long long)st.ahi[u] << 32);
            const unsigned long long incoming = __shfl_down_sync(kFull, full0, 1, 8);
            const unsigned long long base = WARM ? ii.pre[t < order ? t : order] : steady;
            const unsigned long long tt = base - full0;
            vnext = rt - (int32_t)((long long)tt >> kQ);
            vnext = __shfl_sync(kFull, vnext, 0, 8);
            st.alo[u] = incoming;
            st.ahi[u] = 0;
        }
        if (writer)
            stage_row[e] = vnext;
        st.sp = (uint32_t)(vnext + kSampleBias);
    }
}

// Runs the recurrence for the quarter's subframe; after every block calls
// emit(B, kx, ky) with this lane's two finished samples s[16B + 2*hl], s[16B + 2*hl + 1].
template <int TPL, typename Emit>
__device__ void warp_iir_quad(const CoefSmem &cf, const IirSmem &ii, int order, int order_max, const QuadIo io, bool has_res, Emit emit)
{
    const int hl = lane_id() & 7;
    QuadState<TPL> st;
#pragma unroll
    for (int m = 0; m < TPL; m++) {
        const int j = TPL * hl + m; // tap j+1
        st.cl[m] = j < 112 ? cf.clo[j] : 0u;
        st.ch[m] = j < 112 ? cf.chi[j] : 0;
        st.alo[m] = 0;
        st.ahi[m] = 0;
    }
    st.sp = 0;
    const unsigned long long steady = ii.pre[order];
    const bool writer = hl == 0;
    const int2 *r2 = reinterpret_cast<const int2 *>(io.res);
    // plain loads: a difference subframe writes its result back into this row behind the reads
    int2 rcur = has_res ? r2[hl] : make_int2(0, 0);
    int2 rnext = has_res ? r2[8 + hl] : make_int2(0, 0);
    const int warm_blocks = order_max / 16 + 1; // blocks that contain some t <= order
    for (int B = 0; B < kFrame / 16; B++) {
        int32_t *row = io.stage + (B & 1) * 16;
        if (B < warm_blocks)
            quad_block<TPL, true>(st, ii, order, steady, B, rcur, row, writer);
        else
            quad_block<TPL, false>(st, ii, order, steady, B, rcur, row, writer);
        rcur = rnext;
        if (B + 2 < kFrame / 16 && has_res)
            rnext = r2[(B + 2) * 8 + hl];
        __syncwarp();
        const int2 kept = *reinterpret_cast<const int2 *>(row + 2 * hl);
        emit(B, kept.x, kept.y);
    }
}

} // namespace selab200

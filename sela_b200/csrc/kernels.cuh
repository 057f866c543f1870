// kernels.cuh -- the __global__ kernels of the SELA hot path (sm_100a).
//
//   k_encode<STEREO>     fused encode of one scan unit per CTA: PCM -> autocorrelation
//                        -> Schur -> quantise -> step-up -> FIR -> Rice size -> arena
//                        offset (decoupled look-back) -> Rice pack + descriptor.
//                        STEREO: unit = frame, 3 warps (ch0, ch1, ch0-ch1);
//                        otherwise unit = subframe, 1 warp.
//   k_rice_decode        K5: one lane per Rice stream (reflection or residue streams).
//   k_synthesise         K6: CTA per frame, warp per subframe: step-up + IIR, difference
//                        reconstruction, interleave to int16 PCM.
//   k_lpc_residues / k_lpc_samples / k_rice_encode   stage-level entry points.
#pragma once

#include "lpc.cuh"
#include "rice.cuh"

namespace selab200 {

// ------------------------------------------------------------------ encode --

struct EncodeParams {
    const int16_t *pcm;            // [n_frames][2048][channels] interleaved
    uint32_t n_frames, channels;
    selab200_subframe_desc *descs; // [n_frames*channels]
    uint32_t *words;
    unsigned long long capacity;   // words
    unsigned long long *words_used;
    int32_t *status;
    uint32_t *ticket;              // workspace: dynamic unit counter
    unsigned long long *scan;      // workspace: [n_units] look-back state
};

struct SubframeResult {
    uint32_t order;
    RiceChoice refl, res;
};

// flag in the two top bits of a scan entry
constexpr unsigned long long kScanAggregate = 1ull << 62;
constexpr unsigned long long kScanPrefix    = 2ull << 62;
constexpr unsigned long long kScanValueMask = (1ull << 62) - 1;

// Exclusive prefix of the unit sizes (decoupled look-back, single thread).
// Units take tickets in increasing order, so every predecessor is resident or done.
__device__ unsigned long long scan_exclusive(unsigned long long *scan, uint32_t unit, unsigned long long agg)
{
    volatile unsigned long long *vs = scan;
    if (unit == 0) {
        vs[0] = kScanPrefix | agg;
        __threadfence();
        return 0;
    }
    vs[unit] = kScanAggregate | agg;
    __threadfence();
    unsigned long long sum = 0;
    uint32_t p = unit - 1;
    while (true) {
        unsigned long long v = vs[p];
        if ((v >> 62) == 0)
            continue; // predecessor has not published yet
        sum += v & kScanValueMask;
        if (v & kScanPrefix)
            break;
        p--;
    }
    vs[unit] = kScanPrefix | (sum + agg);
    __threadfence();
    return sum;
}

__device__ __forceinline__ void write_desc(selab200_subframe_desc *d, uint32_t channel, uint32_t type,
                                           uint32_t parent, const SubframeResult &r,
                                           unsigned long long refl_off)
{
    selab200_subframe_desc v;
    v.channel = (uint8_t)channel;
    v.subframe_type = (uint8_t)type;
    v.parent_channel = (uint8_t)parent;
    v.refl_rice_param = (uint8_t)r.refl.k;
    v.refl_words = (uint16_t)r.refl.words;
    v.lpc_order = (uint8_t)r.order;
    v.res_rice_param = (uint8_t)r.res.k;
    v.res_words = (uint16_t)r.res.words;
    v.samples = (uint16_t)kFrame;
    v.reserved = 0;
    v.refl_offset = refl_off;
    v.res_offset = refl_off + r.refl.words;
    *d = v;
}

template <bool STEREO>
__global__ void __launch_bounds__(STEREO ? 96 : 32) k_encode(EncodeParams p)
{
    constexpr int kWarps = STEREO ? 3 : 1;
    constexpr int kChan = STEREO ? 2 : 1; // channels held in shared memory per CTA
    constexpr int kRow = kHistoryPad + kFrame;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int16_t *s16 = reinterpret_cast<int16_t *>(smem_raw);                        // [kChan][pad + 2048]
    WarpScratch *scratch_all = reinterpret_cast<WarpScratch *>(smem_raw + kChan * kRow * 2);
    CoefSmem *coef_all = reinterpret_cast<CoefSmem *>(scratch_all + kWarps);
    SubframeResult *results = reinterpret_cast<SubframeResult *>(coef_all + kWarps);
    unsigned long long *base_slot = reinterpret_cast<unsigned long long *>(results + 4);
    uint32_t *unit_slot = reinterpret_cast<uint32_t *>(base_slot + 1);

    const int warp = warp_id(), lane = lane_id();
    if (threadIdx.x == 0)
        *unit_slot = atomicAdd(p.ticket, 1u);
    __syncthreads();
    const uint32_t unit = *unit_slot;
    const uint32_t n_units = STEREO ? p.n_frames : p.n_frames * p.channels;
    if (unit >= n_units)
        return;
    const uint32_t frame = STEREO ? unit : unit / p.channels;
    const uint32_t chan0 = STEREO ? 0 : unit % p.channels;

    // ---- stage the PCM of this unit (de-interleave to planar int16, zero history in front) ----
    const int16_t *src = p.pcm + (size_t)frame * kFrame * p.channels;
    for (int j = threadIdx.x; j < kChan * kHistoryPad / 2; j += blockDim.x) {
        const int c = j / (kHistoryPad / 2), o = j % (kHistoryPad / 2);
        reinterpret_cast<uint32_t *>(s16 + c * kRow)[o] = 0;
    }
    int16_t *ch0 = s16 + kHistoryPad;
    if (STEREO) {
        int16_t *ch1 = ch0 + kRow;
        const uint4 *src128 = reinterpret_cast<const uint4 *>(src); // 4 stereo sample pairs per load
        for (int j = threadIdx.x; j < kFrame / 4; j += blockDim.x) {
            const uint4 v = src128[j];
            const uint32_t l0 = __byte_perm(v.x, v.y, 0x5410), r0 = __byte_perm(v.x, v.y, 0x7632);
            const uint32_t l1 = __byte_perm(v.z, v.w, 0x5410), r1 = __byte_perm(v.z, v.w, 0x7632);
            reinterpret_cast<uint2 *>(ch0)[j] = make_uint2(l0, l1);
            reinterpret_cast<uint2 *>(ch1)[j] = make_uint2(r0, r1);
        }
    } else {
        for (int j = threadIdx.x; j < kFrame; j += blockDim.x)
            ch0[j] = src[(size_t)j * p.channels + chan0];
    }
    __syncthreads();

    // ---- per-warp analysis ----
    Signal sig;
    if (STEREO) {
        sig.a = (warp == 1) ? ch0 + kRow : ch0;
        sig.b = (warp == 2) ? ch0 + kRow : nullptr;
    } else {
        sig.a = ch0;
        sig.b = nullptr;
    }
    WarpScratch &scratch = scratch_all[warp];
    CoefSmem &cf = coef_all[warp];
    int32_t *res = scratch.res;
    warp_autocorrelation(sig, scratch.a);
    warp_schur(scratch.a);
    const int order = warp_order_and_quantise(scratch.a, cf);
    warp_coefficients(cf, scratch.a.t, order);
    warp_fir_residual(sig, cf, order, res);

    const int32_t *qv = cf.q;
    auto q_at = [qv](int i) { return qv[i]; };
    auto r_at = [res](int i) { return res[i]; };
    SubframeResult mine;
    mine.order = order;
    mine.refl = warp_rice_choose(q_at, order);
    mine.res = warp_rice_choose(r_at, kFrame);
    if (lane == 0)
        results[warp] = mine;
    __syncthreads();

    // ---- stereo decision + arena offset ----
    bool emit = true;
    uint32_t channel = chan0, type = 0, parent = chan0;
    unsigned long long my_off = 0, unit_words;
    if (STEREO) {
        const unsigned long long w0 = (unsigned long long)results[0].refl.words + results[0].res.words;
        const unsigned long long wa = (unsigned long long)results[1].refl.words + results[1].res.words;
        const unsigned long long wd = (unsigned long long)results[2].refl.words + results[2].res.words;
        const bool diff_wins = wd < wa; // strictly smaller (src/frame/frame_encoder.cpp:63-72)
        unit_words = w0 + (diff_wins ? wd : wa);
        if (warp == 0) {
            channel = 0; parent = 0;
        } else {
            emit = (warp == 2) == diff_wins;
            channel = 1;
            type = diff_wins ? 1 : 0;
            parent = diff_wins ? 0 : 1;
            my_off = w0;
        }
    } else {
        unit_words = (unsigned long long)mine.refl.words + mine.res.words;
    }
    if (threadIdx.x == 0) {
        *base_slot = scan_exclusive(p.scan, unit, unit_words);
        if (unit == n_units - 1)
            *p.words_used = *base_slot + unit_words;
    }
    __syncthreads();
    const unsigned long long base = *base_slot + my_off;

    if (!emit)
        return;
    if (mine.refl.words > 0xffffu || mine.res.words > 0xffffu) {
        // the uint16 word-count fields cannot hold this (the reference would truncate)
        if (lane == 0)
            raise_status(p.status, SELAB200_ERR_RANGE);
        return;
    }
    if (*base_slot + unit_words > p.capacity) {
        if (lane == 0)
            raise_status(p.status, SELAB200_ERR_CAPACITY);
        return;
    }
    warp_rice_pack(q_at, order, mine.refl.k, mine.refl.words, p.words + base);
    warp_rice_pack(r_at, kFrame, mine.res.k, mine.res.words, p.words + base + mine.refl.words);
    if (lane == 0)
        write_desc(p.descs + (size_t)frame * p.channels + channel, channel, type, parent, mine, base);
}

template <bool STEREO>
constexpr size_t encode_smem_bytes()
{
    constexpr int kWarps = STEREO ? 3 : 1;
    constexpr int kChan = STEREO ? 2 : 1;
    return (size_t)kChan * (kHistoryPad + kFrame) * 2 + kWarps * (sizeof(WarpScratch) + sizeof(CoefSmem)) +
           4 * sizeof(SubframeResult) + 16;
}

// ------------------------------------------------------------------ decode --

__device__ __forceinline__ bool desc_ok(const selab200_subframe_desc &d, uint32_t channels,
                                        unsigned long long n_words)
{
    return d.channel < channels && d.parent_channel < channels && d.subframe_type <= 1 &&
           d.lpc_order <= kMaxOrder && d.refl_rice_param < 32 && d.res_rice_param < 32 &&
           d.samples == kFrame && d.refl_offset + d.refl_words <= n_words &&
           d.res_offset + d.res_words <= n_words &&
           !(d.subframe_type == 1 && d.parent_channel == d.channel);
}

struct DecodeParams {
    const selab200_subframe_desc *descs;
    uint32_t n_frames, channels;
    const uint32_t *words;
    unsigned long long n_words;
    int16_t *pcm_out;
    int32_t *status;
    int32_t *ws_q;   // [n_sub][128]
    int32_t *ws_res; // [n_sub][2048]
};

// K5: one lane per stream.  which = 0: reflection streams, 1: residue streams.
__global__ void __launch_bounds__(128) k_rice_decode(DecodeParams p, int which)
{
    const uint32_t sub = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_sub = p.n_frames * p.channels;
    if (sub >= n_sub)
        return;
    const selab200_subframe_desc d = p.descs[sub];
    if (!desc_ok(d, p.channels, p.n_words)) {
        raise_status(p.status, SELAB200_ERR_BITSTREAM);
        return;
    }
    bool ok;
    if (which == 0)
        ok = lane_rice_decode(p.words + d.refl_offset, d.refl_words, d.refl_rice_param, d.lpc_order,
                              p.ws_q + (size_t)sub * 128);
    else
        ok = lane_rice_decode(p.words + d.res_offset, d.res_words, d.res_rice_param, d.samples,
                              p.ws_res + (size_t)sub * kFrame);
    if (!ok)
        raise_status(p.status, SELAB200_ERR_BITSTREAM);
}

// K6: CTA per frame, warp per subframe.
__global__ void k_synthesise(DecodeParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t ch = p.channels;
    int32_t *planes = reinterpret_cast<int32_t *>(smem_raw);                  // [ch][2048], by channel
    CoefSmem *coef_all = reinterpret_cast<CoefSmem *>(planes + (size_t)ch * kFrame);
    double *t_all = reinterpret_cast<double *>(coef_all + ch);                // [ch][104]
    int *meta = reinterpret_cast<int *>(t_all + (size_t)ch * 104); // [ch] type, [ch] parent, [1] valid

    const uint32_t frame = blockIdx.x;
    const int warp = warp_id(), lane = lane_id();
    const selab200_subframe_desc *fd = p.descs + (size_t)frame * ch;

    if (threadIdx.x == 0) {
        // frame-level validation: channels form a permutation, parents are independent subframes
        bool ok = true;
        unsigned seen = 0;
        for (uint32_t i = 0; i < ch && ok; i++) {
            const selab200_subframe_desc d = fd[i];
            ok = desc_ok(d, ch, p.n_words) && !((seen >> d.channel) & 1);
            if (ok) {
                seen |= 1u << d.channel;
                meta[d.channel] = d.subframe_type;
                meta[ch + d.channel] = d.parent_channel;
            }
        }
        for (uint32_t c = 0; c < ch && ok; c++)
            if (meta[c] == 1 && meta[meta[ch + c]] != 0)
                ok = false;
        meta[2 * ch] = ok;
        if (!ok)
            raise_status(p.status, SELAB200_ERR_BITSTREAM);
    }
    __syncthreads();
    const bool valid = meta[2 * ch];
    int16_t *out = p.pcm_out + (size_t)frame * kFrame * ch;
    if (!valid) {
        for (uint32_t e = threadIdx.x; e < kFrame * ch; e += blockDim.x)
            out[e] = 0;
        return;
    }

    {
        const uint32_t sub = frame * ch + warp;
        const selab200_subframe_desc d = fd[warp];
        CoefSmem &cf = coef_all[warp];
        int32_t *buf = planes + (size_t)d.channel * kFrame;
        const int order = d.lpc_order;
        for (int i = lane; i < 104; i += 32)
            cf.q[i] = i < order ? p.ws_q[(size_t)sub * 128 + i] : 0;
        const int32_t *r = p.ws_res + (size_t)sub * kFrame;
        for (int i = lane; i < kFrame; i += 32)
            buf[i] = r[i];
        __syncwarp();
        // order 0 behaves like order 1 with a zero predictor (linear_predictor.cpp:19-22)
        warp_coefficients(cf, t_all + warp * 104, order);
        warp_iir_synthesis(cf, order, buf, kFrame);
    }
    __syncthreads();

    // difference reconstruction + interleave (frame_decoder.cpp:64-67, wav_file.cpp:244-266)
    for (uint32_t e = threadIdx.x; e < kFrame * ch; e += blockDim.x) {
        const uint32_t c = e % ch, j = e / ch;
        int32_t v = planes[(size_t)c * kFrame + j];
        if (meta[c] == 1)
            v = planes[(size_t)meta[ch + c] * kFrame + j] - v;
        out[e] = (int16_t)(uint16_t)v;
    }
}

inline size_t synthesise_smem_bytes(uint32_t ch)
{
    return (size_t)ch * kFrame * 4 + ch * (sizeof(CoefSmem) + 104 * sizeof(double)) + (2 * ch + 1) * sizeof(int);
}

// ------------------------------------------------------------ stage level --

// lpc::ResidueGenerator::process for one signal per (1-warp) CTA.
__global__ void __launch_bounds__(32) k_lpc_residues(const int32_t *samples, uint32_t n_sub, uint8_t *order_out,
                                                     int32_t *q_out, int32_t *residues)
{
    __shared__ __align__(16) WarpScratch scratch;
    __shared__ __align__(16) CoefSmem cf;
    __shared__ __align__(16) int32_t s_pad[kHistoryPad + kFrame];
    const uint32_t sub = blockIdx.x;
    const int lane = lane_id();
    int32_t *s = s_pad + kHistoryPad;
    for (int i = lane; i < kHistoryPad; i += 32)
        s_pad[i] = 0;
    for (int i = lane; i < kFrame; i += 32)
        s[i] = samples[(size_t)sub * kFrame + i];
    __syncwarp();
    PlainSignal sig{s};
    warp_autocorrelation(sig, scratch.a);
    warp_schur(scratch.a);
    const int order = warp_order_and_quantise(scratch.a, cf);
    warp_coefficients(cf, scratch.a.t, order);
    warp_fir_residual(sig, cf, order, scratch.res);
    for (int i = lane; i < kFrame; i += 32)
        residues[(size_t)sub * kFrame + i] = scratch.res[i];
    for (int i = lane; i < kMaxOrder; i += 32)
        q_out[(size_t)sub * kMaxOrder + i] = i < order ? cf.q[i] : 0;
    if (lane == 0)
        order_out[sub] = (uint8_t)order;
}

// lpc::SampleGenerator::process for one signal per (1-warp) CTA.
__global__ void __launch_bounds__(32) k_lpc_samples(const int32_t *residues, uint32_t n_sub, const uint8_t *order_in,
                                                    const int32_t *q_in, int32_t *samples)
{
    __shared__ __align__(16) CoefSmem cf;
    __shared__ double t[104];
    __shared__ __align__(16) int32_t buf[kFrame];
    const uint32_t sub = blockIdx.x;
    const int lane = lane_id();
    int order = order_in[sub];
    if (order > kMaxOrder)
        order = kMaxOrder;
    for (int i = lane; i < 104; i += 32)
        cf.q[i] = i < order ? q_in[(size_t)sub * kMaxOrder + i] : 0;
    for (int i = lane; i < kFrame; i += 32)
        buf[i] = residues[(size_t)sub * kFrame + i];
    __syncwarp();
    warp_coefficients(cf, t, order);
    warp_iir_synthesis(cf, order, buf, kFrame);
    for (int i = lane; i < kFrame; i += 32)
        samples[(size_t)sub * kFrame + i] = buf[i];
}

// Exhaustive device check of sample_to_x against IEEE division over |s| <= 65535.
__global__ void k_selftest_scaling(uint32_t *mismatches)
{
    const int s = (int)(blockIdx.x * blockDim.x + threadIdx.x) - 65535;
    if (s > 65535)
        return;
    if (__double_as_longlong(sample_to_x(s)) != __double_as_longlong(sample_to_x_div(s)))
        atomicAdd(mismatches, 1u);
}

// rice::RiceEncoder::process, one stream per (1-warp) CTA.
__global__ void __launch_bounds__(32) k_rice_encode(const int32_t *values, const uint32_t *counts, uint32_t stride,
                                                    uint32_t *k_out, uint32_t *n_words_out, uint32_t *words,
                                                    uint32_t words_stride, int32_t *status)
{
    const uint32_t st = blockIdx.x;
    const int32_t *v = values + (size_t)st * stride;
    const int n = (int)counts[st];
    auto at = [v](int i) { return v[i]; };
    RiceChoice c = warp_rice_choose(at, n);
    if (lane_id() == 0) {
        k_out[st] = c.k;
        n_words_out[st] = c.words;
    }
    if (c.words > words_stride) {
        if (lane_id() == 0)
            raise_status(status, SELAB200_ERR_CAPACITY);
        return;
    }
    warp_rice_pack(at, n, c.k, c.words, words + (size_t)st * words_stride);
}

// rice::RiceDecoder::process, one stream per lane.
__global__ void __launch_bounds__(128) k_rice_decode_streams(const uint32_t *words, const uint32_t *n_words,
                                                             uint32_t words_stride, const uint32_t *k,
                                                             const uint32_t *counts, uint32_t n_streams,
                                                             int32_t *out, uint32_t out_stride, int32_t *status)
{
    const uint32_t st = blockIdx.x * blockDim.x + threadIdx.x;
    if (st >= n_streams)
        return;
    if (k[st] >= 32 || counts[st] > out_stride || n_words[st] > words_stride) {
        raise_status(status, SELAB200_ERR_BITSTREAM);
        return;
    }
    lane_rice_decode(words + (size_t)st * words_stride, n_words[st], k[st], counts[st],
                     out + (size_t)st * out_stride);
}

} // namespace selab200

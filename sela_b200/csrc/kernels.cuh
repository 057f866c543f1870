// kernels.cuh -- the __global__ kernels of the SELA hot path (sm_100a).
//
//   encode   k_encode_units<STEREO> (warp per analysis unit: PCM -> ... -> Rice pack into a private slot)
//            k_encode_sizes + k_encode_scan (stereo decision, prefix sum, descriptors)
//            k_encode_gather / k_encode_gather_container (slot -> word arena / .sela byte stream)
//   decode   k_container_unpack (.sela bytes -> word arena)
//            k_decode_class_counts + k_decode_classify (subframes grouped by predictor-order class)
//            k_rice_decode<RING,BATCH> (lane per stream: reflection streams, flagged residue streams)
//            k_rice_split_index / k_rice_decode_vc (rice_vs.cuh: residue streams, lane per part of a stream)
//            k_synthesise_quad (+ k_diff_fixup; k_synthesise for frames the batch kernel declines)
//   stage-level entry points   k_lpc_residues, k_lpc_samples, k_rice_encode, k_rice_decode_streams
#pragma once

#include "lpc.cuh"
#include "rice.cuh"
#include "rice_vs.cuh"

namespace selab200 {

// ------------------------------------------------------------------ encode --
//
// Four launches, no inter-CTA dependency inside any of them:
//   k_encode_units   one WARP per analysis unit (a channel, or for stereo the three
//                    candidates ch0 / ch1 / ch0-ch1 of a frame): PCM -> autocorrelation ->
//                    Schur -> quantise -> step-up -> FIR -> Rice parameter search -> Rice
//                    pack into the unit's private slot of a scratch arena + a 32-byte
//                    unit record.  No barriers, no atomics between units.
//   k_encode_sizes,  1024 subframes per CTA: stereo decision (difference wins iff strictly fewer
//   k_encode_scan    words, src/frame/frame_encoder.cpp:63-72), exclusive prefix sum of the chosen
//                    sizes in file order, descriptors, total (see there).
//   k_encode_gather  one warp per emitted subframe: slot -> final arena offset.

constexpr uint32_t kSlotWords = 1600;     // per-unit scratch slot (refl words first, then residue words)
constexpr uint32_t kSlotReflWords = 32;   // 100 coefficients * (8 + 1) bits <= 29 words

struct __align__(16) UnitRecord {         // 32 bytes
    uint32_t order;
    uint32_t refl_k, refl_words;
    uint32_t res_k, res_words;
    uint32_t flags;                       // 1 = too large for the slot / the uint16 fields
    uint32_t pad[2];
};

struct EncodeParams {
    const int16_t *pcm;            // [n_frames][2048][channels] interleaved
    uint32_t n_frames, channels;
    selab200_subframe_desc *descs; // [n_frames*channels]
    uint32_t *words;
    unsigned long long capacity;   // words
    unsigned long long *words_used;
    int32_t *status;
    UnitRecord *units;             // workspace [n_units]
    uint32_t *slots;               // workspace [n_units][kSlotWords]
    int32_t *residues;             // workspace [n_units][2048]: FIR output, re-read by the Rice stages (L2-resident)
};

__host__ __device__ inline uint32_t encode_units(uint32_t n_frames, uint32_t channels)
{
    return channels == 2 ? n_frames * 3u : n_frames * channels;
}

template <bool STEREO>
__global__ void __launch_bounds__(32) k_encode_units(EncodeParams p)
{
    constexpr int kRow = kHistoryPad + kFrame;
    constexpr int kLoWords = (kHistoryPad + kFrame) / 32; // parity bits of a difference signal
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int16_t *s16 = reinterpret_cast<int16_t *>(smem_raw);                    // [pad + 2048]
    uint32_t *lo_bits = reinterpret_cast<uint32_t *>(smem_raw + kRow * 2);    // [pad/32 + 64] (stereo only)
    constexpr size_t kSigBytes = kRow * 2 + (STEREO ? kLoWords * 4 : 0);
    AnalysisScratch &scratch = *reinterpret_cast<AnalysisScratch *>(smem_raw + kSigBytes);
    // The predictor lives in the part of the analysis scratch that is dead once the Schur recursion has taken
    // the autocorrelation into registers (the tail of the ring and ac[]): 7.7 instead of 9 KB per unit, 29
    // instead of 23 units per SM.
    static_assert(kCoefAlias + sizeof(CoefSmem) <= sizeof(AnalysisScratch) && kCoefAlias % 16 == 0, "predictor alias");
    CoefSmem &cf = *reinterpret_cast<CoefSmem *>(smem_raw + kSigBytes + kCoefAlias);

    const int lane = lane_id();
    const uint32_t unit = blockIdx.x;
    const uint32_t frame = STEREO ? unit / 3 : unit / p.channels;
    const uint32_t role = STEREO ? unit % 3 : unit % p.channels; // stereo: 0 ch0, 1 ch1, 2 ch0-ch1

    // ---- stage the signal (de-interleave to one planar int16 row, zero history in front) ----
    const int16_t *src = p.pcm + (size_t)frame * kFrame * p.channels;
    for (int j = lane; j < kHistoryPad / 2; j += 32)
        reinterpret_cast<uint32_t *>(s16)[j] = 0;
    int16_t *row = s16 + kHistoryPad;
    Signal sig;
    sig.a = row;
    sig.lo = nullptr;
    if (STEREO) {
        const uint4 *src128 = reinterpret_cast<const uint4 *>(src); // 4 stereo sample pairs per load
        if (role < 2) {
            const uint32_t sel = role ? 0x7632 : 0x5410;
            for (int j = lane; j < kFrame / 4; j += 32) {
                const uint4 v = src128[j];
                reinterpret_cast<uint2 *>(row)[j] = make_uint2(__byte_perm(v.x, v.y, sel), __byte_perm(v.z, v.w, sel));
            }
        } else {
            // difference d = ch0 - ch1 (17 bits): d >> 1 into the row, d & 1 into the bit array
            if (lane < kHistoryPad / 32)
                lo_bits[lane] = 0;
            uint32_t *lo = lo_bits + kHistoryPad / 32;
            for (int it = 0; it < kFrame / 128; it++) {
                const uint4 v = src128[it * 32 + lane];
                const uint32_t pr[4] = {v.x, v.y, v.z, v.w};
                int d[4];
#pragma unroll
                for (int e = 0; e < 4; e++)
                    d[e] = ((int)(pr[e] << 16) >> 16) - ((int)pr[e] >> 16);
                const uint32_t h01 = ((uint32_t)(d[0] >> 1) & 0xffffu) | ((uint32_t)(d[1] >> 1) << 16);
                const uint32_t h23 = ((uint32_t)(d[2] >> 1) & 0xffffu) | ((uint32_t)(d[3] >> 1) << 16);
                reinterpret_cast<uint2 *>(row)[it * 32 + lane] = make_uint2(h01, h23);
                uint32_t nib = (d[0] & 1) | ((d[1] & 1) << 1) | ((d[2] & 1) << 2) | ((d[3] & 1) << 3);
                nib <<= 4 * (lane & 7);
                nib |= __shfl_xor_sync(kFull, nib, 1);
                nib |= __shfl_xor_sync(kFull, nib, 2);
                nib |= __shfl_xor_sync(kFull, nib, 4);
                if ((lane & 7) == 0)
                    lo[it * 4 + (lane >> 3)] = nib;
            }
            sig.lo = lo;
        }
    } else {
        for (int j = lane; j < kFrame; j += 32)
            row[j] = src[(size_t)j * p.channels + role];
    }
    __syncwarp();

    // ---- analysis ----
    int32_t *res = p.residues + (size_t)unit * kFrame;
    warp_autocorrelation(sig, scratch);
    warp_schur(scratch);
    const int order = warp_order_and_quantise(scratch, cf);
    warp_coefficients(cf, scratch.t(), order);
    warp_fir_residual(sig, cf, order, res);

    // ---- Rice: parameter search, then pack into this unit's slot ----
    const RiceChoice cq = warp_rice_choose(cf.q, order);
    const RiceChoice cr = warp_rice_choose(res, kFrame);
    const bool too_large = cq.words > kSlotReflWords || cr.words > kSlotWords - kSlotReflWords;
    uint32_t *slot = p.slots + (size_t)unit * kSlotWords;
    if (!too_large) {
        warp_rice_pack(cf.q, order, cq.k, cq.words, slot);
        warp_rice_pack(res, kFrame, cr.k, cr.words, slot + kSlotReflWords);
    }
    // The residue row is dead now.  It only ever lived in L2 (written and re-read by this warp
    // within microseconds); tell L2 to drop the dirty lines instead of writing 8 KB per unit back
    // to HBM.
    __syncwarp();
    for (int l = lane; l < kFrame * 4 / 128; l += 32)
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(res + l * 32) : "memory");
    if (lane == 0) {
        UnitRecord u;
        u.order = order;
        u.refl_k = cq.k;
        u.refl_words = cq.words;
        u.res_k = cr.k;
        u.res_words = cr.words;
        u.flags = too_large ? 1u : 0u;
        u.pad[0] = u.pad[1] = 0;
        p.units[unit] = u;
    }
}

template <bool STEREO>
constexpr size_t encode_smem_bytes()
{
    return (size_t)(kHistoryPad + kFrame) * 2 + (STEREO ? (kHistoryPad + kFrame) / 8 : 0) + sizeof(AnalysisScratch);
}

// Warp copy of n words, eight loads in flight per lane before the first store (a load-store-load-store loop leaves
// ONE: the in-order issue stops at every store until its load has landed).  The source is read once and dead after.
__device__ __forceinline__ void warp_copy_words(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, uint32_t n, int lane)
{
    for (uint32_t w = lane; w < n; w += 256) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
            v[j] = w + 32 * j < n ? __ldcs(src + w + 32 * j) : 0u;
#pragma unroll
        for (int j = 0; j < 8; j++)
            if (w + 32 * j < n)
                dst[w + 32 * j] = v[j];
    }
}

// Which unit is emitted for output subframe (frame, channel), and as what.
struct Emit {
    uint32_t unit, type, parent;
};
__device__ __forceinline__ Emit choose_unit(const UnitRecord *units, uint32_t channels, uint32_t sub)
{
    Emit e;
    if (channels != 2) {
        e.unit = sub;
        e.type = 0;
        e.parent = sub % channels;
        return e;
    }
    const uint32_t frame = sub >> 1;
    if ((sub & 1) == 0) {
        e.unit = 3 * frame;
        e.type = 0;
        e.parent = 0;
        return e;
    }
    const UnitRecord a = units[3 * frame + 1], d = units[3 * frame + 2];
    const bool diff_wins = (unsigned long long)d.refl_words + d.res_words <
                           (unsigned long long)a.refl_words + a.res_words; // strictly smaller
    e.unit = 3 * frame + (diff_wins ? 2 : 1);
    e.type = diff_wins ? 1 : 0;
    e.parent = diff_wins ? 0 : 1;
    return e;
}

// The unit emitted for subframe `sub` together with its record (for a stereo side channel the winner is one of the
// two records the decision reads anyway).
__device__ __forceinline__ Emit choose_record(const UnitRecord *units, uint32_t channels, uint32_t sub, UnitRecord &u)
{
    Emit e;
    if (channels != 2 || (sub & 1) == 0) {
        e.unit = channels != 2 ? sub : 3 * (sub >> 1);
        e.type = 0;
        e.parent = channels != 2 ? sub % channels : 0;
        u = units[e.unit];
        return e;
    }
    const uint32_t frame = sub >> 1;
    const UnitRecord a = units[3 * frame + 1], d = units[3 * frame + 2];
    const bool diff_wins = (unsigned long long)d.refl_words + d.res_words <
                           (unsigned long long)a.refl_words + a.res_words; // strictly smaller
    e.unit = 3 * frame + (diff_wins ? 2 : 1);
    e.type = diff_wins ? 1 : 0;
    e.parent = diff_wins ? 0 : 1;
    u = diff_wins ? d : a;
    return e;
}

// Block-wide inclusive scan of one 64-bit value per thread (1024 threads); warp_tot[31] is the block total afterwards.
__device__ __forceinline__ unsigned long long block_scan_inclusive_u64(unsigned long long v, unsigned long long *warp_tot)
{
    const int lane = lane_id(), warp = warp_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(kFull, v, o);
        if (lane >= o)
            v += t;
    }
    if (lane == 31)
        warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(kFull, w, o);
            if (lane >= o)
                w += t;
        }
        warp_tot[lane] = w; // inclusive over warps
    }
    __syncthreads();
    return v + (warp ? warp_tot[warp - 1] : 0ull);
}

// The scan of the emitted sizes (stereo decision, exclusive prefix sum in file order, descriptors, fill level) in two
// launches of 1024 subframes per CTA, coalesced record loads in both:
//   k_encode_sizes   CTA b: total words of its 1024 subframes -> tmp[1 + b]           (independent of earlier chunks)
//   k_encode_scan    CTA b: fill level so far (*words_used, left by the previous chunk of a pipelined call) + the
//                    totals of the CTAs before it + a block scan -> descriptors; the last CTA leaves the new fill level
//                    in tmp[0] (the host side copies it to *words_used: other CTAs may still be reading the old one).
// tmp is the head of the residue workspace, dead once k_encode_units has finished.
constexpr uint32_t kScanTile = 1024;

__global__ void __launch_bounds__(kScanTile) k_encode_sizes(EncodeParams p)
{
    __shared__ unsigned long long warp_tot[32];
    const uint32_t n_sub = p.n_frames * p.channels;
    const uint32_t sub = blockIdx.x * kScanTile + threadIdx.x;
    unsigned long long size = 0;
    if (sub < n_sub) {
        UnitRecord u;
        choose_record(p.units, p.channels, sub, u);
        size = (unsigned long long)u.refl_words + u.res_words;
    }
    block_scan_inclusive_u64(size, warp_tot);
    if (threadIdx.x == 0)
        reinterpret_cast<unsigned long long *>(p.residues)[1 + blockIdx.x] = warp_tot[31];
}

__global__ void __launch_bounds__(kScanTile) k_encode_scan(EncodeParams p)
{
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long before;
    unsigned long long *tmp = reinterpret_cast<unsigned long long *>(p.residues);
    const uint32_t n_sub = p.n_frames * p.channels;
    const uint32_t sub = blockIdx.x * kScanTile + threadIdx.x;
    const int lane = lane_id(), warp = warp_id();
    if (warp == 0) { // words in front of this CTA
        unsigned long long s = lane == 0 ? *p.words_used : 0ull;
        for (uint32_t b = lane; b < blockIdx.x; b += 32)
            s += tmp[1 + b];
        s = warp_sum_u64(s);
        if (lane == 0)
            before = s;
    }
    Emit e;
    UnitRecord u;
    unsigned long long size = 0;
    bool bad = false;
    if (sub < n_sub) {
        e = choose_record(p.units, p.channels, sub, u);
        size = (unsigned long long)u.refl_words + u.res_words;
        bad = u.flags != 0 || u.refl_words > 0xffffu || u.res_words > 0xffffu;
    }
    const unsigned long long incl = block_scan_inclusive_u64(size, warp_tot); // its barriers also publish `before`
    if (sub < n_sub) {
        const unsigned long long off = before + incl - size;
        selab200_subframe_desc v;
        v.channel = (uint8_t)(sub % p.channels);
        v.subframe_type = (uint8_t)e.type;
        v.parent_channel = (uint8_t)e.parent;
        v.refl_rice_param = (uint8_t)u.refl_k;
        v.refl_words = (uint16_t)u.refl_words;
        v.lpc_order = (uint8_t)u.order;
        v.res_rice_param = (uint8_t)u.res_k;
        v.res_words = (uint16_t)u.res_words;
        v.samples = (uint16_t)kFrame;
        v.reserved = 0;
        v.refl_offset = off;
        v.res_offset = off + u.refl_words;
        p.descs[sub] = v;
    }
    if (bad)
        raise_status(p.status, SELAB200_ERR_RANGE); // a stream the uint16 word counts cannot describe
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        const unsigned long long fill = before + warp_tot[31];
        tmp[0] = fill;
        if (fill > p.capacity)
            raise_status(p.status, SELAB200_ERR_CAPACITY);
    }
}

// One warp per emitted subframe; 8 warps per CTA.
__global__ void __launch_bounds__(256) k_encode_gather(EncodeParams p)
{
    const uint32_t sub = blockIdx.x * 8 + warp_id();
    const uint32_t n_sub = p.n_frames * p.channels;
    if (sub >= n_sub || *reinterpret_cast<volatile int32_t *>(p.status) != 0)
        return;
    const int lane = lane_id();
    const Emit e = choose_unit(p.units, p.channels, sub);
    const selab200_subframe_desc d = p.descs[sub];
    const uint32_t *slot = p.slots + (size_t)e.unit * kSlotWords;
    warp_copy_words(p.words + d.refl_offset, slot, d.refl_words, lane);
    warp_copy_words(p.words + d.res_offset, slot + kSlotReflWords, d.res_words, lane);
}

// --------------------------------------------------------------- container --
//
// The .sela container (src/file/sela_file.cpp:105-137) is byte-packed: a 15-byte file header,
// then per frame the sync word 0xAA55FF00 (little endian: 00 FF 55 AA) and per subframe
//   channel, type, parent, reflK (u8 each), reflInts (u16), order (u8), refl words,
//   resK (u8), resInts (u16), samples (u16), residue words.
// With the word arena laid out in file order (refl words then residue words, subframe after
// subframe), the byte position of everything follows from the subframe's global index g, its
// frame f = g / channels and the arena offset W of its first word:
//   header of g at  15 + 4*(f+1) + 12*g + 4*W.
// Word arrays therefore sit at arbitrary byte alignments; the kernels below move them with
// aligned 32-bit accesses and a funnel shift, byte stores only at the two ragged ends.
constexpr unsigned long long kContainerHeaderBytes = 15;
constexpr uint32_t kSubframeHeaderBytes = 12;

__device__ __forceinline__ unsigned long long container_subframe_byte(unsigned long long g, uint32_t channels,
                                                                       unsigned long long first_word)
{
    return kContainerHeaderBytes + 4ull * (g / channels + 1) + (unsigned long long)kSubframeHeaderBytes * g +
           4ull * first_word;
}

// n words from src (aligned) to byte address out + D (any alignment).
__device__ __forceinline__ void put_words_at_byte(uint8_t *out, unsigned long long D, const uint32_t *src, uint32_t n)
{
    const int lane = lane_id();
    if (n == 0)
        return;
    const uint32_t s = (uint32_t)(D & 3);
    uint32_t *dst = reinterpret_cast<uint32_t *>(out + (D - s));
    if (s == 0) {
        warp_copy_words(dst, src, n, lane);
        return;
    }
    // aligned word t = high bytes of src[t-1], low bytes of src[t]; four words per lane loaded before the first store
    for (uint32_t t0 = 1 + lane; t0 < n; t0 += 128) {
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t t = t0 + 32 * j;
            v[j] = t < n ? __funnelshift_l(src[t - 1], src[t], 8 * s) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (t0 + 32 * j < n)
                dst[t0 + 32 * j] = v[j];
    }
    if (lane < (int)(4 - s))
        out[D + lane] = (uint8_t)(src[0] >> (8 * lane));
    else if (lane >= 4 && lane < (int)(4 + s))
        out[D - s + 4ull * n + (lane - 4)] = (uint8_t)(src[n - 1] >> (8 * (4 - s + (lane - 4))));
}

// n words from byte address in + D (any alignment) to dst (aligned).  Reads the aligned word
// that holds the last byte, i.e. at most 3 bytes past the array (the buffer is padded).
__device__ __forceinline__ void get_words_at_byte(const uint8_t *in, unsigned long long D, uint32_t *dst, uint32_t n)
{
    const int lane = lane_id();
    const uint32_t s = (uint32_t)(D & 3);
    const uint32_t *src = reinterpret_cast<const uint32_t *>(in + (D - s));
    for (uint32_t t0 = lane; t0 < n; t0 += 128) { // four words per lane loaded before the first store
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t t = t0 + 32 * j;
            v[j] = t < n ? (s ? __funnelshift_r(src[t], src[t + 1], 8 * s) : src[t]) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (t0 + 32 * j < n)
                dst[t0 + 32 * j] = v[j];
    }
}

// Encoder output straight into the container: k_encode_gather with byte-packed destinations
// and the headers written in place.  One warp per emitted subframe; `sub_base` is the global
// index of this chunk's first subframe.
__global__ void __launch_bounds__(256) k_encode_gather_container(EncodeParams p, uint8_t *container,
                                                                 unsigned long long sub_base)
{
    const uint32_t sub = blockIdx.x * 8 + warp_id();
    const uint32_t n_sub = p.n_frames * p.channels;
    if (sub >= n_sub || *reinterpret_cast<volatile int32_t *>(p.status) != 0)
        return;
    const int lane = lane_id();
    const Emit e = choose_unit(p.units, p.channels, sub);
    const selab200_subframe_desc d = p.descs[sub];
    const uint32_t *slot = p.slots + (size_t)e.unit * kSlotWords;
    const unsigned long long g = sub_base + sub;
    const unsigned long long at = container_subframe_byte(g, p.channels, d.refl_offset);
    if (g % p.channels == 0 && lane >= 12 && lane < 16) {
        const uint32_t sync = 0xAA55FF00u;
        container[at - 4 + (lane - 12)] = (uint8_t)(sync >> (8 * (lane - 12)));
    }
    const unsigned long long at2 = at + 7 + 4ull * d.refl_words;
    if (lane < 12) {
        uint8_t v;
        switch (lane) {
        case 0: v = d.channel; break;
        case 1: v = d.subframe_type; break;
        case 2: v = d.parent_channel; break;
        case 3: v = d.refl_rice_param; break;
        case 4: v = (uint8_t)d.refl_words; break;
        case 5: v = (uint8_t)(d.refl_words >> 8); break;
        case 6: v = d.lpc_order; break;
        case 7: v = d.res_rice_param; break;
        case 8: v = (uint8_t)d.res_words; break;
        case 9: v = (uint8_t)(d.res_words >> 8); break;
        case 10: v = (uint8_t)d.samples; break;
        default: v = (uint8_t)(d.samples >> 8); break;
        }
        container[lane < 7 ? at + lane : at2 + (lane - 7)] = v;
    }
    put_words_at_byte(container, at + 7, slot, d.refl_words);
    put_words_at_byte(container, at2 + 5, slot + kSlotReflWords, d.res_words);
}

// Decoder input straight from the container: realign the word arrays of each subframe into the
// (16-byte aligned) arena the Rice decoder reads.  descs[] come from the host's header walk and
// carry arena offsets in file order.  One warp per subframe.
__global__ void __launch_bounds__(256) k_container_unpack(const uint8_t *container, const selab200_subframe_desc *descs,
                                                          uint32_t n_sub, uint32_t channels, unsigned long long sub_base,
                                                          uint32_t *arena)
{
    const uint32_t sub = blockIdx.x * 8 + warp_id();
    if (sub >= n_sub)
        return;
    const selab200_subframe_desc d = descs[sub];
    const unsigned long long at = container_subframe_byte(sub_base + sub, channels, d.refl_offset);
    get_words_at_byte(container, at + 7, arena + d.refl_offset, d.refl_words);
    get_words_at_byte(container, at + 7 + 4ull * d.refl_words + 5, arena + d.res_offset, d.res_words);
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
    uint32_t *order_index; // [n_sub + 16]: subframe ids grouped by predictor-order class (k_decode_classify)
    int fallback_only;
    const uint32_t *rice_flags; // residue pass of k_rice_decode: when set, only streams with a non-zero flag (rice_vs.cuh)
};

// K5: one lane per stream, 2 warps per CTA.  which = 0: reflection streams (-> ws_q), 1: residue
// streams (-> ws_res).
constexpr int kRiceWarps = 2;
constexpr int kRiceRing = 64, kRiceBatch = 8;
template <int RING, int BATCH>
__global__ void __launch_bounds__(32 * kRiceWarps) k_rice_decode(DecodeParams p, int which)
{
    __shared__ uint32_t ring[kRiceWarps][(RING + 1) * 32];
    const uint32_t sub = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_sub = p.n_frames * p.channels;
    RiceLaneStream st;
    st.src = p.words;
    st.n_words = 0;
    st.k = 0;
    st.count = 0;
    st.out = nullptr;
    if (sub < n_sub) {
        const selab200_subframe_desc d = p.descs[sub];
        if (!desc_ok(d, p.channels, p.n_words)) {
            raise_status(p.status, SELAB200_ERR_BITSTREAM);
        } else if (which == 0) {
            st.src = p.words + d.refl_offset;
            st.n_words = d.refl_words;
            st.k = d.refl_rice_param;
            st.count = d.lpc_order;
            st.out = p.ws_q + (size_t)sub * 128;
        } else if (!p.rice_flags || p.rice_flags[sub]) {
            st.src = p.words + d.res_offset;
            st.n_words = d.res_words;
            st.k = d.res_rice_param;
            st.count = d.samples;
            st.out = p.ws_res + (size_t)sub * kFrame;
        }
    }
    if (which == 1 && p.rice_flags && __ballot_sync(kFull, st.count != 0) == 0)
        return; // nothing flagged in this warp (the usual case)
    if (!warp_rice_decode32<RING, BATCH>(ring[warp_id()], st))
        raise_status(p.status, SELAB200_ERR_BITSTREAM);
}

__device__ __forceinline__ bool frame_needs_general(uint32_t ch, unsigned type_mask) { return ch != 2 && type_mask != 0; }

// K6, general form: CTA per frame; each warp synthesises TWO subframes (one per half-warp) into
// shared-memory planes, so any parent/child arrangement inside the frame can be resolved.
__global__ void k_synthesise(DecodeParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t ch = p.channels;
    int32_t *planes = reinterpret_cast<int32_t *>(smem_raw);                  // [ch][2048], by channel
    IirSmem *iir_all = reinterpret_cast<IirSmem *>(planes + (size_t)ch * kFrame);
    const uint32_t n_warps = (ch + 1) / 2;
    double *t_all = reinterpret_cast<double *>(iir_all + ch);                 // [warps][104]
    CoefSmem *coef_all = reinterpret_cast<CoefSmem *>(t_all + (size_t)n_warps * 104);
    int *meta = reinterpret_cast<int *>(coef_all + ch); // [ch] type, [ch] parent, [1] valid

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
    if (p.fallback_only) {
        // launched behind k_synthesise_quad: only frames it declines (difference coding with a
        // channel count other than 2) are left to do
        unsigned type_mask = 0;
        for (uint32_t c = 0; c < ch; c++)
            type_mask |= (unsigned)(meta[c] == 1) << c;
        if (!valid || !frame_needs_general(ch, type_mask))
            return;
    }
    if (!valid) {
        for (uint32_t e = threadIdx.x; e < kFrame * ch; e += blockDim.x)
            out[e] = 0;
        return;
    }

    {
        // positions 2*warp (half A) and 2*warp+1 (half B); an odd channel count leaves the
        // last half idle: it shadows half A without storing
        const uint32_t pos_a = 2 * warp, pos_b = 2 * warp + 1;
        const bool has_b = pos_b < ch;
        for (uint32_t h = 0; h < (has_b ? 2u : 1u); h++) {
            const uint32_t pos = h ? pos_b : pos_a;
            const uint32_t sub = frame * ch + pos;
            const selab200_subframe_desc d = fd[pos];
            CoefSmem &cf = coef_all[pos];
            int32_t *buf = planes + (size_t)d.channel * kFrame;
            const int order = d.lpc_order;
            for (int i = lane; i < 104; i += 32)
                cf.q[i] = i < order ? p.ws_q[(size_t)sub * 128 + i] : 0;
            const int4 *r4 = reinterpret_cast<const int4 *>(p.ws_res + (size_t)sub * kFrame);
            for (int i = lane; i < kFrame / 4; i += 32)
                reinterpret_cast<int4 *>(buf)[i] = r4[i];
            __syncwarp();
            // order 0 behaves like order 1 with a zero predictor (linear_predictor.cpp:19-22)
            warp_coefficients(cf, t_all + warp * 104, order);
            warp_iir_prepare(cf, iir_all[pos], order);
        }
        const bool upper = lane >= 16;
        const uint32_t my_pos = (upper && has_b) ? pos_b : pos_a;
        const selab200_subframe_desc d = fd[my_pos];
        warp_iir_synthesis_pair(coef_all[my_pos], iir_all[my_pos], d.lpc_order,
                                planes + (size_t)d.channel * kFrame, !upper || has_b, kFrame);
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

// Predictor-order classes of the batch synthesis kernel: taps per lane (8 lanes per subframe)
// 4 / 8 / 16, i.e. orders up to 28 / 56 / 112 (the last lane of a quarter must only hold taps
// beyond the order).  A warp runs all four of its subframes with the taps of the largest order
// among them, so subframes are first grouped by class: on the BASELINE synthetic (orders spread
// evenly over 17..100) that removes about a quarter of the multiplies.
__device__ __forceinline__ int order_class(int order) { return order <= 28 ? 0 : order <= 56 ? 1 : 2; }

// Stable counting sort of the subframes by class into p.order_index; every class segment
// starts on a warp boundary (4 subframes), widest class first; gaps hold 0xffffffff (pre-set by the host side).
// Two launches of 1024 subframes per CTA (coalesced descriptor loads in both):
//   k_decode_class_counts  CTA b: how many of its subframes fall in each class -> tmp[b] (classes 0|1 and 2|3 packed
//                          as the 32-bit halves of two words)
//   k_decode_classify      CTA b: totals of all CTAs (-> where each class segment starts; widest class first) + counts
//                          of the CTAs before it + a block scan -> order_index.  Ascending subframe order inside a
//                          class is kept (stable), so the two subframes of a stereo frame stay neighbours.
// tmp is the head of the Rice decoder's scratch, which is not in use yet.
struct ClassCounts {
    unsigned long long c01, c23;
};
__device__ __forceinline__ ClassCounts class_one(int c)
{
    ClassCounts v;
    const unsigned long long one = 1ull << (32 * (c & 1));
    v.c01 = (c >> 1) ? 0ull : one;
    v.c23 = (c >> 1) ? one : 0ull;
    return v;
}
__device__ __forceinline__ uint32_t class_get(const ClassCounts &v, int c)
{
    return (uint32_t)(((c >> 1) ? v.c23 : v.c01) >> (32 * (c & 1)));
}
__device__ __forceinline__ int subframe_class(const DecodeParams &p, uint32_t u)
{
    const int order = p.descs[u].lpc_order;
    return order_class(order > kMaxOrder ? 0 : order);
}

__global__ void __launch_bounds__(kScanTile) k_decode_class_counts(DecodeParams p, ClassCounts *tmp)
{
    __shared__ ClassCounts warp_cnt[32];
    const uint32_t n_units = p.n_frames * p.channels;
    const uint32_t u = blockIdx.x * kScanTile + threadIdx.x;
    ClassCounts v = {0ull, 0ull};
    if (u < n_units)
        v = class_one(subframe_class(p, u));
    v.c01 = warp_sum_u64(v.c01);
    v.c23 = warp_sum_u64(v.c23);
    if (lane_id() == 0)
        warp_cnt[warp_id()] = v;
    __syncthreads();
    if (warp_id() == 0) {
        v = warp_cnt[lane_id()];
        v.c01 = warp_sum_u64(v.c01);
        v.c23 = warp_sum_u64(v.c23);
        if (lane_id() == 0)
            tmp[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(kScanTile) k_decode_classify(DecodeParams p, const ClassCounts *tmp)
{
    __shared__ unsigned long long warp_tot[2][32];
    __shared__ ClassCounts total, before;
    const uint32_t n_units = p.n_frames * p.channels;
    const uint32_t u = blockIdx.x * kScanTile + threadIdx.x;
    const int lane = lane_id(), warp = warp_id();
    if (warp == 0) {
        ClassCounts all = {0ull, 0ull}, pre = {0ull, 0ull};
        for (uint32_t b = lane; b < gridDim.x; b += 32) {
            const ClassCounts v = tmp[b];
            all.c01 += v.c01, all.c23 += v.c23;
            if (b < blockIdx.x)
                pre.c01 += v.c01, pre.c23 += v.c23;
        }
        all.c01 = warp_sum_u64(all.c01), all.c23 = warp_sum_u64(all.c23);
        pre.c01 = warp_sum_u64(pre.c01), pre.c23 = warp_sum_u64(pre.c23);
        if (lane == 0)
            total = all, before = pre;
    }
    const bool have = u < n_units;
    const int c = have ? subframe_class(p, u) : 0;
    ClassCounts mine = {0ull, 0ull};
    if (have)
        mine = class_one(c);
    ClassCounts incl;
    incl.c01 = block_scan_inclusive_u64(mine.c01, warp_tot[0]); // the barriers in here also publish total / before
    incl.c23 = block_scan_inclusive_u64(mine.c23, warp_tot[1]);
    if (!have)
        return;
    // widest class first: a synthesis warp runs its four subframes start to finish (hundreds of microseconds), so
    // the longest-running warps must be scheduled first and the short ones left to fill the tail; every class
    // segment starts on a multiple of the four subframes a warp takes
    uint32_t base = 0;
    for (int k = 3; k > c; k--)
        base += (class_get(total, k) + 3u) / 4u * 4u;
    ClassCounts excl;
    excl.c01 = before.c01 + incl.c01 - mine.c01;
    excl.c23 = before.c23 + incl.c23 - mine.c23;
    p.order_index[base + class_get(excl, c)] = u;
}

// K6, batch form: one warp = four subframes of one predictor-order class (stereo: two whole
// frames, the pair of a frame in neighbouring quarters).  Handles every frame except those with
// difference-coded subframes and a channel count other than 2 (the reference's encoder never
// produces those; k_synthesise above picks them up).
struct QuadSmem {
    IirSmem ii[4];
    double t[104];
    CoefSmem cf[4];
    // staging rows [2][16] of the four quarters at word offsets 0, 48, 104, 152: the four writer lanes (one per
    // quarter, same row and column) then hit banks 0 / 16 / 8 / 24 apart, and the 8-byte reads of a half-warp
    // (two quarters) cover 32 distinct banks; a plain [4][2][16] puts all four writers on one bank
    int32_t stage[184];
};
__device__ __forceinline__ int quad_stage_offset(int q) { return 48 * q + 8 * (q >> 1); }

__global__ void __launch_bounds__(32) k_synthesise_quad(DecodeParams p)
{
    __shared__ __align__(16) QuadSmem sm;
    const uint32_t ch = p.channels;
    const uint32_t n_sub = p.n_frames * ch;
    const int lane = lane_id(), q = lane >> 3, hl = lane & 7;
    const uint32_t sub = p.order_index[blockIdx.x * 4 + q]; // grouped by order class; stereo pairs stay adjacent
    const bool exists = sub < n_sub;
    const uint32_t frame = exists ? sub / ch : 0, pos = exists ? sub % ch : 0;

    // ---- per-quarter validation of the whole frame (all eight lanes redundantly) ----
    bool frame_ok = exists;
    unsigned seen = 0, type_mask = 0;
    selab200_subframe_desc mine;
    memset(&mine, 0, sizeof mine);
    if (exists) {
        const selab200_subframe_desc *fd = p.descs + (size_t)frame * ch;
        for (uint32_t i = 0; i < ch; i++) {
            const selab200_subframe_desc d = fd[i];
            const bool ok = desc_ok(d, ch, p.n_words) && !((seen >> d.channel) & 1);
            frame_ok &= ok;
            if (ok) {
                seen |= 1u << d.channel;
                type_mask |= (unsigned)(d.subframe_type & 1) << d.channel;
            }
            if (i == pos)
                mine = d;
        }
        for (uint32_t i = 0; i < ch && frame_ok; i++) {
            const selab200_subframe_desc d = fd[i];
            if (d.subframe_type == 1 && ((type_mask >> d.parent_channel) & 1))
                frame_ok = false; // a parent must be an independent subframe
        }
        if (!frame_ok && hl == 0)
            raise_status(p.status, SELAB200_ERR_BITSTREAM);
    }
    const bool general = exists && frame_ok && frame_needs_general(ch, type_mask);
    const bool proc = exists && frame_ok && !general;
    const int order = proc ? mine.lpc_order : 0;

    // ---- predictors of the four subframes (warp-wide routines, one subframe at a time) ----
    for (int h = 0; h < 4; h++) {
        const int order_h = __shfl_sync(kFull, order, 8 * h);
        const uint32_t sub_h = __shfl_sync(kFull, sub, 8 * h);
        CoefSmem &cf = sm.cf[h];
        for (int i = lane; i < 104; i += 32)
            cf.q[i] = i < order_h ? p.ws_q[(size_t)sub_h * 128 + i] : 0;
        __syncwarp();
        warp_coefficients(cf, sm.t, order_h); // order 0 behaves like a zero predictor
        warp_iir_prepare(cf, sm.ii[h], order_h);
    }
    int order_max = order;
    order_max = max(order_max, __shfl_xor_sync(kFull, order_max, 8));
    order_max = max(order_max, __shfl_xor_sync(kFull, order_max, 16));

    // ---- recurrence + output ----
    QuadIo io;
    io.res = p.ws_res + (size_t)(exists ? sub : 0) * kFrame;
    io.stage = sm.stage + quad_stage_offset(q);
    const bool diff = proc && mine.subframe_type == 1;
    const uint32_t channel = mine.channel;
    int16_t *out = p.pcm_out + (size_t)frame * kFrame * ch;
    int32_t *row = p.ws_res + (size_t)(exists ? sub : 0) * kFrame;
    auto emit = [&](int B, int kx, int ky) {
        if (!proc)
            return;
        const size_t t = 16 * B + 2 * hl;
        if (diff) {
            // a difference signal goes back into its (already consumed) residue row;
            // k_diff_fixup turns it into parent - difference once the parent is complete
            *reinterpret_cast<int2 *>(row + t) = make_int2(kx, ky);
        } else {
            out[t * ch + channel] = (int16_t)(uint16_t)kx;
            out[(t + 1) * ch + channel] = (int16_t)(uint16_t)ky;
        }
    };
    switch (order_class(order_max)) {
    case 0: warp_iir_quad<4>(sm.cf[q], sm.ii[q], order, order_max, io, proc, emit); break;
    case 1: warp_iir_quad<8>(sm.cf[q], sm.ii[q], order, order_max, io, proc, emit); break;
    default: warp_iir_quad<16>(sm.cf[q], sm.ii[q], order, order_max, io, proc, emit); break;
    }

    if (exists && !frame_ok) { // malformed frame: silence
        for (int t = hl; t < kFrame; t += 8)
            out[(size_t)t * ch + (pos < ch ? pos : 0)] = 0;
    }
}

// Difference reconstruction for stereo (frame_decoder.cpp:40-69): after k_synthesise_quad the
// residue row of a difference-coded subframe holds the decoded difference; its channel is
// parent - difference.  One CTA per frame; frames without a difference subframe return at once.
__global__ void __launch_bounds__(128) k_diff_fixup(DecodeParams p)
{
    const uint32_t frame = blockIdx.x;
    const selab200_subframe_desc *fd = p.descs + (size_t)frame * 2;
    const selab200_subframe_desc d0 = fd[0], d1 = fd[1];
    // same acceptance rules as the synthesis kernel (which has already reported any violation)
    if (!desc_ok(d0, 2, p.n_words) || !desc_ok(d1, 2, p.n_words) || d0.channel == d1.channel)
        return;
    if ((d0.subframe_type | d1.subframe_type) == 0 || (d0.subframe_type & d1.subframe_type))
        return; // nothing to do / both dependent (rejected upstream)
    const uint32_t pos = d1.subframe_type ? 1 : 0;
    const selab200_subframe_desc d = pos ? d1 : d0;
    const int32_t *diff = p.ws_res + ((size_t)frame * 2 + pos) * kFrame;
    int16_t *out = p.pcm_out + (size_t)frame * kFrame * 2;
    for (int t = threadIdx.x; t < kFrame; t += blockDim.x)
        out[2 * t + d.channel] = (int16_t)(uint16_t)((int)out[2 * t + d.parent_channel] - diff[t]);
}

inline size_t synthesise_smem_bytes(uint32_t ch)
{
    return (size_t)ch * kFrame * 4 + ch * (sizeof(CoefSmem) + sizeof(IirSmem)) +
           ((ch + 1) / 2) * 104 * sizeof(double) + (2 * ch + 1) * sizeof(int);
}

// ------------------------------------------------------------ stage level --

// lpc::ResidueGenerator::process for one signal per (1-warp) CTA.
__global__ void __launch_bounds__(32) k_lpc_residues(const int32_t *samples, uint32_t n_sub, uint8_t *order_out,
                                                     int32_t *q_out, int32_t *residues)
{
    __shared__ __align__(16) AnalysisScratch scratch;
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
    warp_autocorrelation(sig, scratch);
    warp_schur(scratch);
    const int order = warp_order_and_quantise(scratch, cf);
    warp_coefficients(cf, scratch.t(), order);
    warp_fir_residual(sig, cf, order, residues + (size_t)sub * kFrame);
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
    __shared__ __align__(16) IirSmem ii;
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
    warp_iir_prepare(cf, ii, order);
    warp_iir_synthesis_pair(cf, ii, order, buf, lane < 16, kFrame); // upper half shadows, never stores
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
    RiceChoice c = warp_rice_choose(v, n);
    if (lane_id() == 0) {
        k_out[st] = c.k;
        n_words_out[st] = c.words;
    }
    if (c.words > words_stride) {
        if (lane_id() == 0)
            raise_status(status, SELAB200_ERR_CAPACITY);
        return;
    }
    warp_rice_pack(v, n, c.k, c.words, words + (size_t)st * words_stride);
}

// rice::RiceDecoder::process, one stream per lane.
__global__ void __launch_bounds__(32 * kRiceWarps) k_rice_decode_streams(
    const uint32_t *words, const uint32_t *n_words, uint32_t words_stride, const uint32_t *k,
    const uint32_t *counts, uint32_t n_streams, int32_t *out, uint32_t out_stride, int32_t *status)
{
    __shared__ uint32_t ring[kRiceWarps][(kRiceRing + 1) * 32];
    const uint32_t st_i = blockIdx.x * blockDim.x + threadIdx.x;
    RiceLaneStream st;
    st.src = words;
    st.n_words = 0;
    st.k = 0;
    st.count = 0;
    st.out = nullptr;
    if (st_i < n_streams) {
        if (k[st_i] >= 32 || counts[st_i] > out_stride || n_words[st_i] > words_stride) {
            raise_status(status, SELAB200_ERR_BITSTREAM);
        } else {
            st.src = words + (size_t)st_i * words_stride;
            st.n_words = n_words[st_i];
            st.k = k[st_i];
            st.count = counts[st_i];
            st.out = out + (size_t)st_i * out_stride;
        }
    }
    warp_rice_decode32<kRiceRing, kRiceBatch>(ring[warp_id()], st);
}

} // namespace selab200

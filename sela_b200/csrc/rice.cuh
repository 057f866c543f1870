// rice.cuh -- Golomb-Rice coding (kernels K4 / K5 of SURVEY.md 2).
//
// Bit layout (src/rice/rice_encoder.cpp:35-71): stream bit b lives in word b/32 at
// bit b%32 (LSB first); a symbol is u>>k ones, one zero, then the k low bits of u
// MSB FIRST; the last word is zero padded.  u is the zig-zag of the int32 input.
#pragma once

#include "common.cuh"

namespace selab200 {

// convertSignedToUnsigned (rice_encoder.cpp:12-18), 32-bit domain
__device__ __forceinline__ uint32_t zigzag(int32_t v)
{
    return ((uint32_t)v << 1) ^ (uint32_t)(v >> 31);
}
// convertUnsignedToSigned (src/rice/rice_decoder.cpp:46-52)
__device__ __forceinline__ int32_t unzigzag(uint32_t u)
{
    return (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
}

struct RiceChoice {
    uint32_t k;
    uint32_t bits;  // requiredBits (saturated at 0xffffffff)
    uint32_t words; // ceil(bits / 32)
};

// calculateOptimumRiceParam (rice_encoder.cpp:20-33): first arg-min over k = 0..19 of
// sum(u >> k) + n*(1 + k).  Val(i) returns the i-th int32 input, i < n <= 2048.
// Every lane returns the same result.
template <typename Val>
__device__ RiceChoice warp_rice_choose(const Val &val, int n)
{
    const int lane = lane_id();
    unsigned long long sums[kMaxRice];
#pragma unroll
    for (int k = 0; k < kMaxRice; k++)
        sums[k] = 0;
    for (int i = lane; i < n; i += 32) {
        const uint32_t u = zigzag(val(i));
#pragma unroll
        for (int k = 0; k < kMaxRice; k++)
            sums[k] += u >> k;
    }
    unsigned long long best = ~0ull;
    uint32_t best_k = 0;
#pragma unroll
    for (int k = 0; k < kMaxRice; k++) {
        unsigned long long total = warp_sum_u64(sums[k]) + (unsigned long long)n * (1 + k);
        if (total < best) {
            best = total;
            best_k = k;
        }
    }
    RiceChoice c;
    c.k = best_k;
    c.bits = best > 0xffffffffull ? 0xffffffffu : (uint32_t)best;
    unsigned long long w = (best + 31) >> 5;
    c.words = w > 0xffffffffull ? 0xffffffffu : (uint32_t)w;
    return c;
}

// generateEncodedBits + writeInts (rice_encoder.cpp:35-71).  Lane l codes the
// contiguous values [l*per, (l+1)*per); a warp prefix sum of the code lengths gives
// each lane its first bit.  Words wholly inside a lane's range are plain stores;
// the (at most two) words it shares with a neighbour are OR-ed into the
// pre-zeroed destination.  dst: `words` uint32 in global memory.
template <typename Val>
__device__ void warp_rice_pack(const Val &val, int n, uint32_t k, uint32_t words, uint32_t *dst)
{
    const int lane = lane_id();
    for (uint32_t w = lane; w < words; w += 32)
        dst[w] = 0;
    __syncwarp();

    const int per = (n + 31) >> 5;
    const int lo = lane * per < n ? lane * per : n;
    const int hi = lo + per < n ? lo + per : n;
    uint32_t my_bits = 0;
    for (int i = lo; i < hi; i++)
        my_bits += (zigzag(val(i)) >> k) + 1 + k;
    const uint32_t end = warp_scan_inclusive_u32(my_bits);
    const uint32_t start = end - my_bits;

    // 64-bit staging: bits [0, fill) of `stage` are pending for word index `widx`
    unsigned long long stage = 0;
    uint32_t fill = start & 31;
    uint32_t widx = start >> 5;
    auto flush = [&](uint32_t value) {
        const uint32_t b0 = widx << 5;
        if (b0 >= start && b0 + 32 <= end)
            dst[widx] = value;
        else if (value)
            atomicOr(&dst[widx], value);
        widx++;
    };
    for (int i = lo; i < hi; i++) {
        const uint32_t u = zigzag(val(i));
        uint32_t ones = u >> k;
        while (ones >= 32) { // long unary run: 32 ones at a time
            stage |= 0xffffffffull << fill;
            flush((uint32_t)stage);
            stage >>= 32;
            ones -= 32;
        }
        // remaining ones (< 32), the zero, then k payload bits MSB first == the
        // bit-reversed low k bits placed LSB first after the zero
        unsigned long long code = (1ull << ones) - 1;
        const uint32_t payload = k ? (__brev(u) >> (32 - k)) : 0u; // brev of low k bits
        code |= (unsigned long long)payload << (ones + 1);
        const uint32_t len = ones + 1 + k; // <= 31 + 1 + 19
        stage |= code << fill;             // fill < 32, len <= 51: may exceed 64 bits
        const uint32_t total = fill + len;
        if (total >= 32) {
            flush((uint32_t)stage);
            if (total >= 64) {
                flush((uint32_t)(stage >> 32));
                // bits of `code` that did not fit: code >> (64 - fill)
                stage = fill ? (code >> (64 - fill)) : 0ull;
                fill = total - 64;
            } else {
                stage >>= 32;
                fill = total - 32;
            }
        } else {
            fill = total;
        }
    }
    if (fill && hi > lo)
        flush((uint32_t)stage);
    __syncwarp();
}

// rice::RiceDecoder (src/rice/rice_decoder.cpp:11-52), ONE LANE PER STREAM: the parse
// is inherently sequential, so the parallelism is across streams (subframes) and
// every lane runs the tight scalar parser on its own stream.  The words are fetched
// 16 bytes at a time, one vector ahead of use, so the dependent chain never waits on
// memory; decoded values leave as 16-byte stores.  Reads beyond n_words see zero
// bits (bounded, unlike the reference).  `src` must be readable up to the next
// 16-byte boundary past its last word.  Returns false if the stream ran out of
// words before `count` symbols were complete.
__device__ bool lane_rice_decode(const uint32_t *__restrict__ src, uint32_t n_words, uint32_t k,
                                 uint32_t count, int32_t *__restrict__ out)
{
    const uintptr_t addr = reinterpret_cast<uintptr_t>(src);
    const uint32_t skip = (uint32_t)(addr >> 2) & 3u;
    const uint4 *vp = reinterpret_cast<const uint4 *>(addr & ~(uintptr_t)15);
    const uint32_t total = n_words + skip; // words counted from the aligned base
    const uint32_t nvec = (total + 3) >> 2;
    uint32_t vi = 2, wpos = skip;
    uint4 cur = nvec > 0 ? __ldg(vp) : make_uint4(0, 0, 0, 0);
    uint4 nxt = nvec > 1 ? __ldg(vp + 1) : make_uint4(0, 0, 0, 0);
    auto next_word = [&]() -> uint32_t {
        const uint32_t sel = wpos & 3u;
        uint32_t w = sel == 0 ? cur.x : sel == 1 ? cur.y : sel == 2 ? cur.z : cur.w;
        if (wpos >= total)
            w = 0;
        wpos++;
        if ((wpos & 3u) == 0) {
            cur = nxt;
            nxt = vi < nvec ? __ldg(vp + vi) : make_uint4(0, 0, 0, 0);
            vi++;
        }
        return w;
    };

    unsigned long long buf = 0;
    uint32_t avail = 0;
    unsigned long long consumed = 0;
    auto refill = [&]() {
        if (avail <= 32) {
            buf |= (unsigned long long)next_word() << avail;
            avail += 32;
        }
    };
    const bool vec_out = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    int32_t o0 = 0, o1 = 0, o2 = 0;
    refill();
    for (uint32_t i = 0; i < count; i++) {
        uint32_t q = 0;
        while (true) {
            const uint32_t inv = ~(uint32_t)buf;
            if (inv == 0) { // 32 more ones (never past the end: zero bits follow the last word)
                q += 32;
                buf >>= 32;
                avail -= 32;
                consumed += 32;
                refill();
                continue;
            }
            const uint32_t ones = __ffs(inv) - 1;
            q += ones;
            buf >>= ones + 1;
            avail -= ones + 1;
            consumed += ones + 1;
            break;
        }
        refill();
        const uint32_t payload = k ? (__brev((uint32_t)buf) >> (32 - k)) : 0u;
        buf >>= k;
        avail -= k;
        consumed += k;
        refill();
        const uint32_t u = (q << k) | payload; // uint32 shift as in rice_decoder.cpp:37
        const int32_t v = unzigzag(u);
        if (vec_out) {
            const uint32_t sel = i & 3u;
            if (sel == 0) o0 = v;
            else if (sel == 1) o1 = v;
            else if (sel == 2) o2 = v;
            else *reinterpret_cast<int4 *>(out + i - 3) = make_int4(o0, o1, o2, v);
        } else {
            out[i] = v;
        }
    }
    if (vec_out) {
        const uint32_t rem = count & 3u, b = count - rem;
        if (rem > 0) out[b] = o0;
        if (rem > 1) out[b + 1] = o1;
        if (rem > 2) out[b + 2] = o2;
    }
    return consumed <= (unsigned long long)n_words * 32;
}

} // namespace selab200

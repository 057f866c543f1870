// rice_vs.cuh -- Rice decode (K5), second generation: parallelism INSIDE a stream.
//
// rice::RiceDecoder::generateDecodedUnsignedInts (src/rice/rice_decoder.cpp:21-44) is a sequential
// parse: symbol n+1 starts where symbol n ended.  One lane per stream therefore gives a batch of
// 25 838 streams (BASELINE config 3) only 808 warps, each crawling along one dependent chain.
// Here the 2048 symbols of a stream are cut into S parts of 2048/S symbols ("virtual streams") that
// are decoded by S different lanes; what makes that possible is a table of the S-1 bit positions at
// which the parts begin, produced by a first, much cheaper pass:
//
//   k_rice_split_index<S>   S lanes per stream, each parsing ONLY the symbol boundaries of one
//                           contiguous 1/S of the stream's bits, speculatively: a lane that starts
//                           in the middle of the stream does not know where a symbol begins.  Rice
//                           codes resynchronise: the lane before it keeps parsing past its own end
//                           until it lands on a boundary the next lane also found; from there on
//                           the next lane's boundaries are the true ones.  A prefix sum of the
//                           per-lane symbol counts then names the lane (and, through sparse
//                           checkpoints, the bit) where symbol t*2048/S begins.
//   k_rice_decode_vs        one lane per virtual stream: a register bit window over a private
//                           shared-memory ring that the lane tops up with cp.async (no dependent
//                           shared-memory load on the per-symbol chain), values staged in a shared
//                           tile and written as whole 128-byte lines.
//
// Nothing is taken on trust: part l must end exactly where the table says part l+1 begins (part 0
// starts at bit 0, so by induction every part is the sequential parse); a stream that fails that
// check, cannot be split, or meets a symbol the fast paths do not handle is flagged, and the
// general lane-per-stream kernel (k_rice_decode, rice.cuh) decodes it again afterwards.  With S = 1
// the second kernel alone is the large-batch decoder.
#pragma once

#include "rice.cuh"

namespace selab200 {

constexpr uint32_t kNoSplit = 0xffffffffu;

struct RiceVsParams {
    const selab200_subframe_desc *descs;
    uint32_t n_sub, channels;
    const uint32_t *words;
    unsigned long long n_words;
    int32_t *out;      // [n_sub][2048]
    uint32_t *table;   // [n_sub][S-1]: bit position (from the stream's first bit) of symbol t*2048/S
    uint32_t *flags;   // [n_sub]: 1 = the general kernel must decode this stream
    int32_t *status;
    uint32_t cap_words; // k_rice_split_index: staged words per stream
};

__device__ __forceinline__ bool rice_desc_ok(const selab200_subframe_desc &d, uint32_t channels, unsigned long long n_words)
{
    return d.channel < channels && d.parent_channel < channels && d.subframe_type <= 1 && d.lpc_order <= kMaxOrder &&
           d.refl_rice_param < 32 && d.res_rice_param < 32 && d.samples == kFrame &&
           d.refl_offset + d.refl_words <= n_words && d.res_offset + d.res_words <= n_words &&
           !(d.subframe_type == 1 && d.parent_channel == d.channel);
}

// ------------------------------------------------------------------ split index --

constexpr int kCpDense = 16;             // checkpoints at symbols 0, 4, .., 60 of a lane's own parse,
constexpr int kCpMax = kCpDense + 66;    // then at 64, 96, ..: boundary positions relative to the chunk start
__device__ __forceinline__ uint32_t cp_symbol(uint32_t i) { return i < kCpDense ? 4 * i : 64 + 32 * (i - kCpDense); }
__device__ __forceinline__ uint32_t cp_index(uint32_t j) { return j < 64 ? j >> 2 : kCpDense + ((j - 64) >> 5); }

// Position of the symbol boundary after the one at bit q.  sw: the stream's words, bit-reversed
// (stream bit b = bit 31 - b%32 of sw[b/32]), zero beyond total_bits.  Any run length.
__device__ __noinline__ uint32_t rice_next_boundary(const uint32_t *sw, uint32_t total_bits, uint32_t q, uint32_t kp1)
{
    while (q < total_bits) {
        const uint32_t w = q >> 5;
        const uint32_t c = __clz(~__funnelshift_l(sw[w + 1], sw[w], q));
        q += c;
        if (c < 32)
            return q + kp1;
    }
    return total_bits + kp1; // ran off the end: the zero padding terminates the run
}

template <int LOG2S>
__global__ void __launch_bounds__(128) k_rice_split_index(RiceVsParams p)
{
    constexpr int S = 1 << LOG2S, W = 32 >> LOG2S; // lanes per stream, streams per warp
    constexpr uint32_t kPart = kFrame >> LOG2S;
    extern __shared__ __align__(16) unsigned char split_smem[];
    const uint32_t pitch = p.cap_words + 8; // + zero words behind the last one
    const int lane = lane_id(), warp = warp_id();
    uint32_t *stage = reinterpret_cast<uint32_t *>(split_smem) + (size_t)warp * W * pitch;
    uint16_t *cps = reinterpret_cast<uint16_t *>(split_smem + (size_t)(blockDim.x >> 5) * W * pitch * 4) + (size_t)warp * kCpMax * 32;

    const uint32_t g = lane >> LOG2S, l = lane & (S - 1), gb = lane & ~(S - 1);
    const uint32_t st = (blockIdx.x * (blockDim.x >> 5) + warp) * W + g;
    const bool exists = st < p.n_sub;
    selab200_subframe_desc d;
    memset(&d, 0, sizeof d);
    if (exists)
        d = p.descs[st];
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p.words + d.res_offset);
    const uint32_t skip = (uint32_t)(addr >> 2) & 3u;
    const uint4 *gvec = reinterpret_cast<const uint4 *>(addr & ~(uintptr_t)15);
    const uint32_t total = (uint32_t)d.res_words + skip; // words from the aligned base
    const uint32_t kp1 = (uint32_t)d.res_rice_param + 1;
    // splittable: well-formed, fits the staging area, every lane gets at least four words
    bool ok = exists && rice_desc_ok(d, p.channels, p.n_words) && total <= p.cap_words && d.res_words >= 4u * S;

    // ---- stage the W streams of this warp, bit-reversed, zero behind the end ----
    uint32_t *sw = stage + (size_t)g * pitch;
#pragma unroll 1
    for (int t = 0; t < W; t++) {
        const int src_lane = t << LOG2S;
        const bool ok_t = __shfl_sync(kFull, ok, src_lane);
        const uint32_t total_t = __shfl_sync(kFull, total, src_lane);
        const unsigned long long gv = __shfl_sync(kFull, (unsigned long long)reinterpret_cast<uintptr_t>(gvec), src_lane);
        if (!ok_t)
            continue;
        const uint4 *gp = reinterpret_cast<const uint4 *>((uintptr_t)gv);
        uint4 *dst = reinterpret_cast<uint4 *>(stage + (size_t)t * pitch);
        const uint32_t nvec = (total_t + 8 + 3) >> 2; // through the zero words
        for (uint32_t i = lane; i < nvec; i += 32) {
            uint4 v = 4 * i < total_t ? __ldg(gp + i) : make_uint4(0, 0, 0, 0);
            v.x = 4 * i + 0 < total_t ? __brev(v.x) : 0u;
            v.y = 4 * i + 1 < total_t ? __brev(v.y) : 0u;
            v.z = 4 * i + 2 < total_t ? __brev(v.z) : 0u;
            v.w = 4 * i + 3 < total_t ? __brev(v.w) : 0u;
            if (4 * i + 3 < (uint32_t)pitch)
                dst[i] = v;
        }
    }
    __syncwarp();

    // ---- phase 1: every lane parses the boundaries of its own chunk of bits ----
    const uint32_t total_bits = total * 32;
    const uint32_t cw = total >> LOG2S;                                  // words per chunk (>= 4); the last lane takes the rest
    const uint32_t cstart = l == 0 ? skip * 32 : l * cw * 32;           // lane 0 starts at the stream's first bit: a true boundary
    const uint32_t cend = l == S - 1 ? total_bits : (l + 1) * cw * 32;
    const uint32_t maxfast = 32 - kp1; // a symbol with more ones than this does not fit one 32-bit window
    uint32_t n = 0, exitp = cstart, ncp = 0;
    bool fail = false;
    if (ok && cstart < cend) {
        uint32_t pos = cstart, j = 0, next_cp = 0;
        const uint32_t *wa = sw + (pos >> 5) + 1;
        uint32_t r0 = wa[-1], r1 = wa[0];
        while (true) {
            if (j == next_cp) {
                const uint32_t rel = pos - cstart;
                if (ncp < (uint32_t)kCpMax && rel < 0xffffu)
                    cps[ncp * 32 + lane] = (uint16_t)rel;
                else
                    fail = true;
                ncp++;
                next_cp += j < 64 ? 4 : 32;
            }
            uint32_t pp[5];
            pp[0] = pos;
            uint32_t mx = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const uint32_t ones = __clz(~__funnelshift_l(r1, r0, pos));
                mx = max(mx, ones);
                const uint32_t pn = pos + ones + kp1;
                if ((pos ^ pn) >= 32u) {
                    r0 = r1;
                    wa++;
                    r1 = *wa;
                }
                pos = pn;
                pp[e + 1] = pos;
            }
            if (mx > maxfast) { // a long symbol in this batch: redo it with the general step
                uint32_t q = pp[0];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    q = rice_next_boundary(sw, total_bits, q, kp1);
                    pp[e + 1] = q;
                }
                pos = q;
                const uint32_t w = min(pos >> 5, total + 6);
                wa = sw + w + 1;
                r0 = wa[-1];
                r1 = wa[0];
            }
            if (pp[4] < cend && j + 4 < (uint32_t)kFrame + 64) {
                j += 4;
                continue;
            }
            const uint32_t e = pp[1] >= cend ? 1 : pp[2] >= cend ? 2 : pp[3] >= cend ? 3 : 4;
            n = j + e;
            exitp = pp[e];
            break;
        }
    }
    if (ncp > (uint32_t)kCpMax)
        fail = true;
    __syncwarp();

    // ---- phase 2: run on into the next lane's chunk until landing on one of its boundaries ----
    const uint32_t ncp_s = __shfl_down_sync(kFull, ncp, 1), n_s = __shfl_down_sync(kFull, n, 1);
    const uint32_t exit_s = __shfl_down_sync(kFull, exitp, 1), cstart_s = __shfl_down_sync(kFull, cstart, 1);
    uint32_t x = 0, m_next = 0; // overflow symbols of this lane; index (in the next lane's parse) of the merge boundary
    if (ok && !fail && l < S - 1) {
        uint32_t q = exitp, i = 0;
        const uint16_t *cn = cps + lane + 1;
        for (uint32_t guard = 0;; guard++) {
            const uint32_t rel = q - cstart_s;
            if (rel >= 0xffffu || guard > 4096u) {
                fail = true;
                break;
            }
            while (i < ncp_s && cn[i * 32] < rel)
                i++;
            if (i < ncp_s) {
                if (cn[i * 32] == rel) {
                    m_next = cp_symbol(i);
                    break;
                }
            } else if (q >= exit_s) { // behind the last checkpoint: only the next lane's exit is left to meet
                if (q == exit_s)
                    m_next = n_s;
                else
                    fail = true;
                break;
            }
            q = rice_next_boundary(sw, total_bits, q, kp1);
            x++;
        }
    }
    // ---- phase 3: symbol index of every lane's first true boundary ----
    const uint32_t m_up = __shfl_up_sync(kFull, m_next, 1);
    const uint32_t m = l == 0 ? 0u : m_up;         // this lane's parse is the true one from its symbol m on
    const uint32_t valid = n >= m ? n - m : 0u;
    if (n < m)
        fail = true;
    uint32_t incl = valid + x;
#pragma unroll
    for (int o = 1; o < S; o <<= 1) {
        const uint32_t t = __shfl_up_sync(kFull, incl, o, S);
        if ((int)l >= o)
            incl += t;
    }
    const uint32_t base = incl - (valid + x);
    const uint32_t fail_mask = __ballot_sync(kFull, fail || !ok);
    const bool group_bad = ((fail_mask >> gb) & ((S == 32) ? 0xffffffffu : ((1u << S) - 1u))) != 0;

    // ---- phase 4: lane l looks up where symbol l*2048/S begins ----
    const uint32_t target = l * kPart;
    uint32_t ls = 0;
#pragma unroll
    for (int ll = 1; ll < S; ll++) {
        const uint32_t b = __shfl_sync(kFull, base, gb + ll);
        if (b <= target)
            ls = ll;
    }
    const uint32_t b_s = __shfl_sync(kFull, base, gb + ls), valid_s = __shfl_sync(kFull, valid, gb + ls);
    const uint32_t m_s = __shfl_sync(kFull, m, gb + ls), exitp_s = __shfl_sync(kFull, exitp, gb + ls);
    const uint32_t cst_s = __shfl_sync(kFull, cstart, gb + ls), x_s = __shfl_sync(kFull, x, gb + ls);
    if (exists && l == 0)
        p.flags[st] = group_bad ? 1u : 0u;
    if (exists && l > 0) {
        uint32_t result = kNoSplit;
        if (!group_bad) {
            const uint32_t rel = target - b_s;
            uint32_t q, walk;
            bool found = true;
            if (rel < valid_s) {
                const uint32_t jj = m_s + rel, i = cp_index(jj);
                q = cst_s + cps[i * 32 + gb + ls];
                walk = jj - cp_symbol(i);
            } else {
                q = exitp_s;
                walk = rel - valid_s;
                found = walk < x_s;
            }
            for (uint32_t t = 0; t < walk && t < 512u; t++)
                q = rice_next_boundary(sw, total_bits, q, kp1);
            result = found ? q - skip * 32 : 0u; // a stream with too few symbols: the decoder's end check catches it
        }
        p.table[(size_t)st * (S - 1) + (l - 1)] = result;
    }
}

inline size_t rice_split_smem_bytes(int log2s, uint32_t cap_words, int warps)
{
    const int W = 32 >> log2s;
    return (size_t)warps * W * (cap_words + 8) * 4 + (size_t)warps * kCpMax * 32 * 2;
}

// --------------------------------------------------------------- virtual streams --

constexpr int kVsRing = 64;        // ring words per lane: one 256-byte row
constexpr int kVsRound = 16;       // symbols per lane between two ring top-ups
constexpr int kVsTilePitch = 36;   // words; 32 symbols per lane per flush, rows 16-byte aligned and bank-rotated
constexpr int kVsWarps = 4;

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// A lane's ring: 64 words in one 256-byte row of shared memory, word w at byte ((4w + rot) & 255)
// of the row (rot = 16 * lane spreads the lanes over the banks).
struct VsRing {
    uint32_t row, rot;  // shared-space byte address of the row (256-aligned); rotation
    const uint4 *gvec;  // the stream from its 16-byte aligned base
    int total_bytes;    // bytes from gvec to the end of the stream; everything behind reads as zero
    __device__ __forceinline__ uint32_t word_addr(uint32_t w) const { return row | ((4 * w + rot) & (kVsRing * 4 - 1)); }
    __device__ __forceinline__ uint32_t word(uint32_t w) const { return lds_u32(word_addr(w)); } // reversed words only (below `ce`)
    __device__ __forceinline__ void issue(uint32_t fv) const // vector fv: words [4fv, 4fv+4)
    {
        const int rem = total_bytes - (int)(16 * fv);
        const uint32_t sz = rem <= 0 ? 0u : rem < 16 ? (uint32_t)rem : 16u;
        const uint4 *src = gvec + (sz ? fv : 0u);
        const uint32_t dst = row | ((16 * fv + rot) & (kVsRing * 4 - 1));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    }
    // The parser wants the stream MSB first (leading-zero count finds the terminator, the payload reads as a
    // number); BREV costs three issue slots on sm_100a, so the words are reversed once, in place, when their
    // vector has landed -- not once per symbol.
    __device__ __forceinline__ void reverse(uint32_t v) const
    {
        const uint32_t a = row | ((16 * v + rot) & (kVsRing * 4 - 1));
        uint32_t x, y, z, w;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a));
        x = __brev(x), y = __brev(y), z = __brev(z), w = __brev(w);
        asm volatile("st.shared.v4.u32 [%4], {%0, %1, %2, %3};" ::"r"(x), "r"(y), "r"(z), "r"(w), "r"(a) : "memory");
    }
};

// General decode of `count` symbols from bit `pos`, ring words only up to `ce` (exclusive, complete).
// Out of line: runs for the rare round that holds a symbol longer than one 32-bit window.
// Returns false if it would need words the ring does not hold (the stream is then flagged).
__device__ __noinline__ bool vs_slow_round(const VsRing rg, uint32_t ce, uint32_t &pos, uint32_t k, int32_t *dst, int count)
{
    uint32_t q = pos;
    for (int e = 0; e < count; e++) {
        uint32_t ones = 0;
        while (true) {
            const uint32_t w = q >> 5;
            if (w + 2 > ce)
                return false;
            const uint32_t c = __clz(~__funnelshift_l(rg.word(w + 1), rg.word(w), q));
            ones += c;
            q += c;
            if (c < 32)
                break;
        }
        q += 1;
        const uint32_t w = q >> 5;
        if (w + 2 > ce)
            return false;
        const uint32_t win = __funnelshift_l(rg.word(w + 1), rg.word(w), q);
        const uint32_t pay = __funnelshift_rc(win, 0u, 32 - k);
        q += k;
        dst[e] = unzigzag((ones << k) | pay); // uint32 shift as in rice_decoder.cpp:37
    }
    if ((q >> 5) + 2 > ce)
        return false;
    pos = q;
    return true;
}

__global__ void __launch_bounds__(32 * kVsWarps) k_rice_decode_vs(RiceVsParams p, int log2s)
{
    extern __shared__ __align__(16) unsigned char vs_smem[];
    const int lane = lane_id(), warp = warp_id();
    // ring rows must be 256-byte aligned in the shared window (the word address is formed with an OR)
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(vs_smem);
    const uint32_t pad = (256u - (smem0 & 255u)) & 255u;
    const uint32_t ring_base = smem0 + pad + (uint32_t)(warp * 32 + lane) * (kVsRing * 4);
    int32_t *tile = reinterpret_cast<int32_t *>(vs_smem + pad + kVsWarps * 32 * kVsRing * 4) + warp * 32 * kVsTilePitch;

    const uint32_t S = 1u << log2s, part = (uint32_t)kFrame >> log2s;
    const uint32_t v0 = (blockIdx.x * kVsWarps + warp) * 32;
    const uint32_t v = v0 + lane, st = v >> log2s, l = v & (S - 1);
    const bool exists = st < p.n_sub;
    selab200_subframe_desc d;
    memset(&d, 0, sizeof d);
    if (exists)
        d = p.descs[st];
    bool ok = exists && rice_desc_ok(d, p.channels, p.n_words);
    if (exists && !ok && l == 0)
        raise_status(p.status, SELAB200_ERR_BITSTREAM);
    const bool store_row = ok; // rows of flagged streams may hold garbage: the general kernel rewrites them
    uint32_t sb = 0, expect_end = kNoSplit;
    if (ok && S > 1) {
        const uint32_t *tb = p.table + (size_t)st * (S - 1);
        if (tb[0] == kNoSplit) {
            ok = false; // not split: the general kernel decodes it
        } else {
            if (l > 0)
                sb = tb[l - 1];
            if (l < S - 1)
                expect_end = tb[l];
        }
    }
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p.words + (ok ? d.res_offset : 0));
    const uint32_t skip = (uint32_t)(addr >> 2) & 3u;
    const uint32_t total = ok && d.res_words ? (uint32_t)d.res_words + skip : 0u;
    VsRing rg;
    rg.row = ring_base;
    rg.rot = (16u * lane) & (kVsRing * 4 - 1);
    rg.gvec = reinterpret_cast<const uint4 *>(addr & ~(uintptr_t)15);
    rg.total_bytes = (int)(total * 4);
    const uint32_t k = ok ? d.res_rice_param : 0u, kp1 = k + 1, kk = 32 - k, kpow = 1u << k;
    const uint32_t maxfast = 32 - kp1;
    const uint32_t row_mask = __ballot_sync(kFull, store_row);

    // ---- prime the ring ----
    uint32_t pos = sb + 32 * skip;
    if (pos > total * 32 + 64)
        pos = total * 32 + 64; // a nonsense table entry: parse zeros, fail the end check
    uint32_t wb = (pos >> 5) + 1; // r1 holds word wb, r0 word wb - 1
    uint32_t fv = (wb - 1) >> 2;
    // vector fv overwrites words [4fv - 64, 4fv - 61]; everything below wb - 1 is dead
    {
        const uint32_t lim = (wb + kVsRing - 5) >> 2;
        while (fv <= lim) {
            rg.issue(fv);
            fv++;
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    uint32_t rv = (wb - 1) >> 2; // vectors below rv are reversed
    while (rv < fv) {
        rg.reverse(rv);
        rv++;
    }
    uint32_t ce = 4 * fv, cv_next = fv; // words below ce are in the ring, reversed; vectors below cv_next have been requested
    uint32_t r0 = rg.word(wb - 1), r1 = rg.word(wb);
    uint32_t wa = 4 * wb + rg.rot;      // running byte offset of word wb
    bool dead = false;

    int32_t *out_warp = p.out + (size_t)v0 * part;
    const uint32_t n_rounds = part / kVsRound;
#pragma unroll 1
    for (uint32_t r = 0; r < n_rounds; r++) {
        // ---- boundary: top the ring up, retire the previous top-up ----
        if (r) {
            const uint32_t lim = (wb + kVsRing - 5) >> 2;
            while (fv <= lim) {
                rg.issue(fv);
                fv++;
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory"); // everything but the group just committed has landed
            while (rv < cv_next) {
                rg.reverse(rv);
                rv++;
            }
            ce = 4 * cv_next;
            cv_next = fv;
        }
        const uint32_t pos_s = pos;
        uint32_t mx = 0;
        int32_t *trow = tile + lane * kVsTilePitch + (r & 1) * kVsRound;
#pragma unroll
        for (int e4 = 0; e4 < kVsRound; e4 += 4) {
            int32_t val[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const uint32_t win = __funnelshift_l(r1, r0, pos);
                const uint32_t ones = __clz(~win);
                mx = max(mx, ones);
                const uint32_t t = __funnelshift_lc(0u, win, ones + 1);
                const uint32_t pay = __funnelshift_rc(t, 0u, kk);
                const uint32_t u = ones * kpow + pay;
                val[e] = unzigzag(u);
                const uint32_t pn = pos + ones + kp1;
                if ((pos ^ pn) >= 32u) {
                    r0 = r1;
                    wa += 4;
                    r1 = lds_u32(rg.row | (wa & (kVsRing * 4 - 1)));
                }
                pos = pn;
            }
            *reinterpret_cast<int4 *>(trow + e4) = make_int4(val[0], val[1], val[2], val[3]);
        }
        if (mx > maxfast && !dead) { // a symbol longer than the window: redo the round with the general parser
            pos = pos_s;
            if (vs_slow_round(rg, ce, pos, k, trow, kVsRound)) {
                wb = (pos >> 5) + 1;
                r0 = rg.word(wb - 1);
                r1 = rg.word(wb);
                wa = 4 * wb + rg.rot;
            } else {
                dead = true;
            }
        }
        wb = dead ? wb : (pos >> 5) + 1;
        // ---- every second round: 32 symbols per lane leave as whole 128-byte lines ----
        if (r & 1) {
            __syncwarp();
            int32_t *dst = out_warp + (size_t)(r >> 1) * 32 + (lane & 7) * 4;
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int row = 4 * it + (lane >> 3);
                if ((row_mask >> row) & 1u) {
                    const int4 q = *reinterpret_cast<const int4 *>(tile + row * kVsTilePitch + (lane & 7) * 4);
                    *reinterpret_cast<int4 *>(dst + (size_t)row * part) = q;
                }
            }
            __syncwarp();
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (exists && store_row) {
        bool bad = dead;
        if (ok) {
            if (expect_end != kNoSplit)
                bad |= pos - 32 * skip != expect_end;
            else
                bad |= pos > total * 32;
        }
        if (bad)
            p.flags[st] = 1u; // (several parts may say so: idempotent)
    }
}

constexpr size_t kVsSmemBytes = 256 + (size_t)kVsWarps * 32 * kVsRing * 4 + (size_t)kVsWarps * 32 * kVsTilePitch * 4;

} // namespace selab200

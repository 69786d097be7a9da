// K5 core: warp-level K=7 r=1/2 soft Viterbi over a staged code-order buffer (shared by viterbi.cu and rtchannel.cu).
// See viterbi.cu for the decoding schedule (the written spec shared with oracle/correct_restated.c).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace jb {

static const int VK = 7;
static const int V_MIN_TB = 5 * VK;              // 35
static const int V_GROUP = 15 * VK;              // 105
static const int V_CAP = V_MIN_TB + V_GROUP;     // 140
static const int V_RENORM = 65535 / (2 * 255);   // 128
static const unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned parity7(unsigned x) { return __popc(x) & 1u; }
// table[sr] bit p = parity(sr & poly[p]), polys 109, 79 (jconvolutionalcodec.cpp:13-14)
__device__ __forceinline__ unsigned conv_out(unsigned sr) { return parity7(sr & 109u) | (parity7(sr & 79u) << 1); }

struct WarpHist {
    uint2 *h;                // decision words of a step: x = even successors 2l, y = odd successors 2l+1; V_CAP entries
    uint8_t *obits;          // decoded bits, oldest first
    int index, len, nout;
};

// history_buffer_traceback: every lane walks the same survivor (no divergence); lane 0 stores.
__device__ __forceinline__ void traceback(WarpHist &H, unsigned state, int min_tb, int lane)
{
    int idx = H.index;
    const int len = H.len;
    const int g = len - min_tb;
    uint8_t *ob = H.obits + H.nout + (g - 1) + min_tb;          // step j of the walk decides position ob[-j]
    for (int j = 0; j < len; j++) {
        idx = (idx == 0) ? V_CAP - 1 : idx - 1;
        const uint2 d = H.h[idx];
        const unsigned w = (state & 1u) ? d.y : d.x;
        const unsigned h = (w >> (state >> 1)) & 1u;
        state = (state >> 1) | (h << 5);
        if (j >= min_tb && lane == 0) ob[-j] = (uint8_t)h;
    }
    if (g > 0) { H.nout += g; H.len -= g; }
}

// Lane constants of the add-compare-select step. Both generator polynomials (109, 79) have their first and last taps set,
// so the four branch labels a lane needs are one label t = table[2l] and its complement: table[2l | 64] = table[2l + 1] =
// t ^ 3, table[(2l + 1) | 64] = t. With soft values a, b in 0..255 the distance to label t is (a ^ ma) + (b ^ mb) where
// ma / mb = 0xff when the label's bit is set (|soft - 255| = soft ^ 0xff), and the distance to t ^ 3 is 510 minus it.
struct VitLane {
    unsigned s0, s1, m16, psel;
    int src_lo, src_hi, lane;
};

struct VitState { unsigned m; int rc; WarpHist H; };

// One trellis step with history. TAIL: the zero-tail steps (ties go to the high predecessor, only states whose low bits
// are zero compete for the best state).
template <bool TAIL>
__device__ __forceinline__ void acs_step(VitState &S, const VitLane &L, unsigned d, int i, int sets)
{
    const unsigned dn = 510u - d;
    const unsigned vlo = __shfl_sync(FULL, S.m, L.src_lo), vhi = __shfl_sync(FULL, S.m, L.src_hi);
    const unsigned mlo = __byte_perm(vlo, 0u, L.psel), mhi = __byte_perm(vhi, 0u, L.psel);
    const unsigned lo0 = (d + mlo) & 0xffffu, hi0 = (dn + mhi) & 0xffffu;
    const unsigned lo1 = (dn + mlo) & 0xffffu, hi1 = (d + mhi) & 0xffffu;
    // inner: ties -> low predecessor (<=); tail: ties -> high predecessor (<). The survivor's metric is the minimum either way.
    const bool take_hi0 = TAIL ? (lo0 >= hi0) : (lo0 > hi0);
    const bool take_hi1 = TAIL ? (lo1 >= hi1) : (lo1 > hi1);
    unsigned e0 = min(lo0, hi0), e1 = min(lo1, hi1);
    const unsigned w0 = __ballot_sync(FULL, take_hi0), w1 = __ballot_sync(FULL, take_hi1);
    if (L.lane == 0) S.H.h[S.H.index] = make_uint2(w0, w1);
    // history_buffer_process_skip
    S.H.index++; if (S.H.index == V_CAP) S.H.index = 0;
    S.rc++; S.H.len++;
    const bool renorm = (S.rc == V_RENORM);
    const bool tb = (S.H.len == V_CAP);
    if (renorm || tb) {
        const unsigned skip = TAIL ? (1u << (VK - (sets - i))) : 1u;
        unsigned k0 = ((L.s0 & (skip - 1u)) == 0u) ? ((e0 << 6) | L.s0) : 0xffffffffu;
        unsigned k1 = ((L.s1 & (skip - 1u)) == 0u) ? ((e1 << 6) | L.s1) : 0xffffffffu;
        unsigned best = __reduce_min_sync(FULL, min(k0, k1));
        if (renorm) {
            S.rc = 0;
            unsigned mn = best >> 6;
            if ((L.s0 & (skip - 1u)) == 0u) e0 = (e0 - mn) & 0xffffu;
            if ((L.s1 & (skip - 1u)) == 0u) e1 = (e1 - mn) & 0xffffu;
        }
        if (tb) { __syncwarp(); traceback(S.H, best & 63u, V_MIN_TB, L.lane); }
    }
    S.m = e0 | (e1 << 16);
}

// Decode `sets` trellis steps from sbuf[0 .. 2*sets) (soft values 0..255, code order; 2-byte aligned). hist is a V_CAP-entry
// decision ring. rc = libcorrect's renormalisation counter (persists across calls). Returns the number of decoded bits
// written to obits[0 ..) (= sets - (K-1) once sets >= K-1); later positions are not touched. obits may be sbuf itself: the
// bit of step p is written after step p + V_MIN_TB has been read, at byte p < 2p.
__device__ __forceinline__ int viterbi_decode_warp(const uint8_t *sbuf, int sets, uint2 *hist, uint8_t *obits, int &rc, int lane)
{
    VitLane L;
    L.lane = lane; L.s0 = 2u * lane; L.s1 = L.s0 + 1u;
    const unsigned t = conv_out(L.s0);
    L.m16 = ((t & 1u) ? 0x00ffu : 0u) | ((t & 2u) ? 0xff00u : 0u);
    L.psel = (lane & 1) ? 0x4432u : 0x4410u;         // byte_perm selector: the odd / even half of a packed metric pair
    L.src_lo = lane >> 1; L.src_hi = 16 + (lane >> 1);
    const unsigned short *sb16 = reinterpret_cast<const unsigned short *>(sbuf);
#define VDIST(x16) ((((unsigned)(x16) ^ L.m16) & 0xffu) + (((unsigned)(x16) ^ L.m16) >> 8))

    VitState S;
    S.m = 0;                              // packed metrics: e0 | e1<<16 (error_buffer_reset -> 0)
    S.rc = rc;
    S.H.h = hist; S.H.obits = obits; S.H.index = 0; S.H.len = 0; S.H.nout = 0;

    // warm-up: errors[j] = dist(table[j]) + errors[j>>1] for the states reachable so far; no history
    const int nwarm = sets < VK - 1 ? sets : VK - 1;
    for (int i = 0; i < nwarm; i++) {
        const unsigned d = VDIST(sb16[i]), dn = 510u - d;
        const unsigned vlo = __shfl_sync(FULL, S.m, L.src_lo);
        const unsigned mlo = __byte_perm(vlo, 0u, L.psel);
        const unsigned e0 = (d + mlo) & 0xffffu, e1 = (dn + mlo) & 0xffffu;
        const unsigned lim = 1u << (i + 1);
        unsigned o0 = S.m & 0xffffu, o1 = S.m >> 16;
        if (L.s0 < lim) o0 = e0;
        if (L.s1 < lim) o1 = e1;
        S.m = o0 | (o1 << 16);
    }
    const int tail_start = (sets - (VK - 1) > VK - 1) ? sets - (VK - 1) : VK - 1;
    int i = VK - 1;
    unsigned x = (i < sets) ? sb16[i] : 0u;
    for (; i < tail_start; i++) {
        const unsigned d = VDIST(x);
        x = sb16[i + 1];                  // i + 1 <= sets - (K-1): inside the staged buffer
        acs_step<false>(S, L, d, i, sets);
    }
    for (; i < sets; i++) {
        const unsigned d = VDIST(sb16[i]);
        acs_step<true>(S, L, d, i, sets);
    }
#undef VDIST
    __syncwarp();
    traceback(S.H, 0u, 0, lane);          // history_buffer_flush
    rc = S.rc;
    return S.H.nout;
}

} // namespace jb

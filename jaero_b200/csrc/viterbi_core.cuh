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
    unsigned *h0, *h1;       // decision words of even / odd successors, V_CAP entries each
    uint8_t *obits;          // decoded bits, oldest first
    int index, len, nout;
};

// history_buffer_traceback: every lane walks the same survivor (no divergence); lane 0 stores.
__device__ __forceinline__ void traceback(WarpHist &H, unsigned state, int min_tb, int lane)
{
    int idx = H.index;
    const int len = H.len;
    const int g = len - min_tb;
    for (int j = 0; j < len; j++) {
        idx = (idx == 0) ? V_CAP - 1 : idx - 1;
        unsigned w = (state & 1u) ? H.h1[idx] : H.h0[idx];
        unsigned h = (w >> (state >> 1)) & 1u;
        state = (state >> 1) | (h << 5);
        if (j >= min_tb && lane == 0) H.obits[H.nout + (g - 1 - (j - min_tb))] = (uint8_t)h;
    }
    if (g > 0) { H.nout += g; H.len -= g; }
}

// Decode `sets` trellis steps from sbuf[0 .. 2*sets) (soft values 0..255, code order). obits must be zero-filled for
// [0, sets); h0/h1 are V_CAP-entry decision rings. rc = libcorrect's renormalisation counter (persists across calls).
__device__ __forceinline__ void viterbi_decode_warp(const uint8_t *__restrict__ sbuf, int sets, unsigned *h0, unsigned *h1, uint8_t *obits, int &rc, int lane)
{
    // per-lane branch outputs for successors 2l (e=0) and 2l+1 (e=1)
    const unsigned s0 = 2u * lane, s1 = 2u * lane + 1u;
    const unsigned tl0 = conv_out(s0), th0 = conv_out(s0 | 64u), tl1 = conv_out(s1), th1 = conv_out(s1 | 64u);
    const int src_lo = lane >> 1, src_hi = 16 + (lane >> 1);
    const bool odd = lane & 1;

    unsigned m = 0;                       // packed metrics: e0 | e1<<16 (error_buffer_reset -> 0)
    WarpHist H; H.h0 = h0; H.h1 = h1; H.obits = obits; H.index = 0; H.len = 0; H.nout = 0;

    for (int i = 0; i < sets; i++) {
        const unsigned a0 = sbuf[2 * i], b0 = sbuf[2 * i + 1];
        const unsigned a1 = 255u - a0, b1 = 255u - b0;    // |soft-255|
        const unsigned vlo = __shfl_sync(FULL, m, src_lo), vhi = __shfl_sync(FULL, m, src_hi);
        const unsigned mlo = odd ? (vlo >> 16) : (vlo & 0xffffu);
        const unsigned mhi = odd ? (vhi >> 16) : (vhi & 0xffffu);
#define DSEL(t) (((t) & 1u ? a1 : a0) + ((t) & 2u ? b1 : b0))
        if (i < VK - 1) {
            // warm-up: errors[j] = dist(table[j]) + errors[j>>1] for the states reachable so far; no history
            unsigned e0 = (DSEL(tl0) + mlo) & 0xffffu, e1 = (DSEL(tl1) + mlo) & 0xffffu;
            unsigned lim = 1u << (i + 1);
            unsigned o0 = m & 0xffffu, o1 = m >> 16;
            if (s0 < lim) o0 = e0;
            if (s1 < lim) o1 = e1;
            m = o0 | (o1 << 16);
            continue;
        }
        const bool tail = (i + (VK - 1) >= sets);
        const unsigned lo0 = (DSEL(tl0) + mlo) & 0xffffu, hi0 = (DSEL(th0) + mhi) & 0xffffu;
        const unsigned lo1 = (DSEL(tl1) + mlo) & 0xffffu, hi1 = (DSEL(th1) + mhi) & 0xffffu;
#undef DSEL
        // inner: ties -> low predecessor (<=); tail: ties -> high predecessor (<)
        const bool pick_lo0 = tail ? (lo0 < hi0) : (lo0 <= hi0);
        const bool pick_lo1 = tail ? (lo1 < hi1) : (lo1 <= hi1);
        unsigned e0 = pick_lo0 ? lo0 : hi0, e1 = pick_lo1 ? lo1 : hi1;
        const unsigned w0 = __ballot_sync(FULL, !pick_lo0), w1 = __ballot_sync(FULL, !pick_lo1);
        if (lane == 0) { h0[H.index] = w0; h1[H.index] = w1; }
        __syncwarp();
        const unsigned skip = tail ? (1u << (VK - (sets - i))) : 1u;
        // history_buffer_process_skip
        H.index++; if (H.index == V_CAP) H.index = 0;
        rc++; H.len++;
        const bool renorm = (rc == V_RENORM);
        const bool tb = (H.len == V_CAP);
        if (renorm || tb) {
            unsigned k0 = ((s0 & (skip - 1u)) == 0u) ? ((e0 << 6) | s0) : 0xffffffffu;
            unsigned k1 = ((s1 & (skip - 1u)) == 0u) ? ((e1 << 6) | s1) : 0xffffffffu;
            unsigned best = __reduce_min_sync(FULL, min(k0, k1));
            if (renorm) {
                rc = 0;
                unsigned mn = best >> 6;
                if ((s0 & (skip - 1u)) == 0u) e0 = (e0 - mn) & 0xffffu;
                if ((s1 & (skip - 1u)) == 0u) e1 = (e1 - mn) & 0xffffu;
            }
            if (tb) traceback(H, best & 63u, V_MIN_TB, lane);
        }
        m = e0 | (e1 << 16);
    }
    traceback(H, 0u, 0, lane);            // history_buffer_flush
}

} // namespace jb

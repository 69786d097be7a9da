// K5 — batched K=7, r=1/2 soft-decision Viterbi decoder, one warp per channel.
//
// Replaces (reference): AeroLInterleaver::deinterleave_ba (JAERO/aerol.cpp:603-625, row permutation
// (i*27)%64 at :531-535) + JConvolutionalCodec::Decode_Continuous / Decode_soft
// (JAERO/jconvolutionalcodec.cpp:151-201, 98-125), whose arithmetic lives in the un-vendored
// quiet/libcorrect. The decoding schedule below is the written spec shared with the CPU oracle
// (oracle/correct_restated.c): warm-up (K-1 steps), add-compare-select with ties to the
// "low" predecessor, a zero-tail of K-1 steps with ties to the "high" predecessor, a 140-slice
// history ring that emits 105 decisions after walking back 35, min-subtraction every 128 steps,
// final flush from state 0. Integer arithmetic only -> bit-exact against the oracle.
//
// Mapping: lane l owns successor states 2l and 2l+1 (both have predecessors l and l+32), path
// metrics are two uint16 packed in one register so a step needs two warp shuffles; the 64
// decision bits of a step are two __ballot_sync words kept in a shared-memory ring; the
// best-state search is one __reduce_min_sync on (metric<<6 | state).
#include "common.cuh"
#include "viterbi.cuh"
#include "viterbi_core.cuh"

namespace jb {

__global__ void __launch_bounds__(128)
viterbi_k7_kernel(const uint8_t *__restrict__ soft_in, size_t in_stride, size_t out_stride,
                  const int *__restrict__ active_count, int active_q, int n_soft, int cols, int mode, int pad,
                  uint8_t *__restrict__ overlap, int *__restrict__ overlap_len,
                  int *__restrict__ renorm_counter, uint8_t *__restrict__ bits_out, int *__restrict__ n_valid, int n_channels,
                  int smem_per_warp, int sbuf_bytes)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ch = blockIdx.x * (blockDim.x >> 5) + warp;
    if (ch >= n_channels) return;
    if (active_count && active_count[ch] <= active_q) return;      // nothing queued for this channel
    unsigned char *base = smem + (size_t)warp * smem_per_warp;
    uint8_t *sbuf = base;
    uint2 *hist = reinterpret_cast<uint2 *>(base + sbuf_bytes);
    uint8_t *obits = sbuf;                 // decoded bits overwrite the consumed head of the soft buffer (viterbi_core.cuh)

    // ---- stage the code-order soft buffer: overlap ‖ (de-interleaved) block ‖ erasure padding
    const int ov = (mode == 0) ? overlap_len[ch] : 0;
    const int padn = (mode == 0) ? pad : 0;
    const int total = ov + n_soft + padn;
    const uint8_t *src = soft_in + (size_t)ch * in_stride;
    for (int k = lane; k < total; k += 32) {
        uint8_t v;
        if (k < ov) v = overlap[ch * 64 + k];
        else if (k < ov + n_soft) {
            int q = k - ov;
            if (cols > 0) { int i = q & 63, j = q >> 6; v = src[((i * 27) & 63) * cols + j]; }   // deinterleave_ba
            else v = src[q];
        } else v = 128;
        sbuf[k] = v;
    }
    __syncwarp();
    const int sets = total >> 1;
    // right(62) of the new code-order block, zero-extended to 62: taken before the decoded bits overwrite the buffer
    const int kk = n_soft < 62 ? n_soft : 62;
    uint8_t keep0 = 0, keep1 = 0;
    if (mode == 0) {
        if (lane < kk) keep0 = sbuf[ov + n_soft - kk + lane];
        if (lane + 32 < kk) keep1 = sbuf[ov + n_soft - kk + lane + 32];
    }
    __syncwarp();

    int rc = renorm_counter[ch];
    const int nout = viterbi_decode_warp(sbuf, sets, hist, obits, rc, lane);
    renorm_counter[ch] = rc;
    __syncwarp();

    // ---- outputs
    const int nbits = n_soft >> 1;
    uint8_t *out = bits_out + (size_t)ch * out_stride;
    if (mode == 0) {
        // Decode_Continuous: mid(paddinglength+1, n/2); positions never written by the decoder read as 0
        const int pos = pad + 1;
        for (int k = lane; k < nbits; k += 32) out[k] = (pos + k < nout) ? obits[pos + k] : (uint8_t)0;
        if (lane == 0 && n_valid) { int nv = (sets - pos < nbits) ? (sets - pos) : nbits; n_valid[ch] = nv < 0 ? 0 : nv; }   // QVector::mid truncation (empty, never negative)
        overlap[ch * 64 + lane] = keep0;
        if (lane + 32 < 62) overlap[ch * 64 + lane + 32] = keep1;
        if (lane == 0) overlap_len[ch] = 62;
    } else {
        for (int k = lane; k < nbits; k += 32) out[k] = (k < nout) ? obits[k] : (uint8_t)0;   // last K-1 positions stay 0
        if (lane == 0 && n_valid) n_valid[ch] = nbits;
    }
}

size_t viterbi_smem_per_warp(int n_soft, int pad, int *sbuf_bytes)
{
    int total = 62 + n_soft + pad;
    int sb = (total + 2 + 15) & ~15;              // + one look-ahead pair
    *sbuf_bytes = sb;
    return (size_t)sb + V_CAP * sizeof(uint2);
}

int viterbi_launch(const uint8_t *d_soft, int n_soft, int cols, int mode, int pad, uint8_t *d_overlap,
                   int *d_overlap_len, int *d_renorm, uint8_t *d_bits, int *d_valid, int n_channels, cudaStream_t stream,
                   size_t in_stride, size_t out_stride, const int *d_active_count, int active_q)
{
    if (in_stride == 0) in_stride = (size_t)n_soft;
    if (out_stride == 0) out_stride = (size_t)(n_soft / 2);
    int sbuf_bytes = 0;
    size_t per_warp = viterbi_smem_per_warp(n_soft, pad, &sbuf_bytes);
    const int warps = 4;
    size_t smem = per_warp * warps;
    if (smem > 200 * 1024) { set_error("viterbi: block too long for shared memory"); return -1; }
    JB_CUDA(cudaFuncSetAttribute(viterbi_k7_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int grid = (n_channels + warps - 1) / warps;
    viterbi_k7_kernel<<<grid, warps * 32, smem, stream>>>(d_soft, in_stride, out_stride, d_active_count, active_q, n_soft, cols, mode, pad, d_overlap, d_overlap_len,
                                                          d_renorm, d_bits, d_valid, n_channels, (int)per_warp, sbuf_bytes);
    JB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace jb

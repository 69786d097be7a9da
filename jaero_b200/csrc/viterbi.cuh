#pragma once
#include <cuda_runtime.h>
#include <cstdint>
namespace jb {
// mode 0: Decode_Continuous (overlap + erasure padding carried per channel); mode 1: one-shot Decode_soft
int viterbi_launch(const uint8_t *d_soft, int n_soft, int cols, int mode, int pad, uint8_t *d_overlap,
                   int *d_overlap_len, int *d_renorm, uint8_t *d_bits, int *d_valid, int n_channels, cudaStream_t stream,
                   size_t in_stride = 0, size_t out_stride = 0, const int *d_active_count = nullptr, int active_q = 0);
}

// R/T burst channel layer — device data layout and launch prototypes (see rtchannel.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace jb {

static const int RT_BLOCK = 64 * 95;       // RTChannelDeleaveFECScram::block (aerol.h:575)
static const int RT_SLOTS = 8;             // packets that may be open per channel between two trial passes
static const int RT_OUT = 8;               // decoded packets held per channel until read
static const int RT_OUT_BYTES = 400;       // 16-byte header + up to 379 payload bytes

struct RtState {                           // AeroL members used by the burst branch of Decode (aerol.h:937-1016)
    unsigned sr_imag, sr_real, sr_msk;
    int inv_imag, inv_real, inv_msk, realimag, gotsync_last;
    int cntr, muw, datacd, datacdcountdown;
    int slot_head, slot_cur, n_open;
    int rc, lastpacketstate, n_bad, n_trials, out_count, overflow;
    long long bits_seen;
};
struct RtSlot { int fill, next_trial, done, closed, targetSUSize, targetBlocks; long long start_bit; };

struct RtParams {
    int n_channels, oqpsk, ifb, number_of_bits, total_number_of_bits;
    RtState *state;
    RtSlot *slots;          // [ch][RT_SLOTS]
    uint8_t *blocks;        // [ch][RT_SLOTS][RT_BLOCK] soft values in arrival order
    uint8_t *out;           // [ch][RT_OUT][RT_OUT_BYTES]
};

int rt_set_scrambler(const uint8_t *seq);
int rt_init(const RtParams &rp, cudaStream_t st);
int rt_tick(const RtParams &rp, cudaStream_t st);
int rt_out_reset(const RtParams &rp, cudaStream_t st);
int rt_process(const RtParams &rp, const int16_t *d_soft, const int *d_soft_count, size_t soft_stride, cudaStream_t st, long long *launches, int vmode = 0);

} // namespace jb

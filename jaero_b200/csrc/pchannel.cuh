#pragma once
#include <cuda_runtime.h>
#include <cstdint>
namespace jb {

static const int PCHAN_QUEUE_MIN = 4;        // completed blocks a channel may queue within one process() call: at least this many;
                                             // PChanParams::queue is sized at create time from the demodulator's soft-bit ring
                                             // (2 s of soft bits) so that a full ring never overflows it

struct PChanState {                          // members of AeroL used by Decode (JAERO/aerol.h:937-1016)
    unsigned sr_plain, sr_imag, sr_real;     // preamble detector buffers as 32-bit shift registers
    int inv_imag, inv_real, realimag, gotsync_last;
    int cntr, blockcnt;
    unsigned short frameinfo, lastframeinfo;
    int datacdcountdown, datacd;
    int scr_pos, dl2_ptr, info_len, first_decode_done;
    int nframes, blocks_ready, carry_slot, queue_overflow;
    int su_count;                            // SUs waiting in the output ring
    long long su_total, su_ok, bits_seen;
    int dcd_rises;
};
struct PChanBlockMeta { int scr_pos, info_off, n_valid, frame_done, frame_index; };

struct PChanParams {
    int n_channels, oqpsk, cols, block_len, number_of_bits, bits_in_header, total_number_of_bits, paddinglength;
    int dl2_len, info_cap, su_cap;
    int queue;              // block slots per channel (>= PCHAN_QUEUE_MIN)
    PChanState *state;
    uint8_t *blocks;        // [ch][queue][block_len] interleaved soft values
    uint8_t *decoded;       // [ch][queue][block_len/2]
    PChanBlockMeta *meta;   // [ch][queue]
    int *ready;             // [ch] blocks queued by the frame stage
    uint8_t *dl2;           // [ch][dl2_len]
    uint8_t *infofield;     // [ch][info_cap]
    uint8_t *su_out;        // [ch][su_cap][16]: 12 SU bytes, crc_ok, index in frame, frame number (lo,hi)
};

int pchan_set_scrambler(const uint8_t *seq);
int pchan_init(const PChanParams &pp, cudaStream_t st);
int pchan_tick(const PChanParams &pp, int *demod_dcd, cudaStream_t st);
int pchan_process(const PChanParams &pp, const int16_t *d_soft, const int *d_soft_count, int soft_cap, int *demod_dcd,
                  uint8_t *vit_overlap, int *vit_overlap_len, int *vit_renorm, int *vit_valid, int max_queue, cudaStream_t st,
                  long long *launches, const int *lost_n = nullptr, const int *lost_pos = nullptr, size_t lost_pitch = 0);
int pchan_lost(const PChanParams &pp, int channel, int *demod_dcd, cudaStream_t st);
}

// C-channel (8400 bps) frame layer — device data layout and launch prototypes (see cchannel.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace jb {

static const int CC_FRAME_BITS = 4096;     // AERO_SPEC_NumberOfBits (aerol.cpp:1039)
static const int CC_CODED = 5460;          // de-punctured soft values per frame (4095 + 1365 erasures)
static const int CC_CODED_PITCH = 5472;    // row pitch (16-byte multiple)
static const int CC_DEC = 2730;            // decoded bits per frame before resize(2714)
static const int CC_KEEP = 2714;
static const int CC_QUEUE = 4;             // frames that may complete per channel in one call
static const int CC_OUT = 8;               // decoded frames held per channel until read
static const int CC_RECORD = 352;          // 3 x (12 SU bytes + crc flag + 3 pad) + 300 voice bytes + 4 (frame number)

struct CChanState {
    unsigned long long b1_real, b2_real, b1_imag, b2_imag;     // OQPSKPreambleDetectorAndAmbiguityCorrection x2 (52-bit registers)
    int inv_real, inv_imag, realimag, gotsync_last;
    int cntr, index, datacd, datacdcountdown;
    int dl2_ptr, frames_ready, carry_slot, nframes, out_count, overflow;
    long long bits_seen, su_total, su_ok;
};
struct CChanParams {
    int n_channels, dl2_len;
    CChanState *state;
    uint8_t *coded;         // [ch][CC_QUEUE][CC_CODED_PITCH] code-order soft values (erasures = 128 pre-filled)
    uint8_t *decoded;       // [ch][CC_QUEUE][CC_DEC]
    int *ready;             // [ch]
    uint8_t *dl2;           // [ch][dl2_len]
    uint8_t *out;           // [ch][CC_OUT][CC_RECORD]
};

int cchan_set_scrambler(const uint8_t *seq);
int cchan_init(const CChanParams &cp, cudaStream_t st);
int cchan_tick(const CChanParams &cp, int *demod_dcd, cudaStream_t st);
int cchan_out_reset(const CChanParams &cp, cudaStream_t st);
int cchan_process(const CChanParams &cp, const int16_t *d_soft, const int *d_soft_count, size_t soft_stride, int *demod_dcd,
                  uint8_t *vit_overlap, int *vit_overlap_len, int *vit_renorm, int *vit_valid, cudaStream_t st, long long *launches,
                  const int *lost_n = nullptr, const int *lost_pos = nullptr, size_t lost_pitch = 0);
int cchan_lost(const CChanParams &cp, int channel, int *demod_dcd, cudaStream_t st);

} // namespace jb

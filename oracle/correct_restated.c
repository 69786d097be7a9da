/* TEST INFRASTRUCTURE ONLY (oracle).
 *
 * Restatement of quiet/libcorrect's convolutional codec (src/convolutional/
 * {convolutional,encode,decode,history_buffer,error_buffer,metric,lookup,bit}.c)
 * -- an un-vendored dependency of the reference with NO pinned version
 * (ci-linux-build.sh:118-131; JAERO/JAERO.pro:188). The source is not under
 * /root/reference; this file restates the published algorithm and is anchored on
 * the reference's call sites (JAERO/jconvolutionalcodec.cpp:8-18,98,151-201) and
 * on the contract in SURVEY.md App. B. PARITY UNPINNED: no reference test or
 * golden vector covers it; the end-to-end anchor is CRC-16-valid signal units
 * decoded from the reference's own recordings (tests/test_oracle_samples.py).
 *
 * The written spec shared with the CUDA kernel (jaero_b200/csrc/viterbi.cu):
 *  - shift register sr = (sr<<1 | bit) & (2^K-1), newest bit = LSB;
 *    table[sr] bit p = parity(sr & poly[p]); output p=0 is sent first.
 *  - state = low K-1 bits. successor s has predecessors s>>1 ("low", oldest bit 0)
 *    and (s>>1)|2^(K-2) ("high", oldest bit 1); the branch outputs are table[s]
 *    and table[s | 2^(K-1)].
 *  - soft metric: sum_p |soft[p] - (bit_p ? 255 : 0)| ("linear"), uint16 path metrics.
 *  - phases: warm-up (K-1 steps, no history, states reachable from 0 only),
 *    inner (add-compare-select, ties -> low predecessor), tail (last K-1 steps,
 *    zero-input successors only, ties -> high predecessor), flush from state 0.
 *  - history ring of cap = 5K + 15K slices; when len==cap: best state (first
 *    minimum), walk back 5K slices silently, then emit the remaining 15K decisions
 *    oldest-first; renormalise (subtract min) every 65535/(rate*255) steps.
 *  - decoded bit for a slice = history bit of the state on the survivor
 *    (= the bit shifted out = input bit K-1 steps earlier); bits packed MSB-first.
 */
#include "correct.h"
#include <stdlib.h>
#include <string.h>
#include <limits.h>

typedef uint16_t distance_t;
typedef unsigned int shift_register_t;

struct correct_convolutional {
    unsigned int *table;
    size_t rate, order;
    unsigned int numstates;          /* 1<<order */
    /* decoder state */
    int has_init_decode;
    distance_t *errors[2];
    unsigned int err_index;
    /* history buffer */
    unsigned int min_traceback_length, traceback_group_length, cap, num_states_hist;
    shift_register_t highbit;
    uint8_t **history;
    uint8_t *fetched;
    unsigned int index, len, renormalize_counter, renormalize_interval;
    /* bit writer */
    uint8_t *out_bytes; size_t out_cap; size_t out_len; uint8_t wbyte; unsigned int wbyte_len;
};

static unsigned int popcount_u(unsigned int x) { unsigned int c = 0; while (x) { c += x & 1u; x >>= 1; } return c; }

correct_convolutional *correct_convolutional_create(size_t rate, size_t order,
                                                    const correct_convolutional_polynomial_t *poly)
{
    if (order > 8 * sizeof(shift_register_t) || rate < 2) return NULL;
    correct_convolutional *conv = (correct_convolutional *)calloc(1, sizeof(*conv));
    conv->rate = rate; conv->order = order; conv->numstates = 1u << order;
    conv->table = (unsigned int *)malloc(sizeof(unsigned int) * conv->numstates);
    for (shift_register_t i = 0; i < conv->numstates; i++) {
        unsigned int out = 0, mask = 1;
        for (size_t j = 0; j < rate; j++) { if (popcount_u(i & poly[j]) & 1u) out |= mask; mask <<= 1; }
        conv->table[i] = out;
    }
    conv->has_init_decode = 0;
    return conv;
}

void correct_convolutional_destroy(correct_convolutional *conv)
{
    if (!conv) return;
    free(conv->table);
    if (conv->has_init_decode) {
        free(conv->errors[0]); free(conv->errors[1]);
        for (unsigned int i = 0; i < conv->cap; i++) free(conv->history[i]);
        free(conv->history); free(conv->fetched);
    }
    free(conv);
}

/* ---- bit writer (MSB-first packing) ---- */
static void bw_reset(correct_convolutional *c, uint8_t *bytes, size_t cap)
{ c->out_bytes = bytes; c->out_cap = cap; c->out_len = 0; c->wbyte = 0; c->wbyte_len = 0; }
static void bw_write1(correct_convolutional *c, unsigned int bit)
{
    c->wbyte = (uint8_t)((c->wbyte << 1) | (bit & 1u));
    c->wbyte_len++;
    if (c->wbyte_len == 8) { c->out_bytes[c->out_len++] = c->wbyte; c->wbyte = 0; c->wbyte_len = 0; }
}
static size_t bw_finish(correct_convolutional *c)
{
    if (c->wbyte_len) { c->out_bytes[c->out_len++] = (uint8_t)(c->wbyte << (8 - c->wbyte_len)); c->wbyte = 0; c->wbyte_len = 0; }
    return c->out_len;
}

/* ---- encoder ---- */
size_t correct_convolutional_encode_len(correct_convolutional *conv, size_t msg_len)
{
    size_t msgbits = 8 * msg_len;
    return conv->rate * (msgbits + conv->order + 1);
}
size_t correct_convolutional_encode(correct_convolutional *conv, const uint8_t *msg, size_t msg_len, uint8_t *encoded)
{
    shift_register_t sr = 0, mask = (1u << conv->order) - 1u;
    size_t nbits = correct_convolutional_encode_len(conv, msg_len);
    size_t nbytes = (nbits + 7) / 8;
    bw_reset(conv, encoded, nbytes);
    for (size_t i = 0; i < 8 * msg_len + conv->order + 1; i++) {
        unsigned int bit = (i < 8 * msg_len) ? ((msg[i >> 3] >> (7 - (i & 7))) & 1u) : 0u;
        sr = ((sr << 1) | bit) & mask;
        unsigned int out = conv->table[sr];
        for (size_t p = 0; p < conv->rate; p++) { bw_write1(conv, out & 1u); out >>= 1; }
    }
    bw_finish(conv);
    return nbits;
}

/* ---- decoder ---- */
static distance_t soft_distance_linear(unsigned int hard_x, const uint8_t *soft_y, size_t len)
{
    distance_t dist = 0;
    for (size_t i = 0; i < len; i++) {
        unsigned int soft_x = (0u - (hard_x & 1u)) & 0xffu;
        hard_x >>= 1;
        int d = (int)soft_y[i] - (int)soft_x;
        dist = (distance_t)(dist + ((d < 0) ? -d : d));
    }
    return dist;
}
static distance_t hard_distance(unsigned int x, unsigned int y) { return (distance_t)popcount_u(x ^ y); }

static void decode_init(correct_convolutional *conv)
{
    const unsigned int soft_max = 255, distance_max = 65535;
    conv->has_init_decode = 1;
    conv->renormalize_interval = distance_max / (unsigned int)(conv->rate * soft_max);
    conv->min_traceback_length = 5 * (unsigned int)conv->order;
    conv->traceback_group_length = 15 * (unsigned int)conv->order;
    conv->cap = conv->min_traceback_length + conv->traceback_group_length;
    conv->num_states_hist = conv->numstates / 2;
    conv->highbit = 1u << (conv->order - 1);
    conv->history = (uint8_t **)malloc(conv->cap * sizeof(uint8_t *));
    for (unsigned int i = 0; i < conv->cap; i++) conv->history[i] = (uint8_t *)calloc(conv->num_states_hist, 1);
    conv->fetched = (uint8_t *)malloc(conv->cap);
    conv->errors[0] = (distance_t *)calloc(conv->numstates, sizeof(distance_t));
    conv->errors[1] = (distance_t *)calloc(conv->numstates, sizeof(distance_t));
    conv->index = 0; conv->len = 0; conv->renormalize_counter = 0;
}

static shift_register_t hist_search(correct_convolutional *c, const distance_t *d, unsigned int every)
{
    shift_register_t best = 0; distance_t least = USHRT_MAX;
    for (shift_register_t s = 0; s < c->num_states_hist; s += every)
        if (d[s] < least) { least = d[s]; best = s; }
    return best;
}
static shift_register_t hist_renormalize(correct_convolutional *c, distance_t *d, unsigned int every)
{
    shift_register_t best = 0; distance_t mn = d[0];
    for (shift_register_t s = 0; s < c->num_states_hist; s += every)
        if (d[s] < mn) { mn = d[s]; best = s; }
    for (shift_register_t s = 0; s < c->num_states_hist; s += every) d[s] = (distance_t)(d[s] - mn);
    return best;
}
static void hist_traceback(correct_convolutional *c, shift_register_t bestpath, unsigned int min_tb)
{
    unsigned int fetched = 0, index = c->index, cap = c->cap;
    shift_register_t highbit = c->highbit;
    for (unsigned int j = 0; j < min_tb; j++) {
        index = (index == 0) ? cap - 1 : index - 1;
        uint8_t h = c->history[index][bestpath];
        bestpath |= h ? highbit : 0;
        bestpath >>= 1;
    }
    for (unsigned int j = min_tb; j < c->len; j++) {
        index = (index == 0) ? cap - 1 : index - 1;
        uint8_t h = c->history[index][bestpath];
        shift_register_t pathbit = h ? highbit : 0;
        bestpath |= pathbit;
        bestpath >>= 1;
        c->fetched[fetched++] = pathbit ? 1 : 0;
    }
    for (unsigned int j = fetched; j > 0; j--) bw_write1(c, c->fetched[j - 1]);   /* oldest first */
    c->len -= fetched;
}
static void hist_process_skip(correct_convolutional *c, distance_t *d, unsigned int skip)
{
    c->index++; if (c->index == c->cap) c->index = 0;
    c->renormalize_counter++; c->len++;
    if (c->renormalize_counter == c->renormalize_interval) {
        c->renormalize_counter = 0;
        shift_register_t best = hist_renormalize(c, d, skip);
        if (c->len == c->cap) hist_traceback(c, best, c->min_traceback_length);
    } else if (c->len == c->cap) {
        shift_register_t best = hist_search(c, d, skip);
        hist_traceback(c, best, c->min_traceback_length);
    }
}

static ssize_t decode_common(correct_convolutional *conv, size_t num_encoded_bits, uint8_t *msg,
                             const uint8_t *soft, const uint8_t *hard_bytes)
{
    if (num_encoded_bits % conv->rate) return -1;
    if (!conv->has_init_decode) decode_init(conv);
    const unsigned int order = (unsigned int)conv->order, rate = (unsigned int)conv->rate;
    unsigned int sets = (unsigned int)(num_encoded_bits / rate);
    size_t num_encoded_bytes = (num_encoded_bits + 7) / 8;
    bw_reset(conv, msg, num_encoded_bytes);
    memset(conv->errors[0], 0, conv->numstates * sizeof(distance_t));
    memset(conv->errors[1], 0, conv->numstates * sizeof(distance_t));
    conv->err_index = 0;
    conv->len = 0; conv->index = 0;                 /* history_buffer_reset */
    distance_t *rd = conv->errors[0], *wr = conv->errors[1], *t;
    const unsigned int *table = conv->table;
    distance_t distances[16];
    size_t hard_pos = 0;
#define HARD_OUT(o) do { o = 0; for (unsigned int p_ = 0; p_ < rate; p_++) { \
        unsigned int b_ = (hard_bytes[hard_pos >> 3] >> (7 - (hard_pos & 7))) & 1u; hard_pos++; o |= b_ << p_; } } while (0)
    /* warm-up: load the shift register, K-1 steps, only states reachable from 0 */
    for (unsigned int i = 0; i < order - 1 && i < sets; i++) {
        unsigned int out = 0;
        if (!soft) HARD_OUT(out);
        for (unsigned int j = 0; j < (1u << (i + 1)); j++) {
            unsigned int last = j >> 1;
            distance_t dist = soft ? soft_distance_linear(table[j], soft + i * rate, rate) : hard_distance(table[j], out);
            wr[j] = (distance_t)(dist + rd[last]);
        }
        t = rd; rd = wr; wr = t;
    }
    /* inner: full add-compare-select */
    shift_register_t highbit = 1u << (order - 1);
    for (unsigned int i = order - 1; i + order - 1 < sets; i++) {
        unsigned int out = 0;
        if (!soft) HARD_OUT(out);
        for (unsigned int j = 0; j < (1u << rate); j++)
            distances[j] = soft ? soft_distance_linear(j, soft + i * rate, rate) : hard_distance(j, out);
        uint8_t *history = conv->history[conv->index];
        for (shift_register_t s = 0; s < highbit; s++) {
            distance_t low_err = (distance_t)(distances[table[s]] + rd[s >> 1]);
            distance_t high_err = (distance_t)(distances[table[s | highbit]] + rd[(s >> 1) | (highbit >> 1)]);
            if (low_err <= high_err) { wr[s] = low_err; history[s] = 0; }
            else { wr[s] = high_err; history[s] = 1; }
        }
        hist_process_skip(conv, wr, 1);
        t = rd; rd = wr; wr = t;
    }
    /* tail: only zero-input successors */
    for (unsigned int i = (sets >= order - 1 ? sets - (order - 1) : 0); i < sets; i++) {
        if (i < order - 1) continue;                 /* degenerate tiny inputs */
        unsigned int out = 0;
        if (!soft) HARD_OUT(out);
        for (unsigned int j = 0; j < (1u << rate); j++)
            distances[j] = soft ? soft_distance_linear(j, soft + i * rate, rate) : hard_distance(j, out);
        uint8_t *history = conv->history[conv->index];
        unsigned int skip = 1u << (order - (sets - i));
        for (shift_register_t s = 0; s < highbit; s += skip) {
            distance_t low_err = (distance_t)(distances[table[s]] + rd[s >> 1]);
            distance_t high_err = (distance_t)(distances[table[s | highbit]] + rd[(s >> 1) | (highbit >> 1)]);
            if (low_err < high_err) { wr[s] = low_err; history[s] = 0; }
            else { wr[s] = high_err; history[s] = 1; }
        }
        hist_process_skip(conv, wr, skip);
        t = rd; rd = wr; wr = t;
    }
#undef HARD_OUT
    hist_traceback(conv, 0, 0);                      /* flush from state 0 */
    return (ssize_t)bw_finish(conv);                 /* bytes written */
}

ssize_t correct_convolutional_decode_soft(correct_convolutional *conv, const correct_convolutional_soft_t *encoded,
                                          size_t num_encoded_bits, uint8_t *msg)
{ return decode_common(conv, num_encoded_bits, msg, encoded, NULL); }

ssize_t correct_convolutional_decode(correct_convolutional *conv, const uint8_t *encoded,
                                     size_t num_encoded_bits, uint8_t *msg)
{ return decode_common(conv, num_encoded_bits, msg, NULL, encoded); }

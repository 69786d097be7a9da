// TEST INFRASTRUCTURE: a stand-in for the entry points of libjaero_b200.so that include/jaero_b200_host.hpp calls, with a
// trivially predictable "demodulator" (every (int)(Fs/fb)-th PCM sample becomes one soft value = sample & 0xff; a burst handle prefixes
// a -1 marker once). Lets the host mirror's plumbing (chunking, group-of-32/12 emits, setters, failure path) be checked
// on a machine without a GPU. Not linked into anything shipped.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "jaero_b200.h"

struct Fake {
    jaero_settings s; std::vector<int16_t> soft; long long samples; int dcd, afc, sql, cpu, burst, marker_done; double center;
};
struct jaero_batch { Fake f; };
struct jaero_burst { Fake f; };
struct jaero_viterbi { int pad; };
static std::string g_err;
int mock_fail_create = 0;            // set by the test through mock_set_fail()
int mock_log[8];                     // [0] creates [1] destroys [2] writes [3] largest write [4] afc calls [5] sql [6] cpu [7] dcd

extern "C" {
void mock_set_fail(int v) { mock_fail_create = v; }
int *mock_get_log(void) { return mock_log; }
const char *jaero_last_error(void) { return g_err.c_str(); }

static int create(Fake &f, const jaero_settings *s, int burst)
{
    memset(&f.s, 0, sizeof f.s); f.s = *s; f.samples = 0; f.dcd = 0; f.afc = s->afc; f.sql = s->sql; f.cpu = s->cpu_reduce; f.burst = burst;
    f.marker_done = 0; f.center = s->freq_center;
    mock_log[0]++;
    return JAERO_OK;
}
static int write(Fake &f, const int16_t *pcm, size_t n)
{
    mock_log[2]++; if ((int)n > mock_log[3]) mock_log[3] = (int)n;
    for (size_t i = 0; i < n; i++, f.samples++) {
        if (f.burst && !f.marker_done && f.samples == 10) { f.soft.push_back(-1); f.marker_done = 1; }
        const int step = (int)(f.s.Fs / f.s.fb) > 0 ? (int)(f.s.Fs / f.s.fb) : 1;
        if (f.samples % step == 0) f.soft.push_back((int16_t)(pcm[i] & 0xff));
    }
    return JAERO_OK;
}
static int read(Fake &f, int16_t *out, size_t cap, int32_t *counts)
{
    if (f.soft.size() > cap) { g_err = "soft-bit ring overflow"; return JAERO_E_OVERFLOW; }
    memcpy(out, f.soft.data(), f.soft.size() * 2); counts[0] = (int32_t)f.soft.size(); f.soft.clear();
    return JAERO_OK;
}
int jaero_batch_create(const jaero_settings *s, int n, const double *, int, jaero_batch **out)
{
    if (mock_fail_create || n != 1) { g_err = "CUDA error 100 (no CUDA-capable device is detected) [mock]"; return JAERO_E_CUDA; }
    *out = new jaero_batch(); return create((*out)->f, s, 0);
}
void jaero_batch_destroy(jaero_batch *b) { mock_log[1]++; delete b; }
int jaero_batch_write(jaero_batch *b, const int16_t *p, size_t n, size_t) { return write(b->f, p, n); }
int jaero_batch_read_softbits(jaero_batch *b, int16_t *o, size_t cap, int32_t *c) { return read(b->f, o, cap, c); }
int jaero_batch_set_dcd(jaero_batch *b, int, int d) { b->f.dcd = d; mock_log[7]++; return JAERO_OK; }
int jaero_batch_set_center_freq(jaero_batch *b, int, double hz) { b->f.center = hz; return JAERO_OK; }
int jaero_batch_set_afc(jaero_batch *b, int v) { b->f.afc = v; mock_log[4]++; return JAERO_OK; }
int jaero_batch_set_sql(jaero_batch *b, int v) { b->f.sql = v; mock_log[5]++; return JAERO_OK; }
int jaero_batch_set_cpu_reduce(jaero_batch *b, int v) { b->f.cpu = v; mock_log[6]++; return JAERO_OK; }
int jaero_batch_get_status(jaero_batch *b, int, jaero_status *st)
{
    memset(st, 0, sizeof *st);
    st->center_freq = b->f.center; st->mixer2_freq = b->f.center + 1.5; st->mse = b->f.samples >= 1000 ? 0.1 : 0.9; st->ebno = 12.5;
    st->samples = b->f.samples; st->dcd = b->f.dcd;
    return JAERO_OK;
}
int jaero_burst_msk_create(const jaero_settings *s, int n, int, jaero_burst **out)
{
    if (mock_fail_create || n != 1) { g_err = "CUDA error 100 [mock]"; return JAERO_E_CUDA; }
    *out = new jaero_burst(); return create((*out)->f, s, 1);
}
int jaero_burst_oqpsk_create(const jaero_settings *s, int n, int d, jaero_burst **out) { return jaero_burst_msk_create(s, n, d, out); }
void jaero_burst_destroy(jaero_burst *b) { mock_log[1]++; delete b; }
int jaero_burst_write(jaero_burst *b, const int16_t *p, size_t n, size_t) { return write(b->f, p, n); }
int jaero_burst_read_softbits(jaero_burst *b, int16_t *o, size_t cap, int32_t *c) { return read(b->f, o, cap, c); }
int jaero_burst_set_dcd(jaero_burst *b, int, int d) { b->f.dcd = d; mock_log[7]++; return JAERO_OK; }
int jaero_burst_set_afc(jaero_burst *b, int v) { b->f.afc = v; mock_log[4]++; return JAERO_OK; }
int jaero_burst_set_sql(jaero_burst *b, int v) { b->f.sql = v; mock_log[5]++; return JAERO_OK; }
int jaero_burst_get_status_all(jaero_burst *b, jaero_burst_status *st)
{
    memset(st, 0, sizeof *st); st->center_freq = b->f.center; st->mixer2_freq = b->f.center; st->mse = 0.2; st->ebno = 7.0;
    return JAERO_OK;
}
int jaero_viterbi_create(int, int pad, int, jaero_viterbi **out)
{
    if (mock_fail_create) { g_err = "CUDA error 100 [mock]"; return JAERO_E_CUDA; }
    *out = new jaero_viterbi(); (*out)->pad = pad; return JAERO_OK;
}
void jaero_viterbi_destroy(jaero_viterbi *v) { delete v; }
int jaero_viterbi_decode_continuous(jaero_viterbi *, const uint8_t *soft, size_t n, int, uint8_t *bits, int32_t *nv)
{
    for (size_t i = 0; i < n / 2; i++) bits[i] = soft[2 * i] > 127;
    *nv = (int32_t)(n / 2) - 13;
    return JAERO_OK;
}
int jaero_viterbi_decode_block(jaero_viterbi *, const uint8_t *soft, size_t n, uint8_t *bits)
{
    for (size_t i = 0; i < n / 2; i++) bits[i] = soft[2 * i + 1] > 127;
    return JAERO_OK;
}
}

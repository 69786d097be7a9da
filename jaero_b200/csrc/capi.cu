// extern "C" boundary of libjaero_b200.so (declared in include/jaero_b200.h).
// Host-side object management only: device allocation, staging copies, stream ordering, kernel
// launches. No CPU implementation of any DSP lives here — without a CUDA device every create
// call fails loudly.
#include "../../include/jaero_b200.h"
#include "common.cuh"
#include "viterbi.cuh"
#include "demod.cuh"
#include "prefilter.cuh"
#include <cstring>
#include <complex>
#include <cstdlib>
#include <new>
#include <vector>

namespace jb {
static thread_local std::string g_err;
void set_error(const std::string &m) { g_err = m; }
int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_err = buf;
    return JAERO_E_CUDA;
}
} // namespace jb
using namespace jb;

// A create call that fails half-way (any JB_CUDA early return) hands the partly built object to its destroy function.
template <class T> struct CreateGuard {
    T *obj; void (*destroy)(T *);
    CreateGuard(T *o, void (*d)(T *)) : obj(o), destroy(d) {}
    ~CreateGuard() { if (obj) destroy(obj); }
    void release() { obj = 0; }
    CreateGuard(const CreateGuard &) = delete;
    CreateGuard &operator=(const CreateGuard &) = delete;
};

struct jaero_viterbi {
    int n_channels, pad, device;
    cudaStream_t stream;
    uint8_t *d_overlap; int *d_overlap_len; int *d_renorm;
    uint8_t *d_soft, *d_bits; size_t soft_cap, bits_cap;
    int *d_valid;
    int64_t launches;
};

extern "C" {

const char *jaero_last_error(void) { return g_err.c_str(); }
int jaero_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ------------------------------------------------------------------ Viterbi
int jaero_viterbi_create(int n_channels, int paddinglength, int device, jaero_viterbi **out)
{
    if (!out || n_channels <= 0 || paddinglength < 0 || (paddinglength & 1)) { set_error("jaero_viterbi_create: bad argument"); return JAERO_E_ARG; }
    int ndev = 0;
    JB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("jaero_viterbi_create: no such CUDA device"); return JAERO_E_CUDA; }
    JB_CUDA(cudaSetDevice(device));
    jaero_viterbi *v = new (std::nothrow) jaero_viterbi();
    if (!v) { set_error("out of host memory"); return JAERO_E_ARG; }
    CreateGuard<jaero_viterbi> guard(v, jaero_viterbi_destroy);
    memset(v, 0, sizeof *v);
    v->n_channels = n_channels; v->pad = paddinglength; v->device = device;
    JB_CUDA(cudaStreamCreateWithFlags(&v->stream, cudaStreamNonBlocking));
    JB_CUDA(cudaMalloc(&v->d_overlap, (size_t)n_channels * 64));
    JB_CUDA(cudaMalloc(&v->d_overlap_len, (size_t)n_channels * sizeof(int)));
    JB_CUDA(cudaMalloc(&v->d_renorm, (size_t)n_channels * sizeof(int)));
    JB_CUDA(cudaMalloc(&v->d_valid, (size_t)n_channels * sizeof(int)));
    JB_CUDA(cudaMemsetAsync(v->d_overlap, 0, (size_t)n_channels * 64, v->stream));
    JB_CUDA(cudaMemsetAsync(v->d_overlap_len, 0, (size_t)n_channels * sizeof(int), v->stream));
    JB_CUDA(cudaMemsetAsync(v->d_renorm, 0, (size_t)n_channels * sizeof(int), v->stream));
    guard.release();
    *out = v;
    return JAERO_OK;
}
void jaero_viterbi_destroy(jaero_viterbi *v)
{
    if (!v) return;
    cudaSetDevice(v->device);
    cudaStreamSynchronize(v->stream);
    cudaFree(v->d_overlap); cudaFree(v->d_overlap_len); cudaFree(v->d_renorm); cudaFree(v->d_valid); cudaFree(v->d_soft); cudaFree(v->d_bits);
    cudaStreamDestroy(v->stream);
    delete v;
}
int jaero_viterbi_reset(jaero_viterbi *v)
{
    if (!v) { set_error("null handle"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(v->device));
    JB_CUDA(cudaMemsetAsync(v->d_overlap_len, 0, (size_t)v->n_channels * sizeof(int), v->stream));
    return JAERO_OK;
}
int jaero_viterbi_sync(jaero_viterbi *v)
{
    if (!v) { set_error("null handle"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(v->device));
    JB_CUDA(cudaStreamSynchronize(v->stream));
    return JAERO_OK;
}
int64_t jaero_viterbi_launch_count(const jaero_viterbi *v) { return v ? v->launches : 0; }

static int vit_check(jaero_viterbi *v, size_t n_soft, int cols)
{
    if (!v) { set_error("null handle"); return JAERO_E_ARG; }
    if (n_soft < 32 || (n_soft & 1) || n_soft > 60000) { set_error("viterbi: n_soft must be even, 32..60000"); return JAERO_E_ARG; }
    if (cols < 0 || (cols > 0 && (size_t)cols * 64 != n_soft)) { set_error("viterbi: interleaver_cols*64 must equal n_soft"); return JAERO_E_ARG; }
    return JAERO_OK;
}
int jaero_viterbi_decode_continuous_device(jaero_viterbi *v, const uint8_t *d_soft, size_t n_soft, int cols, uint8_t *d_bits, int32_t *d_n_valid)
{
    int r = vit_check(v, n_soft, cols); if (r) return r;
    JB_CUDA(cudaSetDevice(v->device));
    if (viterbi_launch(d_soft, (int)n_soft, cols, 0, v->pad, v->d_overlap, v->d_overlap_len, v->d_renorm, d_bits, d_n_valid, v->n_channels, v->stream)) return JAERO_E_CUDA;
    v->launches++;
    return JAERO_OK;
}
static int vit_stage(jaero_viterbi *v, size_t n_soft)
{
    size_t need = (size_t)v->n_channels * n_soft;
    if (need > v->soft_cap) { cudaFree(v->d_soft); v->d_soft = 0; JB_CUDA(cudaMalloc(&v->d_soft, need)); v->soft_cap = need; }
    size_t needb = (size_t)v->n_channels * (n_soft / 2);
    if (needb > v->bits_cap) { cudaFree(v->d_bits); v->d_bits = 0; JB_CUDA(cudaMalloc(&v->d_bits, needb)); v->bits_cap = needb; }
    return JAERO_OK;
}
int jaero_viterbi_decode_continuous(jaero_viterbi *v, const uint8_t *soft, size_t n_soft, int cols, uint8_t *bits_out, int32_t *n_valid)
{
    int r = vit_check(v, n_soft, cols); if (r) return r;
    if (!soft || !bits_out) { set_error("null buffer"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(v->device));
    r = vit_stage(v, n_soft); if (r) return r;
    JB_CUDA(cudaMemcpyAsync(v->d_soft, soft, (size_t)v->n_channels * n_soft, cudaMemcpyHostToDevice, v->stream));
    r = jaero_viterbi_decode_continuous_device(v, v->d_soft, n_soft, cols, v->d_bits, v->d_valid); if (r) return r;
    JB_CUDA(cudaMemcpyAsync(bits_out, v->d_bits, (size_t)v->n_channels * (n_soft / 2), cudaMemcpyDeviceToHost, v->stream));
    if (n_valid) JB_CUDA(cudaMemcpyAsync(n_valid, v->d_valid, (size_t)v->n_channels * sizeof(int), cudaMemcpyDeviceToHost, v->stream));
    JB_CUDA(cudaStreamSynchronize(v->stream));
    return JAERO_OK;
}
int jaero_viterbi_decode_block(jaero_viterbi *v, const uint8_t *soft, size_t n_soft, uint8_t *bits_out)
{
    int r = vit_check(v, n_soft, 0); if (r) return r;
    if (!soft || !bits_out) { set_error("null buffer"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(v->device));
    r = vit_stage(v, n_soft); if (r) return r;
    JB_CUDA(cudaMemcpyAsync(v->d_soft, soft, (size_t)v->n_channels * n_soft, cudaMemcpyHostToDevice, v->stream));
    if (viterbi_launch(v->d_soft, (int)n_soft, 0, 1, 0, v->d_overlap, v->d_overlap_len, v->d_renorm, v->d_bits, v->d_valid, v->n_channels, v->stream)) return JAERO_E_CUDA;
    v->launches++;
    JB_CUDA(cudaMemcpyAsync(bits_out, v->d_bits, (size_t)v->n_channels * (n_soft / 2), cudaMemcpyDeviceToHost, v->stream));
    JB_CUDA(cudaStreamSynchronize(v->stream));
    return JAERO_OK;
}

} // extern "C"

// ====================================================================== demodulator batches
#include <cmath>
#include <algorithm>

namespace {

// RootRaisedCosine::design (JAERO/DSP.h:316-338): closed-form RRC taps, firsize forced odd.
std::vector<double> rrc_taps(double alpha, int firsize, double samplerate, double symbol_freq)
{
    if ((firsize % 2) == 0) firsize += 1;
    std::vector<double> pts(firsize);
    const double T = (samplerate) / (symbol_freq);
    for (int i = 0; i < firsize; i++) {
        if (i == ((firsize - 1) / 2)) pts[i] = (4.0 * alpha + M_PI - M_PI * alpha) / (M_PI * sqrt(T));
        else {
            const double fi = (((double)i) - ((double)(firsize - 1)) / 2.0);
            if (fabs(1.0 - pow(4.0 * alpha * fi / T, 2)) < 0.0000000001)
                pts[i] = (alpha * ((M_PI - 2.0) * cos(M_PI / (4.0 * alpha)) + (M_PI + 2.0) * sin(M_PI / (4.0 * alpha))) / (M_PI * sqrt(2.0 * T)));
            else
                pts[i] = (4.0 * alpha / (M_PI * sqrt(T)) * (cos((1.0 + alpha) * M_PI * fi / T) + T / (4.0 * alpha * fi) * sin((1.0 - alpha) * M_PI * fi / T)) / (1.0 - pow(4.0 * alpha * fi / T, 2)));
        }
    }
    return pts;
}

// Delay<T>::update interpolation weight (JAERO/DSP.h:357-374) for every ring position. The kernels keep the delay line
// as a shift register; the weight the reference derives from (buffptr - fractdelay) can differ in the last bit between
// ring positions, so it is tabulated per position and indexed by the lock-step sample count.
bool delay_weights(double fractdelay, int *k_out, double *w_out /*[4]*/)
{
    const int size = (int)std::ceil(fractdelay) + 1;
    if (size > 4 || size < 2) return false;
    for (int bp = 0; bp < size; bp++) {
        double dptr = ((double)bp) - fractdelay;
        while (std::floor(dptr) < 0) dptr += ((double)size);
        const int iptr = (int)std::floor(dptr);
        w_out[bp] = dptr - ((double)iptr);
        // the shift-register form needs the read position to be "ceil(fd) samples ago" at every ring position
        int expect = bp - (int)std::ceil(fractdelay); while (expect < 0) expect += size;
        if (iptr != expect) return false;
    }
    *k_out = (int)std::ceil(fractdelay);
    return true;
}

template <class T> int dev_alloc_zero(T **p, size_t count, cudaStream_t s)
{
    JB_CUDA(cudaMalloc((void **)p, count * sizeof(T)));
    JB_CUDA(cudaMemsetAsync(*p, 0, count * sizeof(T), s));
    return 0;
}
} // namespace

struct jaero_batch {
    jaero_settings set;
    int device;
    cudaStream_t stream;
    DemodParams p;
    CfePlan cfe;
    std::vector<void *> allocs;
    // lock-step counters mirrored on the host
    long long samples;          // samples fully processed
    int bb_pos, coarse_counter;
    int16_t *d_stage; size_t stage_cap;
    // 8400 bps pre-filter (K6)
    bool pre_on; PreParams pre; FirStream fir; int fir_fill; long long fir_blocks; double2 *d_x; size_t x_cap;
    int16_t *h_soft_stage;      // pinned
    int *h_ints; double *h_dbls; long long *h_soft_total;   // pinned mirrors of I / D / soft_total
    long long launches;
    cudaStream_t own_stream;
    bool profiling;
    // asynchronous coarse estimator (see jaero_batch_write_device)
    bool async_cfe; cudaStream_t cfe_stream; cudaEvent_t ev_seg_done, ev_cfe_done[2]; int cfe_count; int bb_phys;
    // host-input pipelining (jaero_batch_write): the H2D copy is cut into column slices on a copy stream; a segment only
    // waits for the slices it reads
    cudaStream_t copy_stream; cudaEvent_t ev_slice[8], ev_stage_free; int n_slices, slice_len, next_slice;
    bool use_pipe;              // 10500 bps: warp-specialised segment kernel (JAERO_OQPSK_PIPE=0 selects the single-warp one, for A/B profiling)
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_seg, ev_cfe;
    double prof_samples;
    int trace_state;            // JAERO_PIPE_TRACE: 0 = armed, 1 = done
    // seating of the channels in the pipelined 10500 bps kernel (regroup)
    int *d_chan_of; double *d_keys, *h_keys; double *d_ring_scratch; size_t ring_scratch_count;
    std::vector<int> slot_of;   // [cpad] seat of channel c
    long long epochs, next_regroup; int regroup_every; long long regroups;
};

namespace {
template <class T> int batch_alloc(jaero_batch *b, T **p, size_t count)
{
    int r = dev_alloc_zero(p, count, b->stream);
    if (r == 0) b->allocs.push_back((void *)*p);
    return r;
}

__global__ void init_state_kernel(DemodParams p, const double *freq_center, double st_freq, double ebno_init)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_channels) return;
    auto D = [&](int i) -> double & { return p.D[(size_t)i * p.cpad + ch]; };
    auto I = [&](int i) -> int & { return p.I[(size_t)i * p.cpad + ch]; };
    // WaveTable::SetFreq(double,int) (DSP.cpp:142-149): WTstep = freq*WTSIZE/(float)samplerate
    double fc = freq_center[ch];
    if (fc > ((p.Fs / 2.0) - (p.lockingbw / 2.0))) fc = ((p.Fs / 2.0) - (p.lockingbw / 2.0));   // oqpskdemodulator.cpp:183
    if (fc < 0) fc = 0;
    const double sr = (double)((float)((int)p.Fs));
    D(D_M2_FREQ) = fc; D(D_M2_STEP) = (fc) * ((double)jb::WTSIZE) / sr;
    D(D_MC_FREQ) = fc; D(D_MC_STEP) = (fc) * ((double)jb::WTSIZE) / sr;
    D(D_ST_FREQ) = st_freq; D(D_ST_STEP) = (st_freq) * ((double)jb::WTSIZE) / sr;
    D(D_SR_FREQ) = st_freq; D(D_SR_STEP) = (st_freq) * ((double)jb::WTSIZE) / sr;
    D(D_MSE) = (p.kind == JAERO_KIND_OQPSK) ? 100.0 : 10.0;       // oqpskdemodulator.cpp:17 / mskdemodulator.cpp:180
    D(D_DIFF_LAST) = -1.0;                                        // DSP.cpp:520
    D(D_EB_EBNO) = ebno_init;
    I(I_COUNTDOWN) = 4; I(I_COUNTDOWN2) = 5;                      // oqpskdemodulator.cpp:641,652 / mskdemodulator.cpp:493
    I(I_EMPTYING) = 1;                                            // coarsefreqestimate.cpp:24
}

// after a host read: move the not-yet-emitted (<32 / <12) soft bits to the front of each ring
__global__ void soft_reset_kernel(DemodParams p)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_channels) return;
    int &count = p.I[(size_t)I_SOFT_COUNT * p.cpad + ch];
    const int pending = p.I[(size_t)I_SOFT_PENDING * p.cpad + ch];
    int16_t *ring = p.soft + (size_t)ch * p.soft_cap;
    for (int k = 0; k < pending; k++) ring[k] = ring[count + k];
    p.soft_total[ch] += count;
    count = 0;
    p.I[(size_t)I_LOST_N * p.cpad + ch] = 0;             // the frame layer consumed the events before the ring is reset
}
__global__ void set_int_kernel(DemodParams p, int idx, int channel, int value)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_channels) return;
    if (channel < 0 || channel == ch) p.I[(size_t)idx * p.cpad + ch] = value;
}
// PeakVolume (oqpskdemodulator.cpp:393-405, mskdemodulator.cpp:329-344): max |sample| of the input since the last read-out.
// One warp per channel row, 16-byte loads; the same for every demodulator kernel variant.
__global__ void peak_kernel(DemodParams p, const int16_t *__restrict__ pcm, size_t stride, int n)
{
    const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (ch >= p.n_channels) return;
    const int4 *row = reinterpret_cast<const int4 *>(pcm + (size_t)ch * stride);
    int m = 0;
    for (int k = lane; k * 8 < n; k += 32) {
        const int4 v = __ldg(row + k);
        const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int lo = (int)(short)(w[q] & 0xffff), hi = w[q] >> 16;
            if (k * 8 + 2 * q < n) m = max(m, abs(lo));
            if (k * 8 + 2 * q + 1 < n) m = max(m, abs(hi));
        }
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if (lane == 0) { int &pk = p.I[(size_t)I_PEAK * p.cpad + ch]; pk = max(pk, m); }
}
// CenterFreqChangedSlot (oqpskdemodulator.cpp:291-310 / mskdemodulator.cpp:265-282)
__global__ void center_freq_kernel(DemodParams p, int channel, double freq_center)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_channels || (channel >= 0 && channel != ch)) return;
    auto D = [&](int i) -> double & { return p.D[(size_t)i * p.cpad + ch]; };
    double fc = freq_center;
    if (p.kind == JAERO_KIND_OQPSK) {
        if (p.fb != 8400) { if (fc < (0.5 * p.fb)) fc = 0.5 * p.fb; if (fc > (p.Fs / 2.0 - 0.5 * p.fb)) fc = p.Fs / 2.0 - 0.5 * p.fb; }
    } else { if (fc < (0.75 * p.fb)) fc = 0.75 * p.fb; if (fc > (p.Fs / 2.0 - 0.75 * p.fb)) fc = p.Fs / 2.0 - 0.75 * p.fb; }
    if (fc < 0) fc = 0;
    const double srf = (double)((float)((int)p.Fs));
    D(D_MC_FREQ) = fc; D(D_MC_STEP) = (fc) * ((double)jb::WTSIZE) / srf;   // SetFreq(freq,Fs)
    auto set_m2 = [&](double f) { if (f < 0) f = 0; D(D_M2_FREQ) = f; D(D_M2_STEP) = (f) * ((double)jb::WTSIZE) / p.Fs; };
    if (p.afc) set_m2(D(D_MC_FREQ));
    if ((D(D_M2_FREQ) - D(D_MC_FREQ)) > (p.lockingbw / 2.0)) set_m2(D(D_MC_FREQ) + (p.lockingbw / 2.0));
    if ((D(D_M2_FREQ) - D(D_MC_FREQ)) < (-p.lockingbw / 2.0)) set_m2(D(D_MC_FREQ) - (p.lockingbw / 2.0));
    double2 *row = p.bb + (size_t)ch * p.bb_len;
    for (int j = 0; j < p.bb_len; j++) row[j] = make_double2(0.0, 0.0);
}
// ---- seating by symbol-timing phase (pipelined 10500 bps kernel)
// key[c] = samples until channel c's next carrier-update strobe, in [0, 2 * samples per strobe): st_osc passes the point ee
// (oqpskdemodulator.cpp:488) every Fs/fb samples and every second passage (yui, :496-503) is a carrier update.
__global__ void regroup_key_kernel(DemodParams p, double *__restrict__ key)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_channels) return;
    const double ptr = p.D[(size_t)D_ST_PTR * p.cpad + ch], step = p.D[(size_t)D_ST_STEP * p.cpad + ch];
    const int yui = p.I[(size_t)I_YUI * p.cpad + ch];
    const double N = (double)jb::WTSIZE;
    double d = p.ee * N - ptr; if (d < 0) d += N;
    const double per = step > 0 ? N / step : 1.0;
    double k = step > 0 ? d / step : 0.0;
    if (yui) k += per;                                    // the next passage only stores pt_d; the one after it updates the carrier
    key[ch] = fmod(k, 2.0 * per);
}
// ring_new[(cta', k, lane')] = ring_old[(cta, k, lane)] for the channel that moves from seat (cta, lane) to (cta', lane')
__global__ void regroup_ring_kernel(const double *__restrict__ src, double *__restrict__ dst, const int *__restrict__ old_seat_of_new, int len, int n_ctas)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);      // (cta', k)
    if (row >= (long long)n_ctas * len) return;
    const int cta_n = (int)(row / len), k = (int)(row % len);
    const int so = old_seat_of_new[cta_n * 32 + lane];
    dst[row * 32 + lane] = src[((size_t)(so >> 5) * len + k) * 32 + (so & 31)];
}
// New seating: slot_of[c] for every channel (pads keep their seats). The rings follow; everything else is indexed by channel.
int batch_apply_seating(jaero_batch *b, const std::vector<int> &new_slot_of)
{
    DemodParams &p = b->p;
    const int cp = p.cpad, n_ctas = cp / 32;
    std::vector<int> old_seat_of_new(cp), chan_of(cp);
    for (int c = 0; c < cp; c++) { old_seat_of_new[new_slot_of[c]] = b->slot_of[c]; chan_of[new_slot_of[c]] = c; }
    const size_t need = (size_t)p.agc_len * cp;
    if (!b->d_ring_scratch) {
        if (cudaMalloc(&b->d_ring_scratch, need * sizeof(double)) != cudaSuccess) { cudaGetLastError(); b->regroup_every = 0; return 0; }   // no room: keep the seating
        b->ring_scratch_count = need;
    }
    int *d_map = b->d_chan_of + cp;                          // second half of the allocation: old seat of each new seat
    JB_CUDA(cudaMemcpyAsync(d_map, old_seat_of_new.data(), cp * sizeof(int), cudaMemcpyHostToDevice, b->stream));
    auto move = [&](double *ring, int len) -> int {
        if (!ring) return 0;
        const long long rows = (long long)n_ctas * len;
        regroup_ring_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, b->stream>>>(ring, b->d_ring_scratch, d_map, len, n_ctas);
        JB_CUDA(cudaGetLastError());
        JB_CUDA(cudaMemcpyAsync(ring, b->d_ring_scratch, (size_t)rows * 32 * sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
        b->launches++;
        return 0;
    };
    if (move(p.agc_ring, p.agc_len) || move(p.ebno_e1, p.ebno_len) || move(p.ebno_e2, p.ebno_len)) return -1;
    JB_CUDA(cudaMemcpyAsync(b->d_chan_of, chan_of.data(), cp * sizeof(int), cudaMemcpyHostToDevice, b->stream));
    JB_CUDA(cudaStreamSynchronize(b->stream));               // the host vectors above go out of scope
    b->slot_of = new_slot_of;
    b->regroups++;
    return 0;
}
int batch_regroup_by_phase(jaero_batch *b, bool force)
{
    DemodParams &p = b->p;
    const int C = p.n_channels, cp = p.cpad;
    regroup_key_kernel<<<(C + 127) / 128, 128, 0, b->stream>>>(p, b->d_keys);
    JB_CUDA(cudaGetLastError());
    JB_CUDA(cudaMemcpyAsync(b->h_keys, b->d_keys, C * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    JB_CUDA(cudaStreamSynchronize(b->stream));
    b->launches++;
    // Is the present seating still coherent? A CTA is coherent when the carrier-update strobes of its channels fall within 1.5
    // samples of each other (circularly, period = two strobe intervals). Moving the rings costs ~30 ms per 4096 channels, so the
    // seating is only changed when more than a quarter of the CTAs have drifted apart (never, for transmitters on one clock).
    if (!force) {
        const double period = 2.0 * p.Fs / p.fb;                                     // keys are in [0, 2*Fs/fb)
        static const double thr = getenv("JAERO_REGROUP_SPREAD") ? atof(getenv("JAERO_REGROUP_SPREAD")) : 1.5;
        int bad = 0, ctas = 0;
        std::vector<double> ks, spreads;
        std::vector<std::vector<int>> members(cp / 32);
        for (int c = 0; c < C; c++) members[b->slot_of[c] >> 5].push_back(c);
        for (auto &m : members) {
            if (m.size() < 2) continue;
            ctas++;
            ks.clear();
            for (int c : m) ks.push_back(b->h_keys[c]);
            std::sort(ks.begin(), ks.end());
            double gap = ks.front() + period - ks.back();
            for (size_t i = 1; i < ks.size(); i++) gap = std::max(gap, ks[i] - ks[i - 1]);
            spreads.push_back(period - gap);
            if (period - gap > thr) bad++;
        }
        if (getenv("JAERO_DEBUG") && !spreads.empty()) {
            std::sort(spreads.begin(), spreads.end());
            fprintf(stderr, "[jaero_b200] seating check at epoch %lld: %d of %d CTAs spread > %.1f samples (median %.2f, p90 %.2f, max %.2f)\n", b->epochs, bad, ctas, thr,
                    spreads[spreads.size() / 2], spreads[spreads.size() * 9 / 10], spreads.back());
        }
        if (bad * 4 <= ctas) return 0;
    }
    std::vector<int> order(C);
    for (int c = 0; c < C; c++) order[c] = c;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return b->h_keys[x] < b->h_keys[y]; });
    std::vector<int> slot_of(cp);
    for (int k = 0; k < C; k++) slot_of[order[k]] = k;
    for (int c = C; c < cp; c++) slot_of[c] = c;
    if (slot_of == b->slot_of) return 0;
    return batch_apply_seating(b, slot_of);
}
} // namespace

extern "C" {

int jaero_batch_create(const jaero_settings *s, int n_channels, const double *freq_center_per_channel, int device, jaero_batch **out)
{
    if (!s || !out || n_channels <= 0) { set_error("jaero_batch_create: bad argument"); return JAERO_E_ARG; }
    if (s->kind != JAERO_KIND_OQPSK && s->kind != JAERO_KIND_MSK) { set_error("jaero_batch_create: unknown kind"); return JAERO_E_ARG; }
    if (s->Fs <= 0 || s->fb <= 0 || s->coarsefreqest_fft_power < 10 || s->coarsefreqest_fft_power > 14) {
        set_error("jaero_batch_create: Fs/fb must be positive and coarsefreqest_fft_power in 10..14"); return JAERO_E_ARG; }
    int ndev = 0;
    JB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("jaero_batch_create: no such CUDA device"); return JAERO_E_CUDA; }
    JB_CUDA(cudaSetDevice(device));
    jaero_batch *b = new (std::nothrow) jaero_batch();
    if (!b) { set_error("out of host memory"); return JAERO_E_ARG; }
    CreateGuard<jaero_batch> guard(b, jaero_batch_destroy);
    b->set = *s; b->device = device; b->samples = 0; b->bb_pos = 0; b->coarse_counter = 0;
    b->d_stage = 0; b->stage_cap = 0; b->launches = 0;
    b->pre_on = false; b->fir_fill = 0; b->fir_blocks = 0; b->d_x = 0; b->x_cap = 0;
    JB_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    b->own_stream = b->stream; b->profiling = false; b->prof_samples = 0;
    { const char *e = getenv("JAERO_OQPSK_PIPE"); b->use_pipe = !(e && e[0] == '0'); }
    b->copy_stream = 0; b->ev_stage_free = 0; for (int k = 0; k < 8; k++) b->ev_slice[k] = 0; b->n_slices = 0; b->slice_len = 0; b->next_slice = 0;
    b->async_cfe = false; b->cfe_stream = 0; b->ev_seg_done = 0; b->ev_cfe_done[0] = b->ev_cfe_done[1] = 0; b->cfe_count = 0; b->bb_phys = 0;
    DemodParams &p = b->p;
    memset(&p, 0, sizeof p);
    p.kind = s->kind; p.n_channels = n_channels; p.cpad = (n_channels + 31) & ~31;
    p.Fs = s->Fs; p.fb = s->fb; p.lockingbw = s->lockingbw; p.signalthreshold = s->signalthreshold;
    p.afc = s->afc; p.sql = s->sql; p.cpu_reduce = s->cpu_reduce; p.report_ebno = s->report_ebno;
    p.bbnfft = 1 << s->coarsefreqest_fft_power;
    std::vector<double> taps;
    double st_freq;
    if (s->kind == JAERO_KIND_OQPSK) {
        taps = (s->fb == 8400) ? rrc_taps(0.6, 55, s->Fs, s->fb / 2) : rrc_taps(1.0, 55, s->Fs, s->fb / 2);   // oqpskdemodulator.cpp:209-211
        p.agc_len = (int)round(4 * s->Fs);                                    // :197 AGC(4,Fs)
        p.ebno_len = 2 * 48000;                                               // :42 (built in the ctor with Fs=48000)
        p.marg_len = 800; p.dt_len = 401; p.mse_len = 400;                    // :44-45,53
        const double T = s->Fs / (s->fb / 2);                                 // :221
        if (!delay_weights(T / 4.0, &p.k41, p.w41v) || !delay_weights(T / 8.0, &p.k8, p.w8v) || p.k41 > 3 || p.k8 > 3) {
            set_error("unsupported fractional delay for this Fs/fb"); return JAERO_E_ARG; }
        if (s->fb == 8400) {                                                  // :243-250 (the 10 Hz set, assigned last, wins)
            p.res_b0 = 0.0012845857864470789; p.res_b1 = 0; p.res_b2 = -0.0012845857864470789;
            p.res_a1 = -0.90681461999279889; p.res_a2 = 0.99743082842710584;
            p.ee = 0.65;
        } else {
            p.res_b0 = 0.00032714218939589035; p.res_b1 = 0; p.res_b2 = 0.00032714218939589035;   // :256-261
            p.res_a1 = -0.39005299948210803; p.res_a2 = 0.99934571562120822;
            p.ee = 0.4;                                                       // :263
        }
        p.lf_b0 = 0.0010275610653672064; p.lf_b1 = 0.0020551221307344128; p.lf_b2 = 0.0010275610653672064;   // :95-100
        p.lf_a1 = -1.9207386815577139; p.lf_a2 = 0.92509247310306331;
        st_freq = s->fb;                                                      // :270
    } else {
        p.sps = (int)(s->Fs / s->fb);                                         // mskdemodulator.cpp:149
        if (2 * p.sps > MAX_TAPS) { set_error("MSK: 2*SamplesPerSymbol exceeds the supported FIR length"); return JAERO_E_ARG; }
        taps.resize(2 * p.sps);
        for (int i = 0; i < 2 * p.sps; i++) taps[i] = sin(M_PI * i / (2.0 * p.sps)) / (2.0 * p.sps);   // :164-170
        p.agc_len = (int)round(1 * s->Fs);                                    // :173
        p.ebno_len = (int)(2.0 * s->Fs);                                      // :176
        p.marg_len = p.sps; p.dt_len = p.sps / 2 + 1; p.mse_len = 600;        // :254-256, ctor :64
        if (s->fb >= 1200) {                                                  // :189-250
            p.correctionfactor = 0.6;
            if (s->Fs == 48000) { p.res_a1 = -1.993312819378528; p.res_a2 = 0.999476538254407; p.res_b0 = 2.617308727964618e-04; p.res_b2 = -2.617308727964618e-04; p.ee = 0.025; }
            else { p.res_a1 = -1.974342917561558; p.res_a2 = 0.998953350377616; p.res_b0 = 5.233248111921052e-04; p.res_b2 = -5.233248111921052e-04; p.ee = 0.05; }
        } else {
            p.correctionfactor = 1.0;
            if (s->Fs == 48000) { p.res_a1 = -1.998196509168551; p.res_a2 = 0.999738234875681; p.res_b0 = 1.308825621597620e-04; p.res_b2 = -1.308825621597620e-04; p.ee = 0.025; }
            else { p.res_a1 = -1.974342917561558; p.res_a2 = 0.998953350377616; p.res_b0 = 5.233248111921052e-04; p.res_b2 = -5.233248111921052e-04; p.ee = 0.0125; }
        }
        p.res_b1 = 0;
        st_freq = s->fb / 2;                                                  // :159
    }
    if ((p.agc_len % 32) || (p.ebno_len % 32) || p.agc_len < 96 || p.ebno_len < 96) {
        set_error("unsupported sample rate: the AGC / EbNo window lengths must be multiples of 32 samples"); return JAERO_E_ARG; }
    p.ntaps = (int)taps.size();
    p.soft_cap = std::max(4096, (int)(2 * s->fb) + 64);
    if (p.ntaps > MAX_TAPS) { set_error("too many FIR taps"); return JAERO_E_ARG; }
    for (int k = 0; k < p.ntaps; k++) p.taps[k] = taps[k];   // per-batch: the taps ride in the kernel parameter block

    const size_t cp = p.cpad;
    int rc = 0;
    rc |= batch_alloc(b, &p.D, (size_t)D_COUNT * cp);
    rc |= batch_alloc(b, &p.I, (size_t)I_COUNT * cp);
    rc |= batch_alloc(b, &p.agc_ring, (size_t)p.agc_len * cp);
    if (p.report_ebno) { rc |= batch_alloc(b, &p.ebno_e1, (size_t)p.ebno_len * cp); rc |= batch_alloc(b, &p.ebno_e2, (size_t)p.ebno_len * cp); }
    rc |= batch_alloc(b, &p.fir_re, (size_t)(p.ntaps + 1) * cp);
    rc |= batch_alloc(b, &p.fir_im, (size_t)(p.ntaps + 1) * cp);
    {
        // The coarse estimate of a trigger is only consumed by channels that are unlocked, have no carrier detect or are about
        // to re-centre (FreqOffsetEstimateSlot, oqpskdemodulator.cpp:629-677), so in steady state the estimator kernels can run
        // on a second stream while the next segment is demodulated. Conditions: the warp-specialised 10500 bps kernel, the
        // non-cpuReduce schedule, and a segment grid that is fully resident (a waiting CTA must never keep the estimator's
        // CTAs from being scheduled). OFF unless JAERO_ASYNC_CFE=1: measured on B200 (4096 channels) the estimator's FP64 work,
        // when it shares SMs with the latency-bound segment warps, slows both kernels by 3-4x (FP64 pipe contention).
        cudaDeviceProp prop;
        JB_CUDA(cudaGetDeviceProperties(&prop, device));
        const char *e = getenv("JAERO_ASYNC_CFE");
        const int grid = (n_channels + 31) / 32;
        b->async_cfe = (e && e[0] == '1') && s->kind == JAERO_KIND_OQPSK && s->fb != 8400 && !s->cpu_reduce && b->use_pipe &&
                       grid <= prop.multiProcessorCount;
        p.bb_len = b->async_cfe ? p.bbnfft + p.bbnfft / 4 : p.bbnfft;
        if (b->async_cfe) {
            JB_CUDA(cudaStreamCreateWithFlags(&b->cfe_stream, cudaStreamNonBlocking));
            JB_CUDA(cudaEventCreateWithFlags(&b->ev_seg_done, cudaEventDisableTiming));
            JB_CUDA(cudaEventCreateWithFlags(&b->ev_cfe_done[0], cudaEventDisableTiming));
            JB_CUDA(cudaEventCreateWithFlags(&b->ev_cfe_done[1], cudaEventDisableTiming));
        }
        rc |= batch_alloc(b, &p.cfe_flag, (size_t)1);
    }
    rc |= batch_alloc(b, &p.bb, (size_t)n_channels * p.bb_len);
    rc |= batch_alloc(b, &p.marg_ring, (size_t)p.marg_len * cp);
    rc |= batch_alloc(b, &p.mse_pm, (size_t)p.mse_len * cp);
    rc |= batch_alloc(b, &p.mse_ma, (size_t)p.mse_len * cp);
    rc |= batch_alloc(b, &p.dt_ring, (size_t)p.dt_len * cp);
    if (s->kind == JAERO_KIND_MSK) {
        rc |= batch_alloc(b, &p.dsmpl_ring, (size_t)(p.sps + 1) * cp);
        rc |= batch_alloc(b, &p.dly8_ring, (size_t)(p.sps / 2 + 1) * cp);
    }
    rc |= batch_alloc(b, &p.soft, (size_t)n_channels * p.soft_cap);
    rc |= batch_alloc(b, &p.soft_total, (size_t)cp);
    rc |= batch_alloc(b, &p.lost_pos, (size_t)LOST_CAP * cp);
    rc |= batch_alloc(b, &p.cfe_est_out, (size_t)cp);
    if (rc) return JAERO_E_CUDA;

    // trig tables exactly as TrigLookUp builds them (DSP.cpp:19-20), computed with the host libm
    {
        std::vector<double> sn(jb::WTSIZE), cs(jb::WTSIZE);
        for (int i = 0; i < jb::WTSIZE; i++) sn[i] = (sin(2 * M_PI * ((double)i) / jb::WTSIZE));
        for (int i = 0; i < jb::WTSIZE; i++) cs[i] = (sin(M_PI_2 + 2 * M_PI * ((double)i) / jb::WTSIZE));
        double *ds, *dc;
        if (batch_alloc(b, &ds, (size_t)jb::WTSIZE) || batch_alloc(b, &dc, (size_t)jb::WTSIZE)) return JAERO_E_CUDA;
        JB_CUDA(cudaMemcpyAsync(ds, sn.data(), sn.size() * sizeof(double), cudaMemcpyHostToDevice, b->stream));
        JB_CUDA(cudaMemcpyAsync(dc, cs.data(), cs.size() * sizeof(double), cudaMemcpyHostToDevice, b->stream));
        JB_CUDA(cudaStreamSynchronize(b->stream));
        p.sin_t = ds; p.cos_t = dc;
    }
    // coarse estimator plan (CoarseFreqEstimate::setSettings, coarsefreqestimate.cpp:39-76)
    {
        CfePlan &c = b->cfe;
        memset(&c, 0, sizeof c);
        c.nfft = p.bbnfft;
        const int lg = s->coarsefreqest_fft_power;
        c.n1 = 1 << ((lg + 1) / 2); c.n2 = 1 << (lg / 2);
        c.hzperbin = s->Fs / ((double)c.nfft);
        const double lbw = (s->kind == JAERO_KIND_OQPSK) ? 2.0 * s->lockingbw / 2.0 : s->lockingbw;   // oqpskdemodulator.cpp:191
        c.startbin = (int)std::max(round(lbw / c.hzperbin), 1.0);
        c.stopbin = c.nfft - c.startbin;
        c.expectedpeakbin = (int)round(s->fb / (2.0 * c.hzperbin));
        c.lo = (int)round((-lbw / c.hzperbin) + ((double)(c.nfft / 2)));
        c.hi = (int)round((lbw / c.hzperbin) + ((double)(c.nfft / 2)));
        c.is8400 = (s->fb == 8400);
        std::vector<double2> tw(c.nfft);
        for (int k = 0; k < c.nfft; k++) { const double a = -2.0 * M_PI * (double)k / (double)c.nfft; tw[k] = make_double2(cos(a), sin(a)); }
        // channels per pass group: the two work buffers of a group (2 x group x nfft x 16 B) (larger groups amortise launch tails; measured best at >= 512 on B200)
        int grp = 1024;
        if (const char *e = getenv("JAERO_CFE_GROUP")) grp = std::max(1, atoi(e));
        c.group = std::min(n_channels, grp);
        if (c.nfft == 16384) {
            const char *e = getenv("JAERO_CFE_CLUSTER");
            if (!(e && e[0] == '0')) c.clusters = cfe_cluster_capacity();
        }
        if (batch_alloc(b, &c.tw, (size_t)c.nfft) || batch_alloc(b, &c.work_a, (size_t)c.group * c.nfft) ||
            batch_alloc(b, &c.work_b, (size_t)c.group * c.nfft) || batch_alloc(b, &c.y, (size_t)n_channels * c.nfft)) return JAERO_E_CUDA;
        JB_CUDA(cudaMemcpyAsync(c.tw, tw.data(), tw.size() * sizeof(double2), cudaMemcpyHostToDevice, b->stream));
        if (c.is8400) {                                                       // raised-cosine window (coarsefreqestimate.cpp:62-74)
            std::vector<double> win(c.nfft, 0.0);
            win[0] = 1;
            for (int i = 1; i <= c.startbin; i++) {
                double val = cos(M_PI_2 * ((double)i) / ((double)c.startbin)); val *= val;
                if ((c.nfft - i) < 0) break;
                if (i >= c.nfft) break;
                win[c.nfft - i] = val; win[i] = val;
            }
            if (batch_alloc(b, &c.window, (size_t)c.nfft)) return JAERO_E_CUDA;
            JB_CUDA(cudaMemcpyAsync(c.window, win.data(), win.size() * sizeof(double), cudaMemcpyHostToDevice, b->stream));
        }
        JB_CUDA(cudaStreamSynchronize(b->stream));
    }
    if (s->kind == JAERO_KIND_OQPSK && s->fb == 8400) {
        // K6: 2049-tap RRC (alpha 0.6) applied by streaming FFT convolution, nfft 4096 (oqpskdemodulator.cpp:280-283)
        b->pre_on = true;
        std::vector<double> kern = rrc_taps(0.6, 2048, s->Fs, s->fb / 2);
        const int NF = 4096;
        std::vector<std::complex<double>> H(NF, 0.0), tw(NF);
        for (size_t i = 0; i < kern.size(); i++) H[i] = kern[i];
        for (int k = 0; k < NF; k++) { const double a = -2.0 * M_PI * (double)k / (double)NF; tw[k] = std::complex<double>(cos(a), sin(a)); }
        {   // host radix-2 FFT of the kernel
            int bits = 12;
            for (int i = 0; i < NF; i++) { int r = 0; for (int q = 0; q < bits; q++) if (i & (1 << q)) r |= 1 << (bits - 1 - q); if (r > i) std::swap(H[i], H[r]); }
            for (int len = 2; len <= NF; len <<= 1)
                for (int i = 0; i < NF; i += len)
                    for (int k = 0; k < len / 2; k++) { auto w = tw[k * (NF / len)]; auto u = H[i + k], v = H[i + k + len / 2] * w; H[i + k] = u + v; H[i + k + len / 2] = u - v; }
        }
        FirStream &f = b->fir; memset(&f, 0, sizeof f);
        PreParams &q = b->pre; memset(&q, 0, sizeof q);
        int rc2 = 0;
        rc2 |= batch_alloc(b, &f.H, (size_t)NF); rc2 |= batch_alloc(b, &f.tw, (size_t)NF);
        rc2 |= batch_alloc(b, &f.hist, (size_t)n_channels * FIR_L); rc2 |= batch_alloc(b, &f.inblk, (size_t)n_channels * FIR_L);
        rc2 |= batch_alloc(b, &f.outblk, (size_t)n_channels * FIR_L);
        rc2 |= batch_alloc(b, &q.osc, (size_t)4 * cp);
        rc2 |= batch_alloc(b, &p.m2_freq_sum, (size_t)cp);
        if (rc2) return JAERO_E_CUDA;
        JB_CUDA(cudaMemcpyAsync(f.H, H.data(), NF * sizeof(double2), cudaMemcpyHostToDevice, b->stream));
        JB_CUDA(cudaMemcpyAsync(f.tw, tw.data(), NF * sizeof(double2), cudaMemcpyHostToDevice, b->stream));
        // mixer_fir_pre.SetFreq(freq_center,Fs) is only done in the ctor, with the ctor's 8000 Hz (oqpskdemodulator.cpp:21,115)
        std::vector<double> osc(4 * cp, 0.0);
        for (size_t c2 = 0; c2 < cp; c2++) { osc[1 * cp + c2] = (8000.0) * ((double)jb::WTSIZE) / ((double)((float)48000)); osc[2 * cp + c2] = 8000.0; }
        JB_CUDA(cudaMemcpyAsync(q.osc, osc.data(), osc.size() * sizeof(double), cudaMemcpyHostToDevice, b->stream));
        JB_CUDA(cudaStreamSynchronize(b->stream));
        q.n_channels = n_channels; q.cpad = p.cpad; q.sin_t = p.sin_t; q.cos_t = p.cos_t;
    }
    // per-channel initial state
    {
        std::vector<double> fc(n_channels);
        for (int i = 0; i < n_channels; i++) fc[i] = freq_center_per_channel ? freq_center_per_channel[i] : s->freq_center;
        double *dfc;
        JB_CUDA(cudaMalloc(&dfc, n_channels * sizeof(double)));
        JB_CUDA(cudaMemcpyAsync(dfc, fc.data(), n_channels * sizeof(double), cudaMemcpyHostToDevice, b->stream));
        init_state_kernel<<<(n_channels + 127) / 128, 128, 0, b->stream>>>(p, dfc, st_freq, 0.0);
        JB_CUDA(cudaGetLastError());
        JB_CUDA(cudaStreamSynchronize(b->stream));
        cudaFree(dfc);
    }
    b->d_chan_of = 0; b->d_keys = 0; b->h_keys = 0; b->d_ring_scratch = 0; b->ring_scratch_count = 0; b->epochs = 0; b->regroups = 0;
    b->regroup_every = 0; b->next_regroup = 0;
    if (s->kind == JAERO_KIND_OQPSK && s->fb > 8400 && b->use_pipe && !s->cpu_reduce) {
        // the pipelined kernel seats channels by symbol-timing phase: first after 2.9 s of signal (the timing loops' phases are final
        // to a few hundredths of a sample by then; at 2 s they are not), a check 2.7 s later (a no-op unless they moved),
        // then every JAERO_REGROUP_EPOCHS estimator epochs (default 128 = 11 s; 0 = never: symbol clocks of different transmitters
        // drift by a sample in minutes, not seconds)
        const char *e = getenv("JAERO_REGROUP_EPOCHS");
        b->regroup_every = e ? atoi(e) : 128;
        if (batch_alloc(b, &b->d_chan_of, (size_t)2 * cp) || batch_alloc(b, &b->d_keys, (size_t)cp)) return JAERO_E_CUDA;
        JB_CUDA(cudaMallocHost(&b->h_keys, cp * sizeof(double)));
        b->slot_of.resize(cp);
        std::vector<int> ident(cp);
        for (size_t c = 0; c < cp; c++) { ident[c] = (int)c; b->slot_of[c] = (int)c; }
        JB_CUDA(cudaMemcpy(b->d_chan_of, ident.data(), cp * sizeof(int), cudaMemcpyHostToDevice));
        p.chan_of = b->d_chan_of;
        b->next_regroup = b->regroup_every > 0 ? 34 : -1;
    }
    JB_CUDA(cudaMallocHost(&b->h_ints, (size_t)I_COUNT * cp * sizeof(int)));
    JB_CUDA(cudaMallocHost(&b->h_dbls, (size_t)D_COUNT * cp * sizeof(double)));
    JB_CUDA(cudaMallocHost(&b->h_soft_total, cp * sizeof(long long)));
    JB_CUDA(cudaMallocHost(&b->h_soft_stage, (size_t)n_channels * p.soft_cap * sizeof(int16_t)));
    guard.release();
    *out = b;
    return JAERO_OK;
}

void jaero_batch_destroy(jaero_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    cudaStreamSynchronize(b->stream);
    if (b->cfe_stream) { cudaStreamSynchronize(b->cfe_stream); cudaStreamDestroy(b->cfe_stream); }
    if (b->copy_stream) { cudaStreamSynchronize(b->copy_stream); cudaStreamDestroy(b->copy_stream); }
    if (b->ev_stage_free) cudaEventDestroy(b->ev_stage_free);
    for (int k = 0; k < 8; k++) if (b->ev_slice[k]) cudaEventDestroy(b->ev_slice[k]);
    if (b->ev_seg_done) cudaEventDestroy(b->ev_seg_done);
    for (int k = 0; k < 2; k++) if (b->ev_cfe_done[k]) cudaEventDestroy(b->ev_cfe_done[k]);
    for (void *q : b->allocs) cudaFree(q);
    cudaFree(b->d_stage); cudaFree(b->d_x);
    cudaFreeHost(b->h_ints); cudaFreeHost(b->h_dbls); cudaFreeHost(b->h_soft_stage); cudaFreeHost(b->h_soft_total); cudaFreeHost(b->h_keys);
    cudaFree(b->d_ring_scratch);
    for (auto &e : b->ev_seg) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto &e : b->ev_cfe) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (b->own_stream) cudaStreamDestroy(b->own_stream);
    delete b;
}
int jaero_batch_channels(const jaero_batch *b) { return b ? b->p.n_channels : 0; }
int64_t jaero_batch_launch_count(const jaero_batch *b) { return b ? b->launches : 0; }

int jaero_batch_set_stream(jaero_batch *b, void *cuda_stream)
{
    if (!b) { set_error("null handle"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    JB_CUDA(cudaStreamSynchronize(b->stream));
    b->stream = cuda_stream ? (cudaStream_t)cuda_stream : b->own_stream;
    return JAERO_OK;
}
int jaero_batch_set_profiling(jaero_batch *b, int enabled)
{
    if (!b) { set_error("null handle"); return JAERO_E_ARG; }
    b->profiling = enabled != 0;
    return JAERO_OK;
}
int jaero_batch_get_profile(jaero_batch *b, double out[5])
{
    if (!b || !out) { set_error("null argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    JB_CUDA(cudaStreamSynchronize(b->stream));
    double seg = 0, cfe = 0;
    for (auto &e : b->ev_seg) { float ms = 0; cudaEventElapsedTime(&ms, e.first, e.second); seg += ms; cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    for (auto &e : b->ev_cfe) { float ms = 0; cudaEventElapsedTime(&ms, e.first, e.second); cfe += ms; cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    out[0] = seg; out[1] = (double)b->ev_seg.size(); out[2] = cfe; out[3] = (double)b->ev_cfe.size(); out[4] = b->prof_samples;
    b->ev_seg.clear(); b->ev_cfe.clear(); b->prof_samples = 0;
    return JAERO_OK;
}
int jaero_batch_sync(jaero_batch *b)
{
    if (!b) { set_error("null handle"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    JB_CUDA(cudaStreamSynchronize(b->stream));
    return JAERO_OK;
}

int jaero_batch_write_device(jaero_batch *b, const int16_t *d_pcm, size_t n, size_t stride)
{
    if (!b || !d_pcm) { set_error("jaero_batch_write_device: null argument"); return JAERO_E_ARG; }
    if (n == 0) return JAERO_OK;                                   // `if(!len)return 0;` oqpskdemodulator.cpp:337
    if (stride < n || n > 0x7fffffff) { set_error("jaero_batch_write_device: bad stride / length"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    const DemodParams &p = b->p;
    if ((((uintptr_t)d_pcm) & 15) || (stride & 7)) {
        // the kernels stage PCM rows with 16-byte bulk copies: re-pitch unaligned caller buffers on the device
        const size_t C = p.n_channels, pitch = (n + 7) & ~(size_t)7;
        if (d_pcm == b->d_stage) { set_error("internal: staging buffer misaligned"); return JAERO_E_STATE; }
        if (C * pitch > b->stage_cap) {
            JB_CUDA(cudaStreamSynchronize(b->stream));
            cudaFree(b->d_stage); b->d_stage = 0;
            JB_CUDA(cudaMalloc(&b->d_stage, C * pitch * sizeof(int16_t)));
            b->stage_cap = C * pitch;
        }
        JB_CUDA(cudaMemcpy2DAsync(b->d_stage, pitch * sizeof(int16_t), d_pcm, stride * sizeof(int16_t), n * sizeof(int16_t), C,
                                  cudaMemcpyDeviceToDevice, b->stream));
        d_pcm = b->d_stage; stride = pitch;
    }
    if (b->pre_on) {
        while (b->next_slice < b->n_slices) { JB_CUDA(cudaStreamWaitEvent(b->stream, b->ev_slice[b->next_slice], 0)); b->next_slice++; }
        // K6 over the whole call first (oqpskdemodulator.cpp:343-381), then the per-sample loop consumes its output
        const size_t C = p.n_channels, xs = (n + 7) & ~(size_t)7;
        if (C * xs > b->x_cap) {
            JB_CUDA(cudaStreamSynchronize(b->stream));
            cudaFree(b->d_x); b->d_x = 0;
            JB_CUDA(cudaMalloc(&b->d_x, C * xs * sizeof(double2)));
            b->x_cap = C * xs;
        }
        b->pre.x = b->d_x; b->pre.xstride = xs;
        b->p.xpre = b->d_x; b->p.xstride = xs;
        if (pre_down_launch(b->pre, d_pcm, stride, (int)n, b->stream)) return JAERO_E_CUDA;
        b->launches++;
        int i = 0;
        while (i < (int)n) {
            const int room = FIR_L - b->fir_fill;
            const int take = std::min(room, (int)n - i);
            if (fir_exchange_up_launch(b->pre, b->fir, i, i + take, b->fir_fill, b->stream)) return JAERO_E_CUDA;
            b->launches++;
            b->fir_fill += take; i += take;
            if (b->fir_fill == FIR_L) {
                if (fir_block_launch(b->fir, p.n_channels, b->fir_blocks == 0 ? 1 : 0, b->stream)) return JAERO_E_CUDA;
                b->launches++;
                b->fir_fill = 0; b->fir_blocks++;
            }
        }
    }
    peak_kernel<<<(p.n_channels + 3) / 4, 128, 0, b->stream>>>(p, d_pcm, stride, (int)n);
    JB_CUDA(cudaGetLastError());
    b->launches++;
    const int N = p.bbnfft, trig_every = p.cpu_reduce ? N : N / 4;
    SegmentArgs a;
    memset(&a, 0, sizeof a);
    a.new_write = 1;
    int seg_start = 0; bool resume = false;
    auto launch = [&](int i0, int i1, bool stop_after_a, int bb0, int cc0) -> int {
        a.sample0 = b->samples; a.i0 = i0; a.i1 = i1; a.skip_a_first = resume ? 1 : 0; a.stop_after_a = stop_after_a ? 1 : 0;
        a.apply_cfe = resume ? 1 : 0; a.bb_pos = bb0; a.coarse_counter = cc0;
        a.cfe_wait = (resume && b->async_cfe) ? b->cfe_count : 0;
        while (b->next_slice < b->n_slices && b->next_slice * b->slice_len < i1) {   // input slices this segment reads
            JB_CUDA(cudaStreamWaitEvent(b->stream, b->ev_slice[b->next_slice], 0));
            b->next_slice++;
        }
        long long *d_trace = 0;
        if (const char *tf = getenv("JAERO_PIPE_TRACE")) {   // development aid: stage time stamps of one K1a launch
            if (!b->trace_state && b->launches > 300 && (i1 - i0) > 4000 && p.kind == JAERO_KIND_OQPSK && b->use_pipe && !p.xpre && p.fb > 8400) {
                (void)tf; cudaMalloc(&d_trace, 64 * 16 * sizeof(long long)); cudaMemset(d_trace, 0, 64 * 16 * sizeof(long long));
                a.trace = d_trace; a.trace_j0 = 2000;
            }
        }
        cudaEvent_t e0 = 0, e1 = 0;
        if (b->profiling) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, b->stream); }
        int r = (p.kind == JAERO_KIND_OQPSK) ? ((b->use_pipe && !p.xpre && p.fb > 8400) ? oqpsk_pipe_launch(p, a, d_pcm, stride, b->stream)
                                                                          : oqpsk_segment_launch(p, a, d_pcm, stride, b->stream))
                                             : ((b->use_pipe && (p.agc_len % 32) == 0 && (p.ebno_len % 32) == 0) ? msk_pipe_launch(p, a, d_pcm, stride, b->stream)
                                                                                                                      : msk_segment_launch(p, a, d_pcm, stride, b->stream));
        if (b->profiling) { cudaEventRecord(e1, b->stream); b->ev_seg.push_back({e0, e1}); b->prof_samples += (i1 - i0 - (stop_after_a ? 1 : 0)); }
        if (d_trace) {
            std::vector<long long> h(64 * 16);
            cudaStreamSynchronize(b->stream);
            cudaMemcpy(h.data(), d_trace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
            if (FILE *f = fopen(getenv("JAERO_PIPE_TRACE"), "w")) {
                for (int jj = 0; jj < 64; jj++) { for (int k = 0; k < 16; k++) fprintf(f, "%lld ", h[jj * 16 + k]); fprintf(f, "\n"); }
                fclose(f);
            }
            cudaFree(d_trace); a.trace = 0; b->trace_state = 1;
        }
        b->launches++;
        a.new_write = 0;
        return r;
    };
    int bb = b->bb_pos, cc = b->coarse_counter, bbp = b->bb_phys;  // bb: bbcycbuff_ptr of the reference; bbp: slot in our ring
    int seg_bb = bbp, seg_cc = cc;                                 // counters at the start of the open segment
    bool cfe_in_flight = false;
    for (int i = 0; i < (int)n; i++) {
        // A(i): ring write + trigger test (oqpskdemodulator.cpp:410-429) — lock-step for the whole batch
        bool trigger = false;
        if (cc >= p.Fs || !p.cpu_reduce) {
            bb++; if (bb >= N) bb = 0;
            bbp++; if (bbp >= p.bb_len) bbp = 0;
            if (bb % trig_every == 0) trigger = true;
        }
        if (trigger) {
            if (launch(seg_start, i + 1, true, seg_bb, seg_cc)) return JAERO_E_CUDA;
            b->samples += (i - seg_start);                         // samples whose B part has run
            cudaEvent_t c0 = 0, c1 = 0;
            int oldest = bbp + (p.bb_len - N); if (oldest >= p.bb_len) oldest -= p.bb_len;
            cudaStream_t cs = b->async_cfe ? b->cfe_stream : b->stream;
            if (b->async_cfe) {
                JB_CUDA(cudaEventRecord(b->ev_seg_done, b->stream));
                JB_CUDA(cudaStreamWaitEvent(cs, b->ev_seg_done, 0));
            }
            if (b->profiling) { cudaEventCreate(&c0); cudaEventCreate(&c1); cudaEventRecord(c0, cs); }
            if (b->cfe.clusters > 0 ? cfe_cluster_run(b->cfe, p, oldest, std::min(b->cfe.clusters, p.n_channels), cs, &b->launches)
                                    : cfe_run(b->cfe, p, oldest, cs, &b->launches)) return JAERO_E_CUDA;
            if (b->profiling) { cudaEventRecord(c1, cs); b->ev_cfe.push_back({c0, c1}); }
            if (b->async_cfe) {
                b->cfe_count++;
                if (cfe_mark_launch(p.cfe_flag, b->cfe_count, cs)) return JAERO_E_CUDA;
                b->launches++;
                JB_CUDA(cudaEventRecord(b->ev_cfe_done[b->cfe_count & 1], cs));
                // the segment after the next one overwrites the quarter this estimate reads first: order the NEXT segment
                // behind the PREVIOUS estimate (a no-op in steady state)
                if (cfe_in_flight) JB_CUDA(cudaStreamWaitEvent(b->stream, b->ev_cfe_done[(b->cfe_count - 1) & 1], 0));
                cfe_in_flight = true;
            }
            b->epochs++;
            // seating check between two launches (the segment kernel has written its ring tiles back; per-channel state is indexed
            // by channel, only the ring rows and chan_of move)
            if (b->regroup_every > 0 && b->next_regroup >= 0 && b->epochs >= b->next_regroup) {
                if (batch_regroup_by_phase(b, false)) return JAERO_E_CUDA;
                b->next_regroup = b->epochs + (b->epochs < 64 ? std::min(32, b->regroup_every) : b->regroup_every);
            }
            cc = 0;                                                // :426
            seg_start = i; resume = true; seg_bb = bbp; seg_cc = 0;
        }
        cc++;                                                      // :431
    }
    if (launch(seg_start, (int)n, false, seg_bb, seg_cc)) return JAERO_E_CUDA;
    b->samples += ((int)n - seg_start);
    while (b->next_slice < b->n_slices) { JB_CUDA(cudaStreamWaitEvent(b->stream, b->ev_slice[b->next_slice], 0)); b->next_slice++; }
    b->n_slices = 0; b->next_slice = 0;
    b->bb_pos = bb; b->coarse_counter = cc; b->bb_phys = bbp;
    if (cfe_in_flight) JB_CUDA(cudaStreamWaitEvent(b->stream, b->ev_cfe_done[b->cfe_count & 1], 0));   // join: a call leaves nothing in flight
    if (b->pre_on) {                                               // :608 mixer_fir_pre.SetFreq(mixer2_freq_sum/i)
        if (pre_finish_launch(b->pre, p.m2_freq_sum, (int)n, p.Fs, b->stream)) return JAERO_E_CUDA;
        b->launches++;
    }
    return JAERO_OK;
}

int jaero_batch_write(jaero_batch *b, const int16_t *pcm, size_t n, size_t stride)
{
    if (!b || !pcm) { set_error("jaero_batch_write: null argument"); return JAERO_E_ARG; }
    if (n == 0) return JAERO_OK;
    if (stride < n) { set_error("jaero_batch_write: channel_stride < n_samples"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    const size_t C = b->p.n_channels;
    const size_t pitch = (n + 7) & ~(size_t)7;
    if (C * pitch > b->stage_cap) {
        JB_CUDA(cudaStreamSynchronize(b->stream));
        cudaFree(b->d_stage); b->d_stage = 0;
        JB_CUDA(cudaMalloc(&b->d_stage, C * pitch * sizeof(int16_t)));
        b->stage_cap = C * pitch;
    }
    if (n < 8192) {
        JB_CUDA(cudaMemcpy2DAsync(b->d_stage, pitch * sizeof(int16_t), pcm, stride * sizeof(int16_t), n * sizeof(int16_t), C,
                                  cudaMemcpyHostToDevice, b->stream));
        return jaero_batch_write_device(b, b->d_stage, n, pitch);
    }
    // long calls: copy in column slices on a second stream so that the transfer of later samples overlaps the
    // demodulation of earlier ones (pinned host memory makes the copies truly asynchronous)
    if (!b->copy_stream) {
        JB_CUDA(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
        JB_CUDA(cudaEventCreateWithFlags(&b->ev_stage_free, cudaEventDisableTiming));
        for (int k = 0; k < 8; k++) JB_CUDA(cudaEventCreateWithFlags(&b->ev_slice[k], cudaEventDisableTiming));
    }
    JB_CUDA(cudaEventRecord(b->ev_stage_free, b->stream));          // everything already queued that reads the staging buffer
    JB_CUDA(cudaStreamWaitEvent(b->copy_stream, b->ev_stage_free, 0));
    b->slice_len = (int)((((n + 7) / 8) + 7) & ~(size_t)7);
    b->n_slices = 0; b->next_slice = 0;
    for (size_t s0 = 0; s0 < n; s0 += (size_t)b->slice_len) {
        const size_t len = std::min((size_t)b->slice_len, n - s0);
        JB_CUDA(cudaMemcpy2DAsync(b->d_stage + s0, pitch * sizeof(int16_t), pcm + s0, stride * sizeof(int16_t), len * sizeof(int16_t), C,
                                  cudaMemcpyHostToDevice, b->copy_stream));
        JB_CUDA(cudaEventRecord(b->ev_slice[b->n_slices], b->copy_stream));
        b->n_slices++;
    }
    return jaero_batch_write_device(b, b->d_stage, n, pitch);
}

static int pull_ints(jaero_batch *b)
{
    const size_t cp = b->p.cpad;
    JB_CUDA(cudaMemcpyAsync(b->h_ints, b->p.I, (size_t)I_COUNT * cp * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    JB_CUDA(cudaStreamSynchronize(b->stream));
    return 0;
}

int jaero_batch_read_softbits(jaero_batch *b, int16_t *out, size_t cap, int32_t *counts)
{
    if (!b || !out || !counts) { set_error("jaero_batch_read_softbits: null argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    if (pull_ints(b)) return JAERO_E_CUDA;
    const DemodParams &p = b->p;
    const size_t cp = p.cpad;
    const int *cnt = b->h_ints + (size_t)I_SOFT_COUNT * cp, *ovf = b->h_ints + (size_t)I_SOFT_OVERFLOW * cp;
    int maxc = 0; bool overflow = false;
    for (int ch = 0; ch < p.n_channels; ch++) { maxc = std::max(maxc, cnt[ch]); overflow |= (ovf[ch] != 0) || ((size_t)cnt[ch] > cap); }
    if (overflow) { set_error("soft-bit ring overflow: drain more often or pass a larger buffer"); return JAERO_E_OVERFLOW; }
    if (maxc > 0) {
        JB_CUDA(cudaMemcpy2DAsync(b->h_soft_stage, (size_t)p.soft_cap * 2, p.soft, (size_t)p.soft_cap * 2, (size_t)maxc * 2, p.n_channels,
                                  cudaMemcpyDeviceToHost, b->stream));
        JB_CUDA(cudaStreamSynchronize(b->stream));
    }
    for (int ch = 0; ch < p.n_channels; ch++) {
        counts[ch] = cnt[ch];
        if (cnt[ch]) memcpy(out + (size_t)ch * cap, b->h_soft_stage + (size_t)ch * p.soft_cap, (size_t)cnt[ch] * 2);
    }
    soft_reset_kernel<<<(p.n_channels + 127) / 128, 128, 0, b->stream>>>(p);
    JB_CUDA(cudaGetLastError());
    return JAERO_OK;
}
int jaero_batch_softbits_device(jaero_batch *b, const int16_t **d_soft, const int32_t **d_counts, size_t *ring_cap)
{
    if (!b || !d_soft || !d_counts || !ring_cap) { set_error("null argument"); return JAERO_E_ARG; }
    *d_soft = b->p.soft; *d_counts = b->p.I + (size_t)I_SOFT_COUNT * b->p.cpad; *ring_cap = (size_t)b->p.soft_cap;
    return JAERO_OK;
}
int jaero_batch_reset_softbits(jaero_batch *b)
{
    if (!b) { set_error("null handle"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    soft_reset_kernel<<<(b->p.n_channels + 127) / 128, 128, 0, b->stream>>>(b->p);
    JB_CUDA(cudaGetLastError());
    return JAERO_OK;
}
int jaero_batch_set_dcd(jaero_batch *b, int channel, int dcd)
{
    if (!b || channel >= b->p.n_channels) { set_error("jaero_batch_set_dcd: bad argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    set_int_kernel<<<(b->p.n_channels + 127) / 128, 128, 0, b->stream>>>(b->p, I_DCD, channel, dcd ? 1 : 0);
    JB_CUDA(cudaGetLastError());
    return JAERO_OK;
}
int jaero_batch_set_center_freq(jaero_batch *b, int channel, double hz)
{
    if (!b || channel >= b->p.n_channels) { set_error("jaero_batch_set_center_freq: bad argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    center_freq_kernel<<<(b->p.n_channels + 127) / 128, 128, 0, b->stream>>>(b->p, channel, hz);
    JB_CUDA(cudaGetLastError());
    return JAERO_OK;
}
// setAFC / setSQL / setCPUReduce (oqpskdemodulator.cpp:149-167, mskdemodulator.cpp:113-131): plain flags the sample loop reads;
// the kernels take them by value at every launch, so a change applies from the next write on
int jaero_batch_set_afc(jaero_batch *b, int state)
{
    if (!b) { set_error("null handle"); return JAERO_E_ARG; }
    b->p.afc = state ? 1 : 0;
    return JAERO_OK;
}
int jaero_batch_set_sql(jaero_batch *b, int state)
{
    if (!b) { set_error("null handle"); return JAERO_E_ARG; }
    b->p.sql = state ? 1 : 0;
    return JAERO_OK;
}
// connect(demodulator, SignalStatus(bool), aerol, SignalStatusSlot(bool)) (JAERO/mainwindow.cpp:432,508)
int jaero_batch_wire_signal_status(jaero_batch *b, int enabled)
{
    if (!b) { set_error("null handle"); return JAERO_E_ARG; }
    b->p.wire_sigstat = enabled ? 1 : 0;
    return JAERO_OK;
}
// Seat the channels of the pipelined 10500 bps kernel: slot_of[c] = seat of channel c (a permutation of 0..n_channels-1), or NULL
// to seat them by symbol-timing phase now. Results never depend on the seating (channels do not interact); throughput does.
int jaero_batch_regroup(jaero_batch *b, const int32_t *slot_of)
{
    if (!b) { set_error("null handle"); return JAERO_E_ARG; }
    if (!b->d_chan_of) return JAERO_OK;                        // this batch's kernel has a fixed seating
    JB_CUDA(cudaSetDevice(b->device));
    if (!slot_of) return batch_regroup_by_phase(b, true) ? JAERO_E_CUDA : JAERO_OK;
    const int C = b->p.n_channels, cp = b->p.cpad;
    std::vector<int> v(cp), seen(C, 0);
    for (int c = 0; c < C; c++) { if (slot_of[c] < 0 || slot_of[c] >= C || seen[slot_of[c]]) { set_error("jaero_batch_regroup: not a permutation"); return JAERO_E_ARG; } seen[slot_of[c]] = 1; v[c] = slot_of[c]; }
    for (int c = C; c < cp; c++) v[c] = c;
    return batch_apply_seating(b, v) ? JAERO_E_CUDA : JAERO_OK;
}
int jaero_batch_set_cpu_reduce(jaero_batch *b, int state)
{
    if (!b) { set_error("null handle"); return JAERO_E_ARG; }
    if (b->async_cfe) { set_error("jaero_batch_set_cpu_reduce: not available with JAERO_ASYNC_CFE=1"); return JAERO_E_STATE; }
    b->p.cpu_reduce = state ? 1 : 0;
    return JAERO_OK;
}
int jaero_batch_get_status_all(jaero_batch *b, jaero_status *out)
{
    if (!b || !out) { set_error("null argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    const size_t cp = b->p.cpad;
    JB_CUDA(cudaMemcpyAsync(b->h_dbls, b->p.D, (size_t)D_COUNT * cp * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    JB_CUDA(cudaMemcpyAsync(b->h_soft_total, b->p.soft_total, cp * sizeof(long long), cudaMemcpyDeviceToHost, b->stream));
    if (pull_ints(b)) return JAERO_E_CUDA;
    for (int ch = 0; ch < b->p.n_channels; ch++) {
        auto D = [&](int i) { return b->h_dbls[(size_t)i * cp + ch]; };
        auto I = [&](int i) { return b->h_ints[(size_t)i * cp + ch]; };
        jaero_status &s = out[ch];
        s.mixer2_freq = D(D_M2_FREQ); s.mixer2_wtptr = D(D_M2_PTR); s.center_freq = D(D_MC_FREQ);
        s.st_freq = D(D_ST_FREQ); s.st_wtptr = D(D_ST_PTR); s.agc = D(D_AGC_VAL); s.mse = D(D_MSE);
        s.ebno = D(D_EB_EBNO); s.marg = D(D_MARG_VAL); s.cfe_est = D(D_CFE_EST);
        s.n_sig_true = I(I_SIG_TRUE); s.n_sig_false = I(I_SIG_FALSE);
        s.center_wtptr = D(D_MC_PTR); s.st_ref_wtptr = D(D_SR_PTR);
        s.samples = b->samples; s.softbits = b->h_soft_total[ch] + I(I_SOFT_COUNT); s.dcd = I(I_DCD); s.reserved = 0;
        s.peak_volume = (double)I(I_PEAK) / 32768.0;
        s.scatter[0] = D(D_SCAT0_RE); s.scatter[1] = D(D_SCAT0_IM); s.scatter[2] = D(D_SCAT1_RE); s.scatter[3] = D(D_SCAT1_IM);
    }
    // `emit PeakVolume(maxval); maxval=0;`: the read-out restarts the maximum
    JB_CUDA(cudaMemsetAsync(b->p.I + (size_t)I_PEAK * cp, 0, cp * sizeof(int), b->stream));
    return JAERO_OK;
}
int jaero_batch_get_status(jaero_batch *b, int channel, jaero_status *out)
{
    if (!b || !out || channel < 0 || channel >= b->p.n_channels) { set_error("jaero_batch_get_status: bad argument"); return JAERO_E_ARG; }
    std::vector<jaero_status> all(b->p.n_channels);
    int r = jaero_batch_get_status_all(b, all.data());
    if (r) return r;
    *out = all[channel];
    return JAERO_OK;
}

} // extern "C"

// ====================================================================== P-channel frame layer
#include "pchannel.cuh"

struct jaero_pchannel {
    int device; cudaStream_t stream;
    PChanParams pp;
    std::vector<void *> allocs;
    uint8_t *vit_overlap; int *vit_overlap_len, *vit_renorm, *vit_valid;
    int16_t *d_soft_stage; int *d_count_stage; size_t stage_cap;
    PChanState *h_state; uint8_t *h_su;
    long long launches;
    // Stream ordering between this layer's own stream and a demodulator batch's stream (which the layer never keeps: the
    // batch may be destroyed first). Work launched on a batch's stream is followed by ev_batch, which `stream` waits for;
    // work on `stream` sets own_dirty, and the next call that uses a batch's stream orders that stream behind ev_own.
    cudaEvent_t ev_batch, ev_own; bool own_dirty;
};

namespace {
template <class T> int pc_alloc(jaero_pchannel *p, T **ptr, size_t count)
{
    int r = dev_alloc_zero(ptr, count, p->stream);
    if (r == 0) p->allocs.push_back((void *)*ptr);
    return r;
}
__global__ void pchan_su_reset_kernel(PChanParams pp)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch < pp.n_channels) { pp.state[ch].su_count = 0; pp.state[ch].queue_overflow = 0; }
}
// a call is about to launch on a batch's stream `bs`: order it behind whatever this layer queued on its own stream
template <class L> int pc_enter(L *p, cudaStream_t bs)
{
    if (p->own_dirty) { JB_CUDA(cudaEventRecord(p->ev_own, p->stream)); JB_CUDA(cudaStreamWaitEvent(bs, p->ev_own, 0)); p->own_dirty = false; }
    return 0;
}
// ... and the layer's own stream behind what was just launched on `bs` (the batch's stream handle is not kept)
template <class L> int pc_leave(L *p, cudaStream_t bs)
{
    JB_CUDA(cudaEventRecord(p->ev_batch, bs)); JB_CUDA(cudaStreamWaitEvent(p->stream, p->ev_batch, 0));
    return 0;
}
} // namespace

extern "C" {

int jaero_pchannel_create(int n_channels, double fb, int device, jaero_pchannel **out)
{
    if (!out || n_channels <= 0) { set_error("jaero_pchannel_create: bad argument"); return JAERO_E_ARG; }
    const int ifb = (int)(fb + 0.5);
    if (ifb != 600 && ifb != 1200 && ifb != 10500) { set_error("jaero_pchannel_create: P-channel rates are 600, 1200, 10500"); return JAERO_E_ARG; }
    int ndev = 0;
    JB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("jaero_pchannel_create: no such CUDA device"); return JAERO_E_CUDA; }
    JB_CUDA(cudaSetDevice(device));
    jaero_pchannel *p = new (std::nothrow) jaero_pchannel();
    if (!p) { set_error("out of host memory"); return JAERO_E_ARG; }
    CreateGuard<jaero_pchannel> guard(p, jaero_pchannel_destroy);
    p->device = device; p->launches = 0; p->d_soft_stage = 0; p->d_count_stage = 0; p->stage_cap = 0;
    JB_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    JB_CUDA(cudaEventCreateWithFlags(&p->ev_batch, cudaEventDisableTiming));
    JB_CUDA(cudaEventCreateWithFlags(&p->ev_own, cudaEventDisableTiming));
    p->own_dirty = false;
    PChanParams &pp = p->pp;
    memset(&pp, 0, sizeof pp);
    pp.n_channels = n_channels; pp.paddinglength = 24;                          // aerol.cpp:940
    switch (ifb) {                                                               // AeroL::setSettings, aerol.cpp:1013-1052
    case 600: pp.cols = 6; pp.number_of_bits = 1152; pp.bits_in_header = 16; pp.total_number_of_bits = 16 + 1152 + 32; pp.oqpsk = 0; pp.dl2_len = 576 - 6 + 1; break;
    case 1200: pp.cols = 9; pp.number_of_bits = 1152; pp.bits_in_header = 16; pp.total_number_of_bits = 16 + 1152 + 32; pp.oqpsk = 0; pp.dl2_len = 576 - 6 + 1; break;
    default: pp.cols = 78; pp.number_of_bits = 4992; pp.bits_in_header = 16 + 178; pp.total_number_of_bits = 16 + 178 + 4992 + 64; pp.oqpsk = 1; pp.dl2_len = 4992 - 6 + 1; break;
    }
    pp.block_len = pp.cols * 64;
    pp.info_cap = pp.number_of_bits / 16 + 16;
    // queue depth: a demodulator soft ring holds max(4096, 2*fb+64) values (jaero_batch_create); a call may hand all of them over
    pp.queue = std::max(PCHAN_QUEUE_MIN, std::max(4096, 2 * ifb + 64) / pp.block_len + 2);
    pp.su_cap = pp.queue * (pp.number_of_bits / 2 / 96) + 8;
    const size_t C = n_channels;
    int rc = 0;
    rc |= pc_alloc(p, &pp.state, C);
    rc |= pc_alloc(p, &pp.blocks, C * pp.queue * pp.block_len);
    rc |= pc_alloc(p, &pp.decoded, C * pp.queue * (pp.block_len / 2));
    rc |= pc_alloc(p, &pp.meta, C * pp.queue);
    rc |= pc_alloc(p, &pp.ready, C);
    rc |= pc_alloc(p, &pp.dl2, C * pp.dl2_len);
    rc |= pc_alloc(p, &pp.infofield, C * pp.info_cap);
    rc |= pc_alloc(p, &pp.su_out, C * pp.su_cap * 16);
    rc |= pc_alloc(p, &p->vit_overlap, C * 64);
    rc |= pc_alloc(p, &p->vit_overlap_len, C);
    rc |= pc_alloc(p, &p->vit_renorm, C);
    rc |= pc_alloc(p, &p->vit_valid, C);
    if (rc) return JAERO_E_CUDA;
    {   // AeroLScrambler::pre_state (aerol.h:397-419)
        int st[15] = {1, 1, 0, 1, 0, 0, 1, 0, 1, 0, 1, 1, 0, 0, 1};
        std::vector<uint8_t> seq(5000);
        for (int a = 0; a < 5000; a++) { int v = st[0] ^ st[14]; seq[a] = (uint8_t)v; for (int i = 14; i > 0; i--) st[i] = st[i - 1]; st[0] = v; }
        if (pchan_set_scrambler(seq.data())) return JAERO_E_CUDA;
    }
    if (pchan_init(pp, p->stream)) return JAERO_E_CUDA;
    JB_CUDA(cudaStreamSynchronize(p->stream));
    JB_CUDA(cudaMallocHost(&p->h_state, C * sizeof(PChanState)));
    JB_CUDA(cudaMallocHost(&p->h_su, C * pp.su_cap * 16));
    guard.release();
    *out = p;
    return JAERO_OK;
}
void jaero_pchannel_destroy(jaero_pchannel *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    for (void *q : p->allocs) cudaFree(q);
    cudaFree(p->d_soft_stage); cudaFree(p->d_count_stage);
    cudaFreeHost(p->h_state); cudaFreeHost(p->h_su);
    if (p->ev_batch) cudaEventDestroy(p->ev_batch);
    if (p->ev_own) cudaEventDestroy(p->ev_own);
    if (p->stream) cudaStreamDestroy(p->stream);
    delete p;
}
int64_t jaero_pchannel_launch_count(const jaero_pchannel *p) { return p ? p->launches : 0; }
int jaero_pchannel_su_capacity(const jaero_pchannel *p) { return p ? p->pp.su_cap : 0; }

int jaero_pchannel_process_batch(jaero_pchannel *p, jaero_batch *b)
{
    if (!p || !b || p->pp.n_channels != b->p.n_channels || p->device != b->device) { set_error("jaero_pchannel_process_batch: batch mismatch"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(p->device));
    const DemodParams &dp = b->p;
    int *dcd = dp.I + (size_t)I_DCD * dp.cpad;
    // everything runs on the batch's stream so it is ordered after the demodulator segments
    if (pc_enter(p, b->stream)) return JAERO_E_CUDA;
    if (pchan_process(p->pp, dp.soft, dp.I + (size_t)I_SOFT_COUNT * dp.cpad, dp.soft_cap, dcd, p->vit_overlap, p->vit_overlap_len,
                      p->vit_renorm, p->vit_valid, p->pp.queue, b->stream, &p->launches,
                      dp.I + (size_t)I_LOST_N * dp.cpad, dp.lost_pos, dp.cpad)) return JAERO_E_CUDA;
    soft_reset_kernel<<<(dp.n_channels + 127) / 128, 128, 0, b->stream>>>(dp);
    JB_CUDA(cudaGetLastError());
    p->launches++;
    return pc_leave(p, b->stream) ? JAERO_E_CUDA : JAERO_OK;
}
int jaero_pchannel_process_softbits(jaero_pchannel *p, const int16_t *soft, size_t cap, const int32_t *counts)
{
    if (!p || !soft || !counts || cap == 0) { set_error("jaero_pchannel_process_softbits: bad argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(p->device));
    const size_t C = p->pp.n_channels;
    if (C * cap > p->stage_cap) {
        JB_CUDA(cudaStreamSynchronize(p->stream));
        cudaFree(p->d_soft_stage); cudaFree(p->d_count_stage); p->d_soft_stage = 0; p->d_count_stage = 0;
        JB_CUDA(cudaMalloc(&p->d_soft_stage, C * cap * sizeof(int16_t)));
        JB_CUDA(cudaMalloc(&p->d_count_stage, C * sizeof(int)));
        p->stage_cap = C * cap;
    }
    p->own_dirty = true;
    JB_CUDA(cudaMemcpyAsync(p->d_soft_stage, soft, C * cap * sizeof(int16_t), cudaMemcpyHostToDevice, p->stream));
    JB_CUDA(cudaMemcpyAsync(p->d_count_stage, counts, C * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    if (pchan_process(p->pp, p->d_soft_stage, p->d_count_stage, (int)cap, nullptr, p->vit_overlap, p->vit_overlap_len,
                      p->vit_renorm, p->vit_valid, p->pp.queue, p->stream, &p->launches)) return JAERO_E_CUDA;
    JB_CUDA(cudaStreamSynchronize(p->stream));
    return JAERO_OK;
}
int jaero_pchannel_tick(jaero_pchannel *p, jaero_batch *b)
{
    if (!p || (b && b->p.n_channels != p->pp.n_channels)) { set_error("jaero_pchannel_tick: bad argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(p->device));
    if (b) { if (pc_enter(p, b->stream)) return JAERO_E_CUDA; } else p->own_dirty = true;
    if (pchan_tick(p->pp, b ? b->p.I + (size_t)I_DCD * b->p.cpad : nullptr, b ? b->stream : p->stream)) return JAERO_E_CUDA;
    p->launches++;
    return (b && pc_leave(p, b->stream)) ? JAERO_E_CUDA : JAERO_OK;
}
// AeroL::SignalStatusSlot(false) -> LostSignal() (aerol.h:920-931): cntr = 1e9, DCD countdown and DCD cleared at once, and
// DataCarrierDetect(false) reaches the demodulator (b may be NULL: frame layer only). channel -1 = every channel.
int jaero_pchannel_lost_signal(jaero_pchannel *p, jaero_batch *b, int channel)
{
    if (!p || channel >= p->pp.n_channels || (b && b->p.n_channels != p->pp.n_channels)) { set_error("jaero_pchannel_lost_signal: bad argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(p->device));
    if (b) { if (pc_enter(p, b->stream)) return JAERO_E_CUDA; } else p->own_dirty = true;
    if (pchan_lost(p->pp, channel, b ? b->p.I + (size_t)I_DCD * b->p.cpad : nullptr, b ? b->stream : p->stream)) return JAERO_E_CUDA;
    p->launches++;
    return (b && pc_leave(p, b->stream)) ? JAERO_E_CUDA : JAERO_OK;
}

// Length of the piece of a write that ends in front of the next coarse-estimator trigger sample (the only points where the OQPSK
// demodulator reads DCD and where SignalStatus is emitted): the lock-step counters of jaero_batch_write_device, read-only.
static size_t piece_before_next_trigger(const jaero_batch *b, size_t n)
{
    const DemodParams &p = b->p;
    const int N = p.bbnfft, trig_every = p.cpu_reduce ? N : N / 4;
    int bb = b->bb_pos, cc = b->coarse_counter;
    for (size_t i = 0; i < n; i++) {
        if (cc >= p.Fs || !p.cpu_reduce) {
            bb++; if (bb >= N) bb = 0;
            if (bb % trig_every == 0) { if (i > 0) return i; cc = 0; }   // a trigger on the first sample opens this piece
        }
        cc++;
    }
    return n;
}
// writeData with the AeroL attached the way JAERO/mainwindow.cpp:198-237,432,508 wires them: the stream is cut in front of
// every estimator trigger sample and the frame layer runs at each cut, so that the DCD the demodulator reads in
// FreqOffsetEstimateSlot, and the LostSignal that follows a SignalStatus(false), see exactly the soft bits emitted before
// that sample (exact for OQPSK, whose only reads of DCD are in that slot; for MSK the timing-loop gain switches at the next
// cut, at most one estimator epoch after the reference's emit-granular switch). HOST pcm.
int jaero_pchannel_write_batch(jaero_pchannel *p, jaero_batch *b, const int16_t *pcm, size_t n, size_t stride)
{
    if (!p || !b || !pcm || p->pp.n_channels != b->p.n_channels || p->device != b->device) { set_error("jaero_pchannel_write_batch: bad argument"); return JAERO_E_ARG; }
    if (stride < n) { set_error("jaero_pchannel_write_batch: channel_stride < n_samples"); return JAERO_E_ARG; }
    b->p.wire_sigstat = 1;
    size_t done = 0;
    while (done < n) {
        const size_t k = piece_before_next_trigger(b, n - done);   // >= 1: up to, not including, the next trigger sample
        int rc = jaero_batch_write(b, pcm + done, k, stride);
        if (rc) return rc;
        rc = jaero_pchannel_process_batch(p, b);
        if (rc) return rc;
        done += k;
    }
    return JAERO_OK;
}
static int pc_pull_state(jaero_pchannel *p)
{
    JB_CUDA(cudaMemcpyAsync(p->h_state, p->pp.state, (size_t)p->pp.n_channels * sizeof(PChanState), cudaMemcpyDeviceToHost, p->stream));
    JB_CUDA(cudaStreamSynchronize(p->stream));          // p->stream is ordered behind every batch-stream call (pc_leave)
    return 0;
}
int jaero_pchannel_read_sus(jaero_pchannel *p, uint8_t *out, size_t cap, int32_t *counts)
{
    if (!p || !out || !counts) { set_error("jaero_pchannel_read_sus: null argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(p->device));
    if (pc_pull_state(p)) return JAERO_E_CUDA;
    const PChanParams &pp = p->pp;
    bool overflow = false; int maxc = 0;
    for (int ch = 0; ch < pp.n_channels; ch++) { maxc = std::max(maxc, p->h_state[ch].su_count); overflow |= p->h_state[ch].queue_overflow != 0 || (size_t)p->h_state[ch].su_count > cap; }
    if (overflow) { set_error("P-channel queue overflow: call process/read more often"); return JAERO_E_OVERFLOW; }
    if (maxc) { JB_CUDA(cudaMemcpyAsync(p->h_su, pp.su_out, (size_t)pp.n_channels * pp.su_cap * 16, cudaMemcpyDeviceToHost, p->stream)); JB_CUDA(cudaStreamSynchronize(p->stream)); }
    for (int ch = 0; ch < pp.n_channels; ch++) {
        counts[ch] = p->h_state[ch].su_count;
        if (counts[ch]) memcpy(out + (size_t)ch * cap * 16, p->h_su + (size_t)ch * pp.su_cap * 16, (size_t)counts[ch] * 16);
    }
    pchan_su_reset_kernel<<<(pp.n_channels + 127) / 128, 128, 0, p->stream>>>(pp);
    JB_CUDA(cudaGetLastError());
    p->own_dirty = true;
    return JAERO_OK;
}
int jaero_pchannel_discard_sus(jaero_pchannel *p)
{
    if (!p) { set_error("null handle"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(p->device));
    pchan_su_reset_kernel<<<(p->pp.n_channels + 127) / 128, 128, 0, p->stream>>>(p->pp);
    JB_CUDA(cudaGetLastError());
    p->own_dirty = true;
    p->launches++;
    return JAERO_OK;
}
int jaero_pchannel_get_stats(jaero_pchannel *p, int32_t *dcd, int64_t *su_total, int64_t *su_ok)
{
    if (!p) { set_error("null handle"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(p->device));
    if (pc_pull_state(p)) return JAERO_E_CUDA;
    for (int ch = 0; ch < p->pp.n_channels; ch++) {
        if (dcd) dcd[ch] = p->h_state[ch].datacd;
        if (su_total) su_total[ch] = p->h_state[ch].su_total;
        if (su_ok) su_ok[ch] = p->h_state[ch].su_ok;
    }
    return JAERO_OK;
}

} // extern "C"

// ====================================================================== burst MSK demodulator
#include "burst.cuh"
namespace jb { int burst_trident_fft_launch(const BurstParams &p, const int *d_ev_list, int n_events, double2 *wa, double2 *wb, const double2 *tw, cudaStream_t s); }

struct jaero_burst {
    int device; cudaStream_t stream;
    BurstParams p; HilbertStream hil;
    std::vector<void *> allocs;
    long long samples; int hil_fill; long long hil_blocks;
    int16_t *d_stage; size_t stage_cap;
    double2 *tw32k, *wa, *wb; int *d_ev_list; int ev_round;
    int *h_ints; double *h_dbls; int16_t *h_soft;
    std::vector<int> h_ev;
    long long launches;
};

namespace {
template <class T> int bu_alloc(jaero_burst *b, T **ptr, size_t count)
{
    int r = dev_alloc_zero(ptr, count, b->stream);
    if (r == 0) b->allocs.push_back((void *)*ptr);
    return r;
}
bool delay_w1(double fd, int *k, double *w)          // Delay<T> weight at ring position 0 (DSP.h:357-374); integer delays -> 0
{
    const int size = (int)std::ceil(fd) + 1;
    double w0 = 0;
    for (int bp = 0; bp < size; bp++) {
        double dptr = ((double)bp) - fd;
        while (std::floor(dptr) < 0) dptr += ((double)size);
        const double ww = dptr - std::floor(dptr);
        if (bp == 0) w0 = ww; else if (ww != w0) return false;
    }
    *k = (int)std::ceil(fd); *w = w0;
    return true;
}
// Delay<T> weight at every ring position (DSP.h:357-374); the ring read position must be "ceil(fd) samples ago"
bool delay_table(double fd, std::vector<double> &w, int *k)
{
    const int size = (int)std::ceil(fd) + 1;
    w.assign(size, 0.0);
    for (int bp = 0; bp < size; bp++) {
        double dptr = ((double)bp) - fd;
        while (std::floor(dptr) < 0) dptr += ((double)size);
        const int iptr = (int)std::floor(dptr);
        w[bp] = dptr - ((double)iptr);
        int expect = bp - (int)std::ceil(fd); while (expect < 0) expect += size;
        if (iptr != expect) return false;
    }
    *k = (int)std::ceil(fd);
    return true;
}
__global__ void burst_init_kernel(BurstParams p, double freq_center, double st_freq)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.cpad) return;
    auto D = [&](int i) -> double & { return p.BD[(size_t)i * p.cpad + ch]; };
    auto I = [&](int i) -> int & { return p.BI[(size_t)i * p.cpad + ch]; };
    const double sr = (double)((float)((int)p.Fs));
    D(BD_M2_FREQ) = freq_center; D(BD_M2_STEP) = (freq_center) * ((double)jb::WTSIZE) / sr;
    D(BD_MC_FREQ) = freq_center; D(BD_MC_STEP) = (freq_center) * ((double)jb::WTSIZE) / sr;
    D(BD_ST_FREQ) = st_freq; D(BD_ST_STEP) = (st_freq) * ((double)jb::WTSIZE) / sr;
    D(BD_SH_FREQ) = st_freq; D(BD_SH_STEP) = (st_freq) * ((double)jb::WTSIZE) / sr;
    D(BD_MSE) = 10.0;                                    // burstmskdemodulator.cpp:195
    D(BD_ROT_RE) = 1.0; D(BD_SAV_RE) = 1.0;              // rotator=1, symboltone_averotator=1 (:201-202); symboltone_rotator stays 0
    D(BD_DIFF_LAST) = -1.0;
    if (p.kind == 1) {                                   // burst OQPSK ctor (burstoqpskdemodulator.cpp:4-133)
        D(BD_MSE) = 100.0; D(BD_VOL_GAIN) = 1.0; D(BD_STR_RE) = 1.0;     // symboltone_rotator=1, never reset
        D(BD_ST_FREQ) = 10500.0; D(BD_ST_STEP) = (10500.0) * ((double)jb::WTSIZE) / sr;
        D(BD_SR_FREQ) = 10500.0; D(BD_SR_STEP) = (10500.0) * ((double)jb::WTSIZE) / sr;
        D(BD_SH_FREQ) = 10500.0 / 4.0; D(BD_SH_STEP) = (10500.0 / 4.0) * ((double)jb::WTSIZE) / sr;
    }
    I(BI_PD_CNTDOWN) = 2 * p.pd_len; I(BI_PD_MAXPOSCNT) = -1;    // PeakDetector::setSettings (DSP.h:502-513)
    I(BI_STARTSTOP) = -1;                                // ctor :69
}
__global__ void burst_soft_reset_kernel(BurstParams p)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_channels) return;
    int &count = p.BI[(size_t)BI_SOFT_COUNT * p.cpad + ch];
    const int pending = p.BI[(size_t)BI_SOFT_PENDING * p.cpad + ch];
    int16_t *ring = p.soft + (size_t)ch * p.soft_cap;
    for (int k = 0; k < pending; k++) ring[k] = ring[count + k];
    count = 0;
}
__global__ void burst_set_int_kernel(BurstParams p, int idx, int channel, int value)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_channels) return;
    if (channel < 0 || channel == ch) p.BI[(size_t)idx * p.cpad + ch] = value;
}
} // namespace

extern "C" {

static int burst_create(const jaero_settings *s, int n_channels, int device, int kind, jaero_burst **out)
{
    if (!s || !out || n_channels <= 0) { set_error("jaero_burst_create: bad argument"); return JAERO_E_ARG; }
    if (kind == 0 && (s->Fs != 48000 || (s->fb != 600 && s->fb != 1200))) { set_error("jaero_burst_msk_create: burst MSK runs at Fs=48000 with fb 600 or 1200"); return JAERO_E_ARG; }
    if (kind == 1 && (s->Fs != 48000 || s->fb != 10500)) { set_error("jaero_burst_oqpsk_create: burst OQPSK runs at Fs=48000, fb=10500"); return JAERO_E_ARG; }
    int ndev = 0;
    JB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("jaero_burst_msk_create: no such CUDA device"); return JAERO_E_CUDA; }
    JB_CUDA(cudaSetDevice(device));
    jaero_burst *b = new (std::nothrow) jaero_burst();
    if (!b) { set_error("out of host memory"); return JAERO_E_ARG; }
    CreateGuard<jaero_burst> guard(b, jaero_burst_destroy);
    b->device = device; b->samples = 0; b->hil_fill = 0; b->hil_blocks = 0; b->d_stage = 0; b->stage_cap = 0; b->launches = 0;
    JB_CUDA(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    BurstParams &p = b->p;
    memset(&p, 0, sizeof p);
    p.kind = kind; p.sql = s->sql;
    p.n_channels = n_channels; p.cpad = (n_channels + 31) & ~31;
    p.Fs = s->Fs; p.fb = s->fb; p.lockingbw = s->lockingbw; p.signalthreshold = s->signalthreshold; p.afc = 1;   // ctor: afc=true (:15)
    double fc = s->freq_center;
    if (fc > ((p.Fs / 2.0) - (p.lockingbw / 2.0))) fc = ((p.Fs / 2.0) - (p.lockingbw / 2.0));
    p.sps = (int)(p.Fs / p.fb);
    const double SPS = kind == 1 ? 2.0 * p.Fs / p.fb : (double)p.sps;            // burstoqpskdemodulator.cpp:219
    p.spsd = SPS;
    std::vector<double> taps;
    if (kind == 1) { p.sps = (int)SPS; p.ntaps = 55; taps = rrc_taps(1.0, 55, 48000, 10500 / 2.0); }    // ctor :38-46
    else {
        p.ntaps = 2 * p.sps;
        if (p.ntaps > MAX_TAPS) { set_error("burst MSK: matched filter too long"); return JAERO_E_ARG; }
        taps.resize(p.ntaps);
        for (int i = 0; i < p.ntaps; i++) taps[i] = sin(M_PI * i / (2.0 * SPS)) / (2.0 * SPS);      // :173-177
    }
    for (int k = 0; k < p.ntaps; k++) p.taps[k] = taps[k];
    p.agc_len = (int)round(1 * p.Fs);
    auto qround = [](double d) { return d >= 0.0 ? int(d + 0.5) : int(d - double(int(d - 1)) + 0.5) + int(d - 1); };
    double btdiff_fd = 0;
    if (kind == 1) {                                              // burstoqpskdemodulator.cpp:202-277
        p.btma_len = qround(128.0 * SPS); p.mav1_len = (int)(SPS * 128); btdiff_fd = SPS * 128; p.btdiff_len = (int)std::ceil(btdiff_fd) + 1;
        p.pd_len = (int)(SPS * 128.0 / 2.0); p.pd_threshold = 0.2;
        p.tri_sz = qround((256.0 + 16.0 + 16.0) * SPS); p.d1_len = (int)(SPS * 128.0 * 2.5 - 190) + 1; p.d2_len = p.tri_sz + 1;
        p.tri_nb = p.tri_nt = qround(128.0 * SPS);
        p.res_b0 = 0.0048847995518126464; p.res_b1 = 0; p.res_b2 = -0.0048847995518126464;       // ctor :69-75 (75 Hz)
        p.res_a1 = -0.3882746897971619; p.res_a2 = 0.99023040089637471;
        p.ee = 0.4;
    } else if (p.fb >= 1200) {                                           // :205-256
        p.btma_len = qround(126.0 * SPS); p.mav1_len = (int)(SPS * 126); p.btdiff_len = (int)std::ceil(SPS * 126) + 1;
        p.pd_len = (int)(SPS * 126.0 / 2.0); p.pd_threshold = 0.1;
        p.tri_sz = qround(200.0 * SPS); p.d1_len = ((int)289 * p.sps) + 20 + 1; p.d2_len = (int)(qround(72 + 120.0) * SPS) + 1;
        p.size_base = 126; p.size_top = 74; p.start_processing = 120; p.end_rotation = (int)((120 + 37) * SPS);
        p.res_a1 = -1.993312819378528; p.res_a2 = 0.999476538254407; p.res_b0 = 2.617308727964618e-04; p.res_b1 = 0; p.res_b2 = -2.617308727964618e-04;
        p.ee = 0.025; btdiff_fd = SPS * 126;
    } else {                                                      // :257-311
        p.btma_len = qround(150.0 * SPS); p.mav1_len = (int)(SPS * 150); p.btdiff_len = (int)std::ceil(SPS * 150) + 1;
        p.pd_len = (int)(SPS * 150.0 / 2.0); p.pd_threshold = 0.2;
        p.tri_sz = qround(224 * SPS); p.d1_len = ((int)397 * p.sps) + 20 + 1; p.d2_len = qround((72 + 150.0) * SPS) + 1;
        p.size_base = 150; p.size_top = 74; p.start_processing = 150; p.end_rotation = (int)((150 + 56) * SPS);
        p.res_a1 = -1.991228154418550; p.res_a2 = 0.997385427096603; p.res_b0 = 0.001307286451699; p.res_b1 = 0; p.res_b2 = -0.001307286451699;
        p.ee = 0.015; btdiff_fd = SPS * 150;
    }
    if (kind == 0) { p.tri_nb = (int)rint(p.size_base * SPS); p.tri_nt = (int)rint(p.size_top * SPS); }
    p.startstopstart = kind == 1 ? (int)(SPS * (1050)) : (int)(SPS * 500);
    p.btd1_len = (int)std::ceil(1.0 * SPS) + 1;
    int kk;
    std::vector<double> w_btd1, w_btdiff, w_a1, w_tmp;
    if (!delay_table(1.0 * SPS, w_btd1, &kk) || !delay_table(btdiff_fd, w_btdiff, &kk) || !delay_table(SPS / 2.0, w_a1, &p.a1_k)) {
        set_error("burst: unsupported delay"); return JAERO_E_ARG; }
    if (kind == 0) {
        if (!delay_w1(SPS / 2.0, &p.d8_k, &p.d8_w)) { set_error("burst MSK: unsupported delay"); return JAERO_E_ARG; }
        p.eb_len = (int)(0.15 * p.Fs); p.agc2_len = (int)round((SPS * 128.0 / p.Fs) * p.Fs); p.ds_len = p.sps + 1; p.msema_len = 75;
    } else {
        const double sps0 = 2.0 * 48000 / 10500;                  // ctor :48-52
        if (!delay_weights(sps0 / 4.0, &p.k41, p.w41v) || !delay_weights(sps0 / 8.0, &p.k8, p.w8v)) { set_error("burst OQPSK: unsupported delay"); return JAERO_E_ARG; }
        p.d8_k = 1; p.ds_len = 1;
        p.eb_len = (int)(SPS * (256.0)); p.agc2_len = (int)round((SPS * 64.0 / p.Fs) * p.Fs); p.msema_len = 128;
    }
    p.soft_cap = std::max(4096, (int)(2 * p.fb) + 64);
    const size_t cp = p.cpad, C = n_channels;
    int rc = 0;
    rc |= bu_alloc(b, &p.BD, (size_t)BD_COUNT * cp); rc |= bu_alloc(b, &p.BI, (size_t)BI_COUNT * cp);
    rc |= bu_alloc(b, &p.agc_ring, (size_t)p.agc_len * cp); rc |= bu_alloc(b, &p.d1_ring, (size_t)p.d1_len * cp);
    rc |= bu_alloc(b, &p.d2_ring, (size_t)p.d2_len * cp); rc |= bu_alloc(b, &p.btd1_ring, (size_t)p.btd1_len * cp);
    rc |= bu_alloc(b, &p.btma_ring, (size_t)p.btma_len * cp); rc |= bu_alloc(b, &p.mav1_ring, (size_t)p.mav1_len * cp);
    rc |= bu_alloc(b, &p.btdiff_ring, (size_t)p.btdiff_len * cp);
    rc |= bu_alloc(b, &p.pd1_ring, (size_t)(2 * p.pd_len + 1) * cp); rc |= bu_alloc(b, &p.pd2_ring, (size_t)(p.pd_len + 1) * cp);
    rc |= bu_alloc(b, &p.pd3_ring, (size_t)(2 * p.pd_len + 1) * cp);
    rc |= bu_alloc(b, &p.a1_ring, (size_t)(p.a1_k + 1) * cp); rc |= bu_alloc(b, &p.eb1_ring, (size_t)p.eb_len * cp);
    rc |= bu_alloc(b, &p.eb2_ring, (size_t)p.eb_len * cp); rc |= bu_alloc(b, &p.agc2_ring, (size_t)p.agc2_len * cp);
    rc |= bu_alloc(b, &p.d8_ring, (size_t)(p.d8_k + 1) * cp); rc |= bu_alloc(b, &p.msema_ring, (size_t)p.msema_len * cp);
    rc |= bu_alloc(b, &p.fir_re, (size_t)(p.ntaps + 1) * cp); rc |= bu_alloc(b, &p.fir_im, (size_t)(p.ntaps + 1) * cp);
    rc |= bu_alloc(b, &p.ds_ring, (size_t)p.ds_len * cp);
    rc |= bu_alloc(b, &p.tri, C * BURST_MAXEV * p.tri_sz); rc |= bu_alloc(b, &p.ev_sample, C * BURST_MAXEV);
    rc |= bu_alloc(b, &p.ev_result, C * BURST_MAXEV * 8);
    p.astride = BURST_CHUNK;
    rc |= bu_alloc(b, &p.analytic, C * p.astride); rc |= bu_alloc(b, &p.vtd, C * p.astride);
    rc |= bu_alloc(b, &p.soft, C * p.soft_cap);
    {
        double *d1 = 0, *d2 = 0, *d3 = 0;
        rc |= bu_alloc(b, &d1, w_btd1.size()); rc |= bu_alloc(b, &d2, w_btdiff.size()); rc |= bu_alloc(b, &d3, w_a1.size());
        if (!rc) {
            JB_CUDA(cudaMemcpyAsync(d1, w_btd1.data(), w_btd1.size() * 8, cudaMemcpyHostToDevice, b->stream));
            JB_CUDA(cudaMemcpyAsync(d2, w_btdiff.data(), w_btdiff.size() * 8, cudaMemcpyHostToDevice, b->stream));
            JB_CUDA(cudaMemcpyAsync(d3, w_a1.data(), w_a1.size() * 8, cudaMemcpyHostToDevice, b->stream));
            JB_CUDA(cudaStreamSynchronize(b->stream));
        }
        p.btd1_wv = d1; p.btdiff_wv = d2; p.a1_wv = d3;
    }
    // Hilbert filter: QJHilbertFilter::setSize(2048) (DSP.cpp:759-789), streaming FFT convolution nfft 8192
    HilbertStream &h = b->hil; memset(&h, 0, sizeof h);
    h.K = 2048; h.nfft = 8192; h.L = h.nfft - h.K + 1;
    rc |= bu_alloc(b, &h.H, (size_t)h.nfft); rc |= bu_alloc(b, &h.tw, (size_t)h.nfft);
    rc |= bu_alloc(b, &h.hist, C * (h.K - 1)); rc |= bu_alloc(b, &h.inblk, C * h.L); rc |= bu_alloc(b, &h.outblk, C * h.L);
    b->ev_round = 128;
    rc |= bu_alloc(b, &b->tw32k, (size_t)TRI_N); rc |= bu_alloc(b, &b->wa, (size_t)2 * b->ev_round * TRI_N); rc |= bu_alloc(b, &b->wb, (size_t)2 * b->ev_round * TRI_N);
    rc |= bu_alloc(b, &b->d_ev_list, (size_t)2 * C * BURST_MAXEV);
    if (rc) return JAERO_E_CUDA;
    {
        std::vector<double> sn(jb::WTSIZE), cs(jb::WTSIZE);
        for (int i = 0; i < jb::WTSIZE; i++) sn[i] = (sin(2 * M_PI * ((double)i) / jb::WTSIZE));
        for (int i = 0; i < jb::WTSIZE; i++) cs[i] = (sin(M_PI_2 + 2 * M_PI * ((double)i) / jb::WTSIZE));
        double *ds, *dc;
        if (bu_alloc(b, &ds, (size_t)jb::WTSIZE) || bu_alloc(b, &dc, (size_t)jb::WTSIZE)) return JAERO_E_CUDA;
        JB_CUDA(cudaMemcpyAsync(ds, sn.data(), sn.size() * 8, cudaMemcpyHostToDevice, b->stream));
        JB_CUDA(cudaMemcpyAsync(dc, cs.data(), cs.size() * 8, cudaMemcpyHostToDevice, b->stream));
        p.sin_t = ds; p.cos_t = dc;
        // Hilbert kernel and its spectrum (host radix-2), twiddle tables
        const int NF = h.nfft;
        std::vector<std::complex<double>> Hk(NF, 0.0), tw(NF), tw32(TRI_N);
        const int N = 2048;
        for (int i = 0; i < N; i++) {
            if (i == N / 2) Hk[i] = std::complex<double>(-1, 0);
            else if ((i % 2) == 0) Hk[i] = 0;
            else Hk[i] = std::complex<double>(0, (2.0 / ((double)N)) / (std::tan(M_PI * (((double)i) / ((double)N) - 0.5))));
        }
        for (int k = 0; k < NF; k++) { const double a = -2.0 * M_PI * (double)k / (double)NF; tw[k] = std::complex<double>(cos(a), sin(a)); }
        for (int k = 0; k < TRI_N; k++) { const double a = -2.0 * M_PI * (double)k / (double)TRI_N; tw32[k] = std::complex<double>(cos(a), sin(a)); }
        {
            int bits = 13;
            for (int i = 0; i < NF; i++) { int r = 0; for (int q = 0; q < bits; q++) if (i & (1 << q)) r |= 1 << (bits - 1 - q); if (r > i) std::swap(Hk[i], Hk[r]); }
            for (int len = 2; len <= NF; len <<= 1)
                for (int i = 0; i < NF; i += len)
                    for (int k = 0; k < len / 2; k++) { auto w = tw[k * (NF / len)]; auto u = Hk[i + k], v = Hk[i + k + len / 2] * w; Hk[i + k] = u + v; Hk[i + k + len / 2] = u - v; }
        }
        JB_CUDA(cudaMemcpyAsync(h.H, Hk.data(), NF * 16, cudaMemcpyHostToDevice, b->stream));
        JB_CUDA(cudaMemcpyAsync(h.tw, tw.data(), NF * 16, cudaMemcpyHostToDevice, b->stream));
        JB_CUDA(cudaMemcpyAsync(b->tw32k, tw32.data(), (size_t)TRI_N * 16, cudaMemcpyHostToDevice, b->stream));
        JB_CUDA(cudaStreamSynchronize(b->stream));
    }
    burst_init_kernel<<<(p.cpad + 127) / 128, 128, 0, b->stream>>>(p, fc, p.fb / 2.0);
    JB_CUDA(cudaGetLastError());
    JB_CUDA(cudaStreamSynchronize(b->stream));
    JB_CUDA(cudaMallocHost(&b->h_ints, (size_t)BI_COUNT * cp * sizeof(int)));
    JB_CUDA(cudaMallocHost(&b->h_dbls, (size_t)BD_COUNT * cp * sizeof(double)));
    JB_CUDA(cudaMallocHost(&b->h_soft, C * p.soft_cap * sizeof(int16_t)));
    guard.release();
    *out = b;
    return JAERO_OK;
}
int jaero_burst_msk_create(const jaero_settings *s, int n_channels, int device, jaero_burst **out) { return burst_create(s, n_channels, device, 0, out); }
int jaero_burst_oqpsk_create(const jaero_settings *s, int n_channels, int device, jaero_burst **out) { return burst_create(s, n_channels, device, 1, out); }
void jaero_burst_destroy(jaero_burst *b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    cudaStreamSynchronize(b->stream);
    for (void *q : b->allocs) cudaFree(q);
    cudaFree(b->d_stage);
    cudaFreeHost(b->h_ints); cudaFreeHost(b->h_dbls); cudaFreeHost(b->h_soft);
    if (b->stream) cudaStreamDestroy(b->stream);
    delete b;
}
int64_t jaero_burst_launch_count(const jaero_burst *b) { return b ? b->launches : 0; }
int jaero_burst_sync(jaero_burst *b)
{
    if (!b) { set_error("null handle"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    JB_CUDA(cudaStreamSynchronize(b->stream));
    return JAERO_OK;
}
int jaero_burst_write_device(jaero_burst *b, const int16_t *d_pcm, size_t n, size_t stride)
{
    if (!b || !d_pcm) { set_error("jaero_burst_write_device: null argument"); return JAERO_E_ARG; }
    if (n == 0) return JAERO_OK;
    if (stride < n || n > 0x7fffffff) { set_error("jaero_burst_write_device: bad stride / length"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    const BurstParams &p = b->p;
    const size_t cp = p.cpad;
    int new_write = 1;
    for (size_t c0 = 0; c0 < n; c0 += BURST_CHUNK) {
        const int m = (int)std::min((size_t)BURST_CHUNK, n - c0);
        // Hilbert transform of this chunk (JFastFir::update: per-sample exchange + block transforms)
        int i = 0;
        while (i < m) {
            const int take = std::min(b->hil.L - b->hil_fill, m - i);
            if (hilbert_exchange_launch(b->hil, p, d_pcm, stride, (int)c0, i, i + take, b->hil_fill, b->stream)) return JAERO_E_CUDA;
            b->launches++;
            b->hil_fill += take; i += take;
            if (b->hil_fill == b->hil.L) {
                if (hilbert_block_launch(b->hil, p.n_channels, b->hil_blocks == 0 ? 1 : 0, b->stream)) return JAERO_E_CUDA;
                b->launches++; b->hil_fill = 0; b->hil_blocks++;
            }
        }
        if (burst_front_launch(p, b->samples, m, b->stream)) return JAERO_E_CUDA;
        b->launches++;
        // trident events of this chunk
        JB_CUDA(cudaMemcpyAsync(b->h_ints, p.BI + (size_t)BI_NEV * cp, cp * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
        JB_CUDA(cudaStreamSynchronize(b->stream));
        b->h_ev.clear();
        for (int ch = 0; ch < p.n_channels; ch++) for (int e = 0; e < b->h_ints[ch]; e++) { b->h_ev.push_back(ch); b->h_ev.push_back(e); }
        const int nevt = (int)b->h_ev.size() / 2;
        if (nevt) JB_CUDA(cudaMemcpyAsync(b->d_ev_list, b->h_ev.data(), b->h_ev.size() * sizeof(int), cudaMemcpyHostToDevice, b->stream));
        for (int e0 = 0; e0 < nevt; e0 += b->ev_round) {
            const int cnt = std::min(b->ev_round, nevt - e0);
            if (burst_trident_fft_launch(p, b->d_ev_list + 2 * e0, cnt, b->wa, b->wb, b->tw32k, b->stream)) return JAERO_E_CUDA;
            b->launches += 2;
        }
        if (burst_back_launch(p, b->samples, m, new_write, b->stream)) return JAERO_E_CUDA;
        new_write = 0;
        b->launches++;
        b->samples += m;
    }
    return JAERO_OK;
}
int jaero_burst_write(jaero_burst *b, const int16_t *pcm, size_t n, size_t stride)
{
    if (!b || !pcm) { set_error("jaero_burst_write: null argument"); return JAERO_E_ARG; }
    if (n == 0) return JAERO_OK;
    if (stride < n) { set_error("jaero_burst_write: channel_stride < n_samples"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    const size_t C = b->p.n_channels, pitch = (n + 7) & ~(size_t)7;
    if (C * pitch > b->stage_cap) {
        JB_CUDA(cudaStreamSynchronize(b->stream));
        cudaFree(b->d_stage); b->d_stage = 0;
        JB_CUDA(cudaMalloc(&b->d_stage, C * pitch * sizeof(int16_t)));
        b->stage_cap = C * pitch;
    }
    JB_CUDA(cudaMemcpy2DAsync(b->d_stage, pitch * 2, pcm, stride * 2, n * 2, C, cudaMemcpyHostToDevice, b->stream));
    return jaero_burst_write_device(b, b->d_stage, n, pitch);
}
int jaero_burst_read_softbits(jaero_burst *b, int16_t *out, size_t cap, int32_t *counts)
{
    if (!b || !out || !counts) { set_error("jaero_burst_read_softbits: null argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    const BurstParams &p = b->p; const size_t cp = p.cpad;
    JB_CUDA(cudaMemcpyAsync(b->h_ints, p.BI, (size_t)BI_COUNT * cp * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    JB_CUDA(cudaStreamSynchronize(b->stream));
    const int *cnt = b->h_ints + (size_t)BI_SOFT_COUNT * cp, *ovf = b->h_ints + (size_t)BI_SOFT_OVERFLOW * cp;
    int maxc = 0; bool overflow = false;
    for (int ch = 0; ch < p.n_channels; ch++) { maxc = std::max(maxc, cnt[ch]); overflow |= ovf[ch] != 0 || (size_t)cnt[ch] > cap; }
    if (overflow) { set_error("soft-bit ring overflow: drain more often or pass a larger buffer"); return JAERO_E_OVERFLOW; }
    if (maxc) {
        JB_CUDA(cudaMemcpy2DAsync(b->h_soft, (size_t)p.soft_cap * 2, p.soft, (size_t)p.soft_cap * 2, (size_t)maxc * 2, p.n_channels, cudaMemcpyDeviceToHost, b->stream));
        JB_CUDA(cudaStreamSynchronize(b->stream));
    }
    for (int ch = 0; ch < p.n_channels; ch++) { counts[ch] = cnt[ch]; if (cnt[ch]) memcpy(out + (size_t)ch * cap, b->h_soft + (size_t)ch * p.soft_cap, (size_t)cnt[ch] * 2); }
    burst_soft_reset_kernel<<<(p.n_channels + 127) / 128, 128, 0, b->stream>>>(p);
    JB_CUDA(cudaGetLastError());
    return JAERO_OK;
}
int jaero_burst_set_dcd(jaero_burst *b, int channel, int dcd)
{
    if (!b || channel >= b->p.n_channels) { set_error("jaero_burst_set_dcd: bad argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    burst_set_int_kernel<<<(b->p.n_channels + 127) / 128, 128, 0, b->stream>>>(b->p, BI_DCD, channel, dcd ? 1 : 0);
    JB_CUDA(cudaGetLastError());
    return JAERO_OK;
}
int jaero_burst_set_afc(jaero_burst *b, int state)                  // burstmskdemodulator.cpp / burstoqpskdemodulator.cpp setAFC
{
    if (!b) { set_error("null handle"); return JAERO_E_ARG; }
    b->p.afc = state ? 1 : 0;
    return JAERO_OK;
}
int jaero_burst_set_sql(jaero_burst *b, int state)
{
    if (!b) { set_error("null handle"); return JAERO_E_ARG; }
    b->p.sql = state ? 1 : 0;
    return JAERO_OK;
}
int jaero_burst_get_status_all(jaero_burst *b, jaero_burst_status *out)
{
    if (!b || !out) { set_error("null argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(b->device));
    const size_t cp = b->p.cpad;
    JB_CUDA(cudaMemcpyAsync(b->h_dbls, b->p.BD, (size_t)BD_COUNT * cp * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    JB_CUDA(cudaMemcpyAsync(b->h_ints, b->p.BI, (size_t)BI_COUNT * cp * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    JB_CUDA(cudaStreamSynchronize(b->stream));
    for (int ch = 0; ch < b->p.n_channels; ch++) {
        auto D = [&](int i) { return b->h_dbls[(size_t)i * cp + ch]; };
        auto I = [&](int i) { return b->h_ints[(size_t)i * cp + ch]; };
        jaero_burst_status &s = out[ch];
        s.mixer2_freq = D(BD_M2_FREQ); s.mixer2_wtptr = D(BD_M2_PTR); s.center_freq = D(BD_MC_FREQ); s.st_freq = D(BD_ST_FREQ); s.st_wtptr = D(BD_ST_PTR);
        s.agc = D(BD_AGC_VAL); s.mse = D(BD_MSE); s.ebno = D(BD_EB_EBNO); s.vol_gain = D(BD_VOL_GAIN); s.rotator_freq = D(BD_ROT_FREQ);
        s.n_sig_true = I(BI_SIG_TRUE); s.n_sig_false = I(BI_SIG_FALSE); s.cntr = I(BI_CNTR); s.startstop = I(BI_STARTSTOP);
        s.last_burst_ebno = D(BD_LAST_EBNO_EMIT); s.n_ebno_emits = I(BI_EBNO_EMITS);
    }
    return JAERO_OK;
}

} // extern "C"

// ====================================================================== R/T burst channel layer (§8(f)2)
#include "rtchannel.cuh"

struct jaero_rt {
    int device; cudaStream_t stream;
    RtParams rp;
    std::vector<void *> allocs;
    int16_t *d_soft_stage; int *d_count_stage; size_t stage_cap;
    RtState *h_state; uint8_t *h_out;
    long long launches;
    int vector_mode;
};

extern "C" {

int jaero_rt_create(double fb, int n_channels, int device, jaero_rt **out)
{
    if (!out || n_channels <= 0) { set_error("jaero_rt_create: bad argument"); return JAERO_E_ARG; }
    const int ifb = (int)(fb >= 0.0 ? fb + 0.5 : fb - 0.5);
    if (ifb != 600 && ifb != 1200 && ifb != 10500) { set_error("jaero_rt_create: burst R/T channels run at 600, 1200 or 10500 bps"); return JAERO_E_ARG; }
    int ndev = 0;
    JB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("jaero_rt_create: no such CUDA device"); return JAERO_E_CUDA; }
    JB_CUDA(cudaSetDevice(device));
    jaero_rt *r = new (std::nothrow) jaero_rt();
    if (!r) { set_error("out of host memory"); return JAERO_E_ARG; }
    CreateGuard<jaero_rt> guard(r, jaero_rt_destroy);
    r->device = device; r->d_soft_stage = 0; r->d_count_stage = 0; r->stage_cap = 0; r->h_state = 0; r->h_out = 0; r->launches = 0;
    JB_CUDA(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
    RtParams &rp = r->rp;
    memset(&rp, 0, sizeof rp);
    rp.n_channels = n_channels; rp.ifb = ifb; rp.oqpsk = (ifb == 10500);
    rp.number_of_bits = (ifb == 10500) ? 4992 : 1152;                       // aerol.cpp:1012-1050
    rp.total_number_of_bits = rp.oqpsk ? ifb : ifb * 3;                     // :1062-1070
    const size_t C = n_channels;
    int rc = 0;
    auto alloc = [&](auto **ptr, size_t count) { int q = dev_alloc_zero(ptr, count, r->stream); if (!q) r->allocs.push_back((void *)*ptr); return q; };
    rc |= alloc(&rp.state, C); rc |= alloc(&rp.slots, C * RT_SLOTS); rc |= alloc(&rp.blocks, C * RT_SLOTS * RT_BLOCK); rc |= alloc(&rp.out, C * RT_OUT * RT_OUT_BYTES);
    if (rc) return JAERO_E_CUDA;
    {   // AeroLScrambler::pre_state (aerol.h:397-437)
        std::vector<uint8_t> seq(5000);
        int st[15] = {1, 1, 0, 1, 0, 0, 1, 0, 1, 0, 1, 1, 0, 0, 1};
        for (int a = 0; a < 5000; a++) { const int v = st[0] ^ st[14]; seq[a] = (uint8_t)v; for (int i = 14; i > 0; i--) st[i] = st[i - 1]; st[0] = v; }
        if (rt_set_scrambler(seq.data())) return JAERO_E_CUDA;
    }
    if (rt_init(rp, r->stream)) return JAERO_E_CUDA;
    JB_CUDA(cudaStreamSynchronize(r->stream));
    JB_CUDA(cudaMallocHost(&r->h_state, C * sizeof(RtState)));
    JB_CUDA(cudaMallocHost(&r->h_out, C * RT_OUT * RT_OUT_BYTES));
    guard.release();
    *out = r;
    return JAERO_OK;
}
void jaero_rt_destroy(jaero_rt *r)
{
    if (!r) return;
    cudaSetDevice(r->device);
    cudaStreamSynchronize(r->stream);
    for (void *q : r->allocs) cudaFree(q);
    cudaFree(r->d_soft_stage); cudaFree(r->d_count_stage);
    cudaFreeHost(r->h_state); cudaFreeHost(r->h_out);
    cudaStreamDestroy(r->stream);
    delete r;
}
int64_t jaero_rt_launch_count(const jaero_rt *r) { return r ? r->launches : 0; }

int jaero_rt_process_softbits(jaero_rt *r, const int16_t *soft, size_t cap, const int32_t *counts)
{
    if (!r || !soft || !counts || cap == 0) { set_error("jaero_rt_process_softbits: bad argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(r->device));
    const size_t C = r->rp.n_channels;
    if (C * cap > r->stage_cap) {
        JB_CUDA(cudaStreamSynchronize(r->stream));
        cudaFree(r->d_soft_stage); cudaFree(r->d_count_stage); r->d_soft_stage = 0; r->d_count_stage = 0;
        JB_CUDA(cudaMalloc(&r->d_soft_stage, C * cap * sizeof(int16_t)));
        JB_CUDA(cudaMalloc(&r->d_count_stage, C * sizeof(int)));
        r->stage_cap = C * cap;
    }
    JB_CUDA(cudaMemcpyAsync(r->d_soft_stage, soft, C * cap * sizeof(int16_t), cudaMemcpyHostToDevice, r->stream));
    JB_CUDA(cudaMemcpyAsync(r->d_count_stage, counts, C * sizeof(int), cudaMemcpyHostToDevice, r->stream));
    if (rt_process(r->rp, r->d_soft_stage, r->d_count_stage, cap, r->stream, &r->launches, r->vector_mode ? -1 : 0)) return JAERO_E_CUDA;
    JB_CUDA(cudaStreamSynchronize(r->stream));
    return JAERO_OK;
}
int jaero_rt_process_burst(jaero_rt *r, jaero_burst *b)
{
    if (!r || !b || r->rp.n_channels != b->p.n_channels || r->device != b->device) { set_error("jaero_rt_process_burst: demodulator mismatch"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(r->device));
    const BurstParams &bp = b->p;
    JB_CUDA(cudaStreamSynchronize(r->stream));
    // on the demodulator's stream: ordered after its kernels; its soft ring is drained afterwards
    if (rt_process(r->rp, bp.soft, bp.BI + (size_t)BI_SOFT_COUNT * bp.cpad, (size_t)bp.soft_cap, b->stream, &r->launches,
                   r->vector_mode ? (bp.kind == 1 ? 32 : 12) : 0)) return JAERO_E_CUDA;   // emit sizes: burstoqpskdemodulator.cpp / burstmskdemodulator.cpp:735
    burst_soft_reset_kernel<<<(bp.n_channels + 127) / 128, 128, 0, b->stream>>>(bp);
    JB_CUDA(cudaGetLastError());
    r->launches++;
    JB_CUDA(cudaStreamSynchronize(b->stream));
    return JAERO_OK;
}
// Opt-in: reproduce AeroL::Decode's return in the middle of a soft-bit vector when the burst time-out fires (aerol.cpp:2018-2027)
int jaero_rt_set_vector_mode(jaero_rt *r, int enabled)
{
    if (!r) { set_error("null handle"); return JAERO_E_ARG; }
    r->vector_mode = enabled ? 1 : 0;
    return JAERO_OK;
}
int jaero_rt_tick(jaero_rt *r)
{
    if (!r) { set_error("null handle"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(r->device));
    if (rt_tick(r->rp, r->stream)) return JAERO_E_CUDA;
    r->launches++;
    return JAERO_OK;
}
int jaero_rt_read_packets(jaero_rt *r, uint8_t *out, int cap_packets, int32_t *counts)
{
    if (!r || !out || !counts || cap_packets <= 0) { set_error("jaero_rt_read_packets: bad argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(r->device));
    const size_t C = r->rp.n_channels;
    JB_CUDA(cudaMemcpyAsync(r->h_state, r->rp.state, C * sizeof(RtState), cudaMemcpyDeviceToHost, r->stream));
    JB_CUDA(cudaMemcpyAsync(r->h_out, r->rp.out, C * RT_OUT * RT_OUT_BYTES, cudaMemcpyDeviceToHost, r->stream));
    JB_CUDA(cudaStreamSynchronize(r->stream));
    bool overflow = false;
    for (size_t ch = 0; ch < C; ch++) {
        const int n = r->h_state[ch].out_count;
        overflow |= r->h_state[ch].overflow != 0 || n > cap_packets;
        counts[ch] = n < cap_packets ? n : cap_packets;
        for (int k = 0; k < counts[ch]; k++)
            memcpy(out + (ch * cap_packets + k) * RT_OUT_BYTES, r->h_out + (ch * RT_OUT + k) * RT_OUT_BYTES, RT_OUT_BYTES);
    }
    if (rt_out_reset(r->rp, r->stream)) return JAERO_E_CUDA;
    r->launches++;
    if (overflow) { set_error("R/T packet queue overflow: read more often"); return JAERO_E_OVERFLOW; }
    return JAERO_OK;
}
int jaero_rt_get_stats(jaero_rt *r, int32_t *n_trials, int32_t *n_bad, int32_t *dcd)
{
    if (!r) { set_error("null handle"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(r->device));
    const size_t C = r->rp.n_channels;
    JB_CUDA(cudaMemcpyAsync(r->h_state, r->rp.state, C * sizeof(RtState), cudaMemcpyDeviceToHost, r->stream));
    JB_CUDA(cudaStreamSynchronize(r->stream));
    for (size_t ch = 0; ch < C; ch++) {
        if (n_trials) n_trials[ch] = r->h_state[ch].n_trials;
        if (n_bad) n_bad[ch] = r->h_state[ch].n_bad;
        if (dcd) dcd[ch] = r->h_state[ch].datacd;
    }
    return JAERO_OK;
}

} // extern "C"

// ====================================================================== C-channel (8400 bps) frame layer (§8(f)3)
#include "cchannel.cuh"

struct jaero_cchannel {
    int device; cudaStream_t stream; cudaEvent_t ev_batch, ev_own; bool own_dirty;   // stream ordering as in jaero_pchannel
    CChanParams cp;
    std::vector<void *> allocs;
    uint8_t *vit_overlap; int *vit_overlap_len, *vit_renorm, *vit_valid;
    int16_t *d_soft_stage; int *d_count_stage; size_t stage_cap;
    CChanState *h_state; uint8_t *h_out;
    long long launches;
};

extern "C" {

int jaero_cchannel_create(int n_channels, int device, jaero_cchannel **out)
{
    if (!out || n_channels <= 0) { set_error("jaero_cchannel_create: bad argument"); return JAERO_E_ARG; }
    int ndev = 0;
    JB_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("jaero_cchannel_create: no such CUDA device"); return JAERO_E_CUDA; }
    JB_CUDA(cudaSetDevice(device));
    jaero_cchannel *c = new (std::nothrow) jaero_cchannel();
    if (!c) { set_error("out of host memory"); return JAERO_E_ARG; }
    CreateGuard<jaero_cchannel> guard(c, jaero_cchannel_destroy);
    c->device = device; c->d_soft_stage = 0; c->d_count_stage = 0; c->stage_cap = 0; c->h_state = 0; c->h_out = 0; c->launches = 0;
    JB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    JB_CUDA(cudaEventCreateWithFlags(&c->ev_batch, cudaEventDisableTiming));
    JB_CUDA(cudaEventCreateWithFlags(&c->ev_own, cudaEventDisableTiming));
    c->own_dirty = false;
    CChanParams &cp = c->cp;
    memset(&cp, 0, sizeof cp);
    cp.n_channels = n_channels; cp.dl2_len = 2714 - 6 + 1;                     // dl2.setLength(2714-6) (aerol.cpp:1037)
    const size_t C = n_channels;
    int rc = 0;
    auto alloc = [&](auto **ptr, size_t count) { int q = dev_alloc_zero(ptr, count, c->stream); if (!q) c->allocs.push_back((void *)*ptr); return q; };
    rc |= alloc(&cp.state, C); rc |= alloc(&cp.coded, C * CC_QUEUE * CC_CODED_PITCH); rc |= alloc(&cp.decoded, C * CC_QUEUE * CC_DEC);
    rc |= alloc(&cp.ready, C); rc |= alloc(&cp.dl2, C * cp.dl2_len); rc |= alloc(&cp.out, C * CC_OUT * CC_RECORD);
    rc |= alloc(&c->vit_overlap, C * 64); rc |= alloc(&c->vit_overlap_len, C); rc |= alloc(&c->vit_renorm, C); rc |= alloc(&c->vit_valid, C);
    if (rc) return JAERO_E_CUDA;
    {
        std::vector<uint8_t> seq(5000);
        int st[15] = {1, 1, 0, 1, 0, 0, 1, 0, 1, 0, 1, 1, 0, 0, 1};
        for (int a = 0; a < 5000; a++) { const int v = st[0] ^ st[14]; seq[a] = (uint8_t)v; for (int i = 14; i > 0; i--) st[i] = st[i - 1]; st[0] = v; }
        if (cchan_set_scrambler(seq.data())) return JAERO_E_CUDA;
    }
    if (cchan_init(cp, c->stream)) return JAERO_E_CUDA;
    JB_CUDA(cudaStreamSynchronize(c->stream));
    JB_CUDA(cudaMallocHost(&c->h_state, C * sizeof(CChanState)));
    JB_CUDA(cudaMallocHost(&c->h_out, C * CC_OUT * CC_RECORD));
    guard.release();
    *out = c;
    return JAERO_OK;
}
void jaero_cchannel_destroy(jaero_cchannel *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (void *q : c->allocs) cudaFree(q);
    cudaFree(c->d_soft_stage); cudaFree(c->d_count_stage);
    cudaFreeHost(c->h_state); cudaFreeHost(c->h_out);
    if (c->ev_batch) cudaEventDestroy(c->ev_batch);
    if (c->ev_own) cudaEventDestroy(c->ev_own);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}
int64_t jaero_cchannel_launch_count(const jaero_cchannel *c) { return c ? c->launches : 0; }

int jaero_cchannel_process_batch(jaero_cchannel *c, jaero_batch *b)
{
    if (!c || !b || c->cp.n_channels != b->p.n_channels || c->device != b->device) { set_error("jaero_cchannel_process_batch: batch mismatch"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(c->device));
    const DemodParams &dp = b->p;
    int *dcd = dp.I + (size_t)I_DCD * dp.cpad;
    if (pc_enter(c, b->stream)) return JAERO_E_CUDA;
    if (cchan_process(c->cp, dp.soft, dp.I + (size_t)I_SOFT_COUNT * dp.cpad, (size_t)dp.soft_cap, dcd, c->vit_overlap, c->vit_overlap_len,
                      c->vit_renorm, c->vit_valid, b->stream, &c->launches, dp.I + (size_t)I_LOST_N * dp.cpad, dp.lost_pos, dp.cpad)) return JAERO_E_CUDA;
    soft_reset_kernel<<<(dp.n_channels + 127) / 128, 128, 0, b->stream>>>(dp);
    JB_CUDA(cudaGetLastError());
    c->launches++;
    return pc_leave(c, b->stream) ? JAERO_E_CUDA : JAERO_OK;
}
int jaero_cchannel_process_softbits(jaero_cchannel *c, const int16_t *soft, size_t cap, const int32_t *counts)
{
    if (!c || !soft || !counts || cap == 0) { set_error("jaero_cchannel_process_softbits: bad argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(c->device));
    const size_t C = c->cp.n_channels;
    if (C * cap > c->stage_cap) {
        JB_CUDA(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_soft_stage); cudaFree(c->d_count_stage); c->d_soft_stage = 0; c->d_count_stage = 0;
        JB_CUDA(cudaMalloc(&c->d_soft_stage, C * cap * sizeof(int16_t)));
        JB_CUDA(cudaMalloc(&c->d_count_stage, C * sizeof(int)));
        c->stage_cap = C * cap;
    }
    c->own_dirty = true;
    JB_CUDA(cudaMemcpyAsync(c->d_soft_stage, soft, C * cap * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream));
    JB_CUDA(cudaMemcpyAsync(c->d_count_stage, counts, C * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (cchan_process(c->cp, c->d_soft_stage, c->d_count_stage, cap, nullptr, c->vit_overlap, c->vit_overlap_len, c->vit_renorm, c->vit_valid,
                      c->stream, &c->launches)) return JAERO_E_CUDA;
    JB_CUDA(cudaStreamSynchronize(c->stream));
    return JAERO_OK;
}
int jaero_cchannel_tick(jaero_cchannel *c, jaero_batch *b)
{
    if (!c || (b && b->p.n_channels != c->cp.n_channels)) { set_error("jaero_cchannel_tick: bad argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(c->device));
    if (b) { if (pc_enter(c, b->stream)) return JAERO_E_CUDA; } else c->own_dirty = true;
    if (cchan_tick(c->cp, b ? b->p.I + (size_t)I_DCD * b->p.cpad : nullptr, b ? b->stream : c->stream)) return JAERO_E_CUDA;
    c->launches++;
    return (b && pc_leave(c, b->stream)) ? JAERO_E_CUDA : JAERO_OK;
}
int jaero_cchannel_lost_signal(jaero_cchannel *c, jaero_batch *b, int channel)     // AeroL::LostSignal, see jaero_pchannel_lost_signal
{
    if (!c || channel >= c->cp.n_channels || (b && b->p.n_channels != c->cp.n_channels)) { set_error("jaero_cchannel_lost_signal: bad argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(c->device));
    if (b) { if (pc_enter(c, b->stream)) return JAERO_E_CUDA; } else c->own_dirty = true;
    if (cchan_lost(c->cp, channel, b ? b->p.I + (size_t)I_DCD * b->p.cpad : nullptr, b ? b->stream : c->stream)) return JAERO_E_CUDA;
    c->launches++;
    return (b && pc_leave(c, b->stream)) ? JAERO_E_CUDA : JAERO_OK;
}
int jaero_cchannel_write_batch(jaero_cchannel *c, jaero_batch *b, const int16_t *pcm, size_t n, size_t stride)   // see jaero_pchannel_write_batch
{
    if (!c || !b || !pcm || c->cp.n_channels != b->p.n_channels || c->device != b->device) { set_error("jaero_cchannel_write_batch: bad argument"); return JAERO_E_ARG; }
    if (stride < n) { set_error("jaero_cchannel_write_batch: channel_stride < n_samples"); return JAERO_E_ARG; }
    b->p.wire_sigstat = 1;
    size_t done = 0;
    while (done < n) {
        const size_t k = piece_before_next_trigger(b, n - done);
        int rc = jaero_batch_write(b, pcm + done, k, stride);
        if (rc) return rc;
        rc = jaero_cchannel_process_batch(c, b);
        if (rc) return rc;
        done += k;
    }
    return JAERO_OK;
}
int jaero_cchannel_read_frames(jaero_cchannel *c, uint8_t *out, int cap_frames, int32_t *counts)
{
    if (!c || !out || !counts || cap_frames <= 0) { set_error("jaero_cchannel_read_frames: bad argument"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(c->device));
    const size_t C = c->cp.n_channels;
    JB_CUDA(cudaMemcpyAsync(c->h_state, c->cp.state, C * sizeof(CChanState), cudaMemcpyDeviceToHost, c->stream));
    JB_CUDA(cudaMemcpyAsync(c->h_out, c->cp.out, C * CC_OUT * CC_RECORD, cudaMemcpyDeviceToHost, c->stream));
    JB_CUDA(cudaStreamSynchronize(c->stream));
    bool overflow = false;
    for (size_t ch = 0; ch < C; ch++) {
        const int n = c->h_state[ch].out_count;
        overflow |= c->h_state[ch].overflow != 0 || n > cap_frames;
        counts[ch] = n < cap_frames ? n : cap_frames;
        for (int k = 0; k < counts[ch]; k++) memcpy(out + (ch * cap_frames + k) * CC_RECORD, c->h_out + (ch * CC_OUT + k) * CC_RECORD, CC_RECORD);
    }
    if (cchan_out_reset(c->cp, c->stream)) return JAERO_E_CUDA;
    c->own_dirty = true;
    c->launches++;
    if (overflow) { set_error("C-channel frame queue overflow: read more often"); return JAERO_E_OVERFLOW; }
    return JAERO_OK;
}
int jaero_cchannel_get_stats(jaero_cchannel *c, int32_t *dcd, int64_t *su_total, int64_t *su_ok)
{
    if (!c) { set_error("null handle"); return JAERO_E_ARG; }
    JB_CUDA(cudaSetDevice(c->device));
    const size_t C = c->cp.n_channels;
    JB_CUDA(cudaMemcpyAsync(c->h_state, c->cp.state, C * sizeof(CChanState), cudaMemcpyDeviceToHost, c->stream));
    JB_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t ch = 0; ch < C; ch++) {
        if (dcd) dcd[ch] = c->h_state[ch].datacd;
        if (su_total) su_total[ch] = c->h_state[ch].su_total;
        if (su_ok) su_ok[ch] = c->h_state[ch].su_ok;
    }
    return JAERO_OK;
}

} // extern "C"

// ====================================================================== ingest router (§8(f)4, host side)
// The many-channel feed of the reference: one ZMQ PUB topic per channel, every message three frames
// [topic][uint32 sample rate][int16 PCM] (JAERO/zmq_audioreceiver.cpp:37-87, subscription = the first 5 bytes of the
// topic, :46), delivered to the demodulator's dataReceived(audio, sampleRate) slot (oqpskdemodulator.cpp:686-693).
// This router takes the three frames as the transport hands them over (no libzmq dependency), files the PCM under the
// channel whose topic matches and, once every channel has n samples, feeds them to a batch in one jaero_batch_write.
struct jaero_ingest {
    int n_channels; uint32_t rate; size_t cap;
    std::vector<std::string> topics;
    std::vector<int16_t> pcm;               // [n_channels][cap]
    std::vector<size_t> fill;
    long long dropped_bytes, messages;
};

extern "C" {

int jaero_ingest_create(int n_channels, const char *const *topics, uint32_t sample_rate, size_t capacity_samples, jaero_ingest **out)
{
    if (!out || !topics || n_channels <= 0 || capacity_samples == 0) { set_error("jaero_ingest_create: bad argument"); return JAERO_E_ARG; }
    jaero_ingest *g = new (std::nothrow) jaero_ingest();
    if (!g) { set_error("out of host memory"); return JAERO_E_ARG; }
    g->n_channels = n_channels; g->rate = sample_rate; g->cap = capacity_samples; g->dropped_bytes = 0; g->messages = 0;
    for (int c = 0; c < n_channels; c++) {
        if (!topics[c]) { delete g; set_error("jaero_ingest_create: null topic"); return JAERO_E_ARG; }
        g->topics.push_back(std::string(topics[c]).substr(0, 5));          // zmq_setsockopt(..., ZMQ_SUBSCRIBE, topic, 5)
    }
    g->pcm.assign((size_t)n_channels * capacity_samples, 0);
    g->fill.assign(n_channels, 0);
    *out = g;
    return JAERO_OK;
}
void jaero_ingest_destroy(jaero_ingest *g) { delete g; }

int jaero_ingest_message(jaero_ingest *g, const void *topic, size_t topic_len, const void *rate, size_t rate_len, const void *pcm, size_t pcm_bytes)
{
    if (!g || !topic || !rate || (!pcm && pcm_bytes)) { set_error("jaero_ingest_message: null argument"); return JAERO_E_ARG; }
    if (rate_len != 4) { set_error("jaero_ingest_message: the sample-rate frame must be 4 bytes"); return JAERO_E_ARG; }
    uint32_t r; memcpy(&r, rate, 4);                                           // memcpy(&sampleRate, rate, 4) (:70)
    int ch = -1;
    for (int c = 0; c < g->n_channels && ch < 0; c++) {
        const std::string &t = g->topics[c];
        if (topic_len >= t.size() && memcmp(topic, t.data(), t.size()) == 0) ch = c;   // prefix match, as a ZMQ subscription does
    }
    if (ch < 0) { set_error("jaero_ingest_message: no channel subscribes to this topic"); return JAERO_E_ARG; }
    if (r != g->rate) { set_error("jaero_ingest_message: sample rate differs from the batch's (the reference re-applies its settings; a batch is fixed-rate)"); return JAERO_E_STATE; }
    g->messages++;
    size_t n = pcm_bytes / 2;                                                   // writeData: len/2 int16 samples
    const size_t room = g->cap - g->fill[ch];
    // a full channel buffer refuses the whole message (nothing is filed, so the caller can flush and re-send): dropping
    // the tail silently would desynchronise this channel against the lock-step batch
    if (n > room) { g->dropped_bytes += (long long)n * 2; set_error("jaero_ingest_message: channel buffer full (flush the batch, then re-send this message)"); return JAERO_E_OVERFLOW; }
    memcpy(g->pcm.data() + (size_t)ch * g->cap + g->fill[ch], pcm, n * 2);
    g->fill[ch] += n;
    return ch;
}
size_t jaero_ingest_available(const jaero_ingest *g)
{
    if (!g) return 0;
    size_t m = g->cap;
    for (int c = 0; c < g->n_channels; c++) m = std::min(m, g->fill[c]);
    return m;
}
int jaero_ingest_flush(jaero_ingest *g, jaero_batch *b, size_t n)
{
    if (!g || !b || b->p.n_channels != g->n_channels) { set_error("jaero_ingest_flush: batch mismatch"); return JAERO_E_ARG; }
    if (n == 0) return JAERO_OK;
    if (n > jaero_ingest_available(g)) { set_error("jaero_ingest_flush: not every channel has that many samples"); return JAERO_E_STATE; }
    const int rc = jaero_batch_write(b, g->pcm.data(), n, g->cap);
    if (rc != JAERO_OK) return rc;
    JB_CUDA(cudaStreamSynchronize(b->stream));                                 // the pageable staging rows are reused below
    for (int c = 0; c < g->n_channels; c++) {
        int16_t *row = g->pcm.data() + (size_t)c * g->cap;
        memmove(row, row + n, (g->fill[c] - n) * 2);
        g->fill[c] -= n;
    }
    return JAERO_OK;
}

} // extern "C"

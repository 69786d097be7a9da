// K6 — 8400 bps pre-filter: mix down -> 2049-tap RRC via streaming FFT convolution (nfft 4096) -> mix up.
//
// Replaces the `fb==8400` block at the top of OqpskDemodulator::writeData (JAERO/oqpskdemodulator.cpp:343-381) and the
// JFastFir object it drives (un-vendored jontio/JFFT; observable contract pinned by JAERO/tests/jfastfir_tests.cpp:
// out[n] = sum_k h[k] x[n-L-k] with L = nfft-K+1 = 2048, zeros for n < 2L). The same streaming FFT-FIR engine serves the
// burst demodulators' Hilbert filter (DSP.cpp:754-794).
//
//   pre_down_kernel        thread/channel: x[i] = mixer_fir_pre.CIS * pcm[i]        (:353-365, phases accumulated serially)
//   fir_exchange_up_kernel thread/channel: per-sample in/out exchange against the L-sample staging block (JFastFir::update)
//                                          followed by the conjugate mix-up from the saved phase (:371-379)
//   fir_block_kernel       CTA/channel:    overlap-save block: [history | new L] -> FFT4096 -> xH -> IFFT4096 -> last L
// The 4096-point transforms run entirely in shared memory (64 KB) as four Stockham radix-8 passes with the butterflies
// in registers (512 threads, one radix-8 butterfly each per pass).
#include "demod_device.cuh"
#include "prefilter.cuh"

namespace jb {

static const int FB_THREADS = 512;

__device__ __forceinline__ double2 pc_add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 pc_sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 pc_mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
template <bool INV> __device__ __forceinline__ double2 pc_rot(double2 a) { return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x); }
template <bool INV> __device__ __forceinline__ void p_dft4(double2 &a0, double2 &a1, double2 &a2, double2 &a3)
{
    const double2 t0 = pc_add(a0, a2), t1 = pc_sub(a0, a2), t2 = pc_add(a1, a3), t3 = pc_rot<INV>(pc_sub(a1, a3));
    a0 = pc_add(t0, t2); a1 = pc_add(t1, t3); a2 = pc_sub(t0, t2); a3 = pc_sub(t1, t3);
}
template <bool INV> __device__ __forceinline__ void p_dft8(double2 *v)
{
    double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    p_dft4<INV>(e0, e1, e2, e3);
    p_dft4<INV>(o0, o1, o2, o3);
    const double h = 0.70710678118654752440;
    const double2 w1 = INV ? make_double2(h, h) : make_double2(h, -h);
    const double2 w3 = INV ? make_double2(-h, h) : make_double2(-h, -h);
    o1 = pc_mul(o1, w1); o2 = pc_rot<INV>(o2); o3 = pc_mul(o3, w3);
    v[0] = pc_add(e0, o0); v[4] = pc_sub(e0, o0);
    v[1] = pc_add(e1, o1); v[5] = pc_sub(e1, o1);
    v[2] = pc_add(e2, o2); v[6] = pc_sub(e2, o2);
    v[3] = pc_add(e3, o3); v[7] = pc_sub(e3, o3);
}
// one Stockham radix-8 pass over a single 4096-point sequence in shared memory; 512 threads = 512 butterflies
template <bool INV> __device__ __forceinline__ void pass8_4096(double2 *s, int Ns, const double2 *__restrict__ tw)
{
    const int j = threadIdx.x, nb = 512;
    double2 v[8];
#pragma unroll
    for (int t = 0; t < 8; t++) v[t] = s[j + t * nb];
    __syncthreads();
    const int k = j % Ns;
    const int wmul = 4096 / (Ns * 8);
#pragma unroll
    for (int t = 1; t < 8; t++) {
        double2 w = tw[t * k * wmul];
        if (INV) w.y = -w.y;
        v[t] = pc_mul(v[t], w);
    }
    p_dft8<INV>(v);
    const int ob = (j / Ns) * Ns * 8 + k;
#pragma unroll
    for (int t = 0; t < 8; t++) s[ob + t * Ns] = v[t];
    __syncthreads();
}
template <bool INV> __device__ __forceinline__ void fft4096(double2 *s, const double2 *__restrict__ tw)
{
    pass8_4096<INV>(s, 1, tw); pass8_4096<INV>(s, 8, tw); pass8_4096<INV>(s, 64, tw); pass8_4096<INV>(s, 512, tw);
}

// overlap-save block (the compute step of JFastFir::update when the staging block fills): one CTA per channel
__global__ void __launch_bounds__(FB_THREADS)
fir_block_kernel(FirStream f, int first_block)
{
    extern __shared__ double2 fs[];
    const int ch = blockIdx.x;
    const int L = FIR_L;
    double2 *hist = f.hist + (size_t)ch * L, *inb = f.inblk + (size_t)ch * L, *outb = f.outblk + (size_t)ch * L;
    for (int j = threadIdx.x; j < L; j += FB_THREADS) {
        const double2 h = hist[j], x = inb[j];
        fs[j] = h; fs[L + j] = x;
        hist[j] = x;                      // new history = last K-1 = L samples of [history | block]
    }
    __syncthreads();
    fft4096<false>(fs, f.tw);
    for (int j = threadIdx.x; j < 2 * L; j += FB_THREADS) fs[j] = pc_mul(fs[j], f.H[j]);
    __syncthreads();
    fft4096<true>(fs, f.tw);
    const double sc = 1.0 / 4096.0;       // JFFT::ifft is 1/N-normalised
    for (int j = threadIdx.x; j < L; j += FB_THREADS) {
        const double2 y = fs[L + j];      // positions K-1 .. K-1+L-1 are the valid (non-circular) outputs
        outb[j] = first_block ? make_double2(0.0, 0.0) : make_double2(y.x * sc, y.y * sc);
    }
}

// mix down with mixer_fir_pre (oqpskdemodulator.cpp:353-365). The oscillator itself is not advanced here: the reference
// rewinds it to the saved phase before the mix-up loop, whose end state is what persists.
__global__ void pre_down_kernel(PreParams q, const int16_t *__restrict__ pcm, size_t stride, int n)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= q.n_channels) return;
    Osc o = {q.osc[0 * q.cpad + ch], q.osc[1 * q.cpad + ch], q.osc[2 * q.cpad + ch], 0.0};
    const int16_t *row = pcm + (size_t)ch * stride;
    double2 *x = q.x + (size_t)ch * q.xstride;
    for (int i = 0; i < n; i++) {
        const double dval = ((double)row[i]) / 32768.0;
        const int t = osc_index(o.ptr);
        x[i] = make_double2(q.cos_t[t] * dval, q.sin_t[t] * dval);
        osc_next_frame(o);
    }
    // savedphase=GetPhaseDeg() (:354) ... SetPhaseDeg(savedphase) (:371): the round trip through degrees is kept
    const double saved = (360.0 * q.osc[0 * q.cpad + ch] / ((double)WTSIZE));
    Osc r = {0, 0, 0, 0};
    osc_set_phase_deg(r, saved);
    q.osc[3 * q.cpad + ch] = r.ptr;       // start pointer of the mix-up loop
}

// JFastFir::update sample exchange for samples [i0,i1) of this call + conjugate mix-up (:371-379)
__global__ void fir_exchange_up_kernel(PreParams q, FirStream f, int i0, int i1, int fill0)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= q.n_channels) return;
    double2 *x = q.x + (size_t)ch * q.xstride;
    double2 *inb = f.inblk + (size_t)ch * FIR_L;
    const double2 *outb = f.outblk + (size_t)ch * FIR_L;
    Osc o = {q.osc[3 * q.cpad + ch], q.osc[1 * q.cpad + ch], q.osc[2 * q.cpad + ch], 0.0};
    int fill = fill0;
    for (int i = i0; i < i1; i++) {
        const double2 xin = x[i];
        const double2 y = outb[fill];
        inb[fill] = xin;
        fill++;
        const int t = osc_index(o.ptr);
        x[i] = pc_mul(y, make_double2(q.cos_t[t], -q.sin_t[t]));          // *= WTCISValue_conj()
        osc_next_frame(o);
    }
    q.osc[3 * q.cpad + ch] = o.ptr;
}

// end of writeData (:608): mixer_fir_pre.SetFreq(mixer2_freq_sum/i); its pointer is where the mix-up loop left it
__global__ void pre_finish_kernel(PreParams q, const double *__restrict__ m2_freq_sum, int n, double Fs)
{
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= q.n_channels) return;
    double f = m2_freq_sum[ch] / ((double)n);
    if (f < 0) f = 0;
    q.osc[2 * q.cpad + ch] = f;
    q.osc[1 * q.cpad + ch] = (f) * ((double)WTSIZE) / Fs;
    q.osc[0 * q.cpad + ch] = q.osc[3 * q.cpad + ch];
}

int fir_block_launch(const FirStream &f, int n_channels, int first_block, cudaStream_t s)
{
    JB_CUDA(cudaFuncSetAttribute(fir_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 16));
    fir_block_kernel<<<n_channels, FB_THREADS, 4096 * 16, s>>>(f, first_block);
    JB_CUDA(cudaGetLastError());
    return 0;
}
int pre_down_launch(const PreParams &q, const int16_t *pcm, size_t stride, int n, cudaStream_t s)
{
    pre_down_kernel<<<(q.n_channels + 63) / 64, 64, 0, s>>>(q, pcm, stride, n);
    JB_CUDA(cudaGetLastError());
    return 0;
}
int fir_exchange_up_launch(const PreParams &q, const FirStream &f, int i0, int i1, int fill0, cudaStream_t s)
{
    fir_exchange_up_kernel<<<(q.n_channels + 63) / 64, 64, 0, s>>>(q, f, i0, i1, fill0);
    JB_CUDA(cudaGetLastError());
    return 0;
}
int pre_finish_launch(const PreParams &q, const double *m2_freq_sum, int n, double Fs, cudaStream_t s)
{
    pre_finish_kernel<<<(q.n_channels + 63) / 64, 64, 0, s>>>(q, m2_freq_sum, n, Fs);
    JB_CUDA(cudaGetLastError());
    return 0;
}

} // namespace jb

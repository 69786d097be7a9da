// TEST INFRASTRUCTURE ONLY (oracle).
// Restatement of the un-vendored third-party dependency jontio/JFFT (no pinned
// version: the reference's CI clones HEAD, ci-linux-build.sh:152-162;
// JAERO/JAERO.pro:24,98 compiles ../../JFFT/jfft.cpp). The source is NOT under
// /root/reference, so this file restates its *observable contract* as pinned by
// the reference's own call sites and tests:
//   JFFT::fft / ifft            JAERO/fftwrapper.cpp:19-35 + tests/fftwrapper_tests.cpp:23-54
//   JFFT::fft_real / ifft_real  JAERO/fftrwrapper.cpp:19-39 + tests/fftrwrapper_tests.cpp:24-55
//   JFastFir::SetKernel/update  JAERO/oqpskdemodulator.cpp:283,368, DSP.cpp:788
//                               + tests/jfastfir_tests.cpp:31-58 (10 000-sample golden)
// forward = e^{-j2*pi*kn/N} unnormalised; inverse = 1/N-normalised (the wrapper
// multiplies by N again). Sizes used by the reference are powers of two only.
#ifndef JFFT_RESTATED_H
#define JFFT_RESTATED_H
#include <complex>
#include <vector>
#include "qt_shim.h"

class JFFT
{
public:
    typedef std::complex<double> cpx_type;
    JFFT() : nfft(0) {}
    void init(int nfft);
    void fft(QVector<cpx_type> &x) { run(x.data(), x.size(), false); }
    void ifft(QVector<cpx_type> &x) { run(x.data(), x.size(), true); }
    void fft(std::vector<cpx_type> &x) { run(x.data(), (int)x.size(), false); }
    void ifft(std::vector<cpx_type> &x) { run(x.data(), (int)x.size(), true); }
    void fft_real(const QVector<double> &in, QVector<cpx_type> &out);
    void ifft_real(const QVector<cpx_type> &in, QVector<double> &out);
    void run(cpx_type *x, int n, bool inverse);
private:
    int nfft;
    std::vector<cpx_type> tw;   // e^{-j2*pi*k/nfft}, k<nfft/2
    std::vector<int> rev;
};

// Streaming FFT (overlap-save) FIR. Observable behaviour pinned by
// tests/jfastfir_tests.cpp for a 2049-tap kernel and nfft=4096:
//   out[n] = sum_k h[k]*x[n-L-k] for n >= 2L, 0 before, with L = nfft-K+1 (=2048).
// The block latency L is what a per-sample in/out exchange against an L-sample
// staging block yields; for the Hilbert use (DSP.cpp:788, default nfft) the
// latency is unobservable downstream.
class JFastFir
{
public:
    typedef std::complex<double> cpx_type;
    JFastFir() : nfft(0), K(0), L(0), fill(0), nblocks(0) {}
    void SetKernel(const QVector<double> &k);
    void SetKernel(const QVector<double> &k, int nfft);
    void SetKernel(const QVector<cpx_type> &k);
    void SetKernel(const QVector<cpx_type> &k, int nfft);
    void update(QVector<cpx_type> &inout);
    void update(cpx_type *inout, int n);
    int latency() const { return L; }
private:
    JFFT fft;
    int nfft, K, L, fill;
    long long nblocks;
    std::vector<cpx_type> H;        // FFT of zero-padded kernel
    std::vector<cpx_type> hist;     // last K-1 input samples
    std::vector<cpx_type> inblk;    // staging block (L new samples)
    std::vector<cpx_type> outblk;   // previous block's result (L samples)
    std::vector<cpx_type> work;
};
#endif

// TEST INFRASTRUCTURE ONLY (oracle). See oracle/shim/jfft.h for the contract
// this restates (jontio/JFFT is an un-vendored, un-pinned dependency of the
// reference: JAERO/JAERO.pro:24,98).
#include "jfft.h"
#include <cmath>

void JFFT::init(int n)
{
    nfft = n;
    tw.resize(n / 2 > 0 ? n / 2 : 1);
    for (int k = 0; k < n / 2; k++) {
        double a = -2.0 * M_PI * (double)k / (double)n;
        tw[k] = cpx_type(cos(a), sin(a));
    }
    rev.resize(n);
    int bits = 0;
    while ((1 << bits) < n) bits++;
    for (int i = 0; i < n; i++) {
        int r = 0;
        for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
        rev[i] = r;
    }
}

// iterative radix-2 decimation-in-time, double precision
void JFFT::run(cpx_type *x, int n, bool inverse)
{
    if (n != nfft) init(n);
    for (int i = 0; i < n; i++) if (rev[i] > i) std::swap(x[i], x[rev[i]]);
    for (int len = 2; len <= n; len <<= 1) {
        int half = len >> 1, step = n / len;
        for (int i = 0; i < n; i += len) {
            for (int k = 0; k < half; k++) {
                cpx_type w = tw[k * step];
                if (inverse) w = std::conj(w);
                cpx_type a = x[i + k], b = x[i + k + half] * w;
                x[i + k] = a + b;
                x[i + k + half] = a - b;
            }
        }
    }
    if (inverse) {
        double s = 1.0 / (double)n;
        for (int i = 0; i < n; i++) x[i] *= s;
    }
}

void JFFT::fft_real(const QVector<double> &in, QVector<cpx_type> &out)
{
    out.resize(in.size());
    for (int i = 0; i < in.size(); i++) out[i] = cpx_type(in[i], 0.0);
    run(out.data(), out.size(), false);
}

void JFFT::ifft_real(const QVector<cpx_type> &in, QVector<double> &out)
{
    // inverse of a half-spectrum: bins above N/2 are rebuilt by Hermitian symmetry
    int n = in.size();
    std::vector<cpx_type> t(n);
    for (int i = 0; i <= n / 2; i++) t[i] = in[i];
    for (int i = n / 2 + 1; i < n; i++) t[i] = std::conj(in[n - i]);
    run(t.data(), n, true);
    out.resize(n);
    for (int i = 0; i < n; i++) out[i] = t[i].real();
}

static int jfastfir_default_nfft(int K)
{
    int n = 1;
    while (n < 4 * K) n <<= 1;     // the "x4 rule of thumb" (oqpskdemodulator.cpp:283 comment)
    return n;
}
void JFastFir::SetKernel(const QVector<double> &k) { SetKernel(k, jfastfir_default_nfft(k.size())); }
void JFastFir::SetKernel(const QVector<cpx_type> &k) { SetKernel(k, jfastfir_default_nfft(k.size())); }
void JFastFir::SetKernel(const QVector<double> &k, int n)
{
    QVector<cpx_type> c(k.size());
    for (int i = 0; i < k.size(); i++) c[i] = cpx_type(k[i], 0.0);
    SetKernel(c, n);
}
void JFastFir::SetKernel(const QVector<cpx_type> &k, int n)
{
    K = k.size();
    nfft = n;
    L = nfft - K + 1;
    assert(L > 0);
    fft.init(nfft);
    H.assign(nfft, cpx_type(0, 0));
    for (int i = 0; i < K; i++) H[i] = k[i];
    fft.run(H.data(), nfft, false);
    hist.assign(K - 1, cpx_type(0, 0));
    inblk.assign(L, cpx_type(0, 0));
    outblk.assign(L, cpx_type(0, 0));
    work.assign(nfft, cpx_type(0, 0));
    fill = 0;
    nblocks = 0;
}
void JFastFir::update(QVector<cpx_type> &inout) { update(inout.data(), inout.size()); }
void JFastFir::update(cpx_type *inout, int n)
{
    for (int i = 0; i < n; i++) {
        cpx_type xin = inout[i];
        inout[i] = outblk[fill];          // result of the previous block (zeros at first)
        inblk[fill] = xin;
        fill++;
        if (fill == L) {
            fill = 0;
            // overlap-save: [K-1 history | L new] -> last L outputs are valid
            for (int j = 0; j < K - 1; j++) work[j] = hist[j];
            for (int j = 0; j < L; j++) work[K - 1 + j] = inblk[j];
            // new history = last K-1 samples of the concatenation
            for (int j = 0; j < K - 1; j++) hist[j] = work[L + j];
            fft.run(work.data(), nfft, false);
            for (int j = 0; j < nfft; j++) work[j] *= H[j];
            fft.run(work.data(), nfft, true);
            // the very first block is discarded: observable output is zero for n < 2L
            // (tests/jfastfir_tests.cpp only checks n >= 4096)
            if (nblocks == 0) for (int j = 0; j < L; j++) outblk[j] = cpx_type(0, 0);
            else for (int j = 0; j < L; j++) outblk[j] = work[K - 1 + j];
            nblocks++;
        }
    }
}

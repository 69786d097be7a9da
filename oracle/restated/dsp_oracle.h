// TEST INFRASTRUCTURE ONLY (oracle): CPU restatement of the reference's DSP primitives
// (JAERO/DSP.h, JAERO/DSP.cpp). Never linked into the product. Every function cites the
// reference lines it follows; arithmetic keeps the reference's operation order in double.
#ifndef JAERO_DSP_ORACLE_H
#define JAERO_DSP_ORACLE_H
#include <cmath>
#include <complex>
#include <vector>
#include <cstdint>

namespace jor {
typedef std::complex<double> cpx;
static const int WTSIZE = 19999;                        // DSP.h:21

// DSP.cpp:9-30 (only the sin/cos tables are used on the hot path)
struct Trig { std::vector<double> SinWT, CosWT; Trig(); };
const Trig &trig();

// WaveTable: DSP.h:40-81, DSP.cpp:32-51,70-93,142-180,222-238
struct WaveTable
{
    double WTptr, WTstep, freq, samplerate, last_WTptr, FractionOfSampleItPassesBy;
    WaveTable() { last_WTptr = 0; samplerate = 48000; freq = 1000; WTstep = (1000.0) * WTSIZE / (48000); WTptr = 0; FractionOfSampleItPassesBy = 0.0; }
    void SetFreq(double f, int sr)                      // DSP.cpp:142-149
    { freq = f; samplerate = sr; if (freq < 0) freq = 0; WTstep = (freq) * ((double)WTSIZE) / ((float)sr); while (((int)WTptr) >= WTSIZE) WTptr -= WTSIZE; }
    void SetFreq(double f)                              // DSP.cpp:151-156
    { freq = f; if (freq < 0) freq = 0; WTstep = (freq) * ((double)WTSIZE) / samplerate; }
    double GetFreqHz() const { return freq; }
    void IncreseFreqHz(double f) { f += freq; SetFreq(f); }                 // DSP.cpp:163-167
    void SetPhaseDeg(double p)                          // DSP.cpp:175-180
    { p = std::fmod(p, 360.0); while (p < 0) p += 360.0; WTptr = (p / 360.0) * ((double)WTSIZE); }
    void IncresePhaseDeg(double p) { p += (360.0 * WTptr / ((double)WTSIZE)); SetPhaseDeg(p); }   // DSP.cpp:169-173
    double GetPhaseDeg() const { return (360.0 * WTptr / ((double)WTSIZE)); }
    void AdvanceFractionOfWave(double x)                // DSP.h:56
    { WTptr += x * WTSIZE; while (WTptr >= WTSIZE) WTptr -= WTSIZE; while (WTptr < 0) WTptr += WTSIZE; }
    void WTnextFrame()                                  // DSP.cpp:70-77
    { if (WTstep < 0) WTstep = 0; last_WTptr = WTptr; WTptr += WTstep; while (((int)WTptr) >= WTSIZE) WTptr -= WTSIZE; }
    int index() const { int t = (int)WTptr; if (t >= WTSIZE) t = 0; if (t < 0) t = WTSIZE - 1; return t; }   // DSP.cpp:81-83
    cpx WTCISValue() const { int t = index(); return cpx(trig().CosWT[t], trig().SinWT[t]); }            // DSP.cpp:79-85
    cpx WTCISValue_conj() const { return std::conj(WTCISValue()); }
    bool IfHavePassedPoint(double FractionOfWave)       // DSP.cpp:222-238
    {
        double t_last = last_WTptr, t = WTptr, pt = (FractionOfWave * WTSIZE);
        t_last -= pt; t -= pt;
        if (t_last < 0.0) t_last += WTSIZE;
        if (t < 0.0) t += WTSIZE;
        if ((t_last > 3.0 * WTSIZE / 4.0) && (t < 1.0 * WTSIZE / 4.0)) { FractionOfSampleItPassesBy = t / WTstep; return true; }
        return false;
    }
};

// FIR: DSP.cpp:271-304 (N+1 ring; the output excludes the sample just written)
struct FIR
{
    std::vector<double> points, buff; int ptr;
    void init(const std::vector<double> &p) { points = p; buff.assign(p.size() + 1, 0.0); ptr = 0; }
    double FIRUpdateAndProcess(double sig)
    {
        int buffsize = (int)buff.size(), N = (int)points.size();
        buff[ptr] = sig; ptr++; if (ptr >= buffsize) ptr = 0;
        int tptr = ptr; double outsum = 0;
        for (int i = 0; i < N; i++) { outsum += points[i] * buff[tptr]; tptr++; if (tptr >= buffsize) tptr = 0; }
        return outsum;
    }
};

// MovingAverage: DSP.cpp:388-426
struct MovingAverage
{
    std::vector<double> buf; double sum, Val; int ptr;
    void init(int n) { buf.assign(n, 0.0); sum = 0; Val = 0; ptr = 0; }
    void Zero() { std::fill(buf.begin(), buf.end(), 0.0); ptr = 0; Val = 0; sum = 0; }
    double Update(double s) { return UpdateSigned(std::fabs(s)); }
    double UpdateSigned(double s)
    { sum = sum - buf[ptr]; sum = sum + s; buf[ptr] = s; ptr++; ptr %= (int)buf.size(); Val = sum / ((double)buf.size()); return Val; }
};

// AGC: DSP.cpp:357-379
struct AGC
{
    MovingAverage ma; double AGCVal;
    void init(double seconds, double Fs) { ma.init((int)std::round(seconds * Fs)); AGCVal = 0; }
    double Update(double sig)
    {
        ma.sum = ma.sum - ma.buf[ma.ptr]; ma.sum = ma.sum + std::fabs(sig); ma.buf[ma.ptr] = std::fabs(sig);
        ma.ptr++; ma.ptr %= (int)ma.buf.size();
        AGCVal = 1.414213562 / std::fmax(ma.sum / ((double)ma.buf.size()), 0.000001);
        AGCVal = std::fmax(AGCVal, 0.000001);
        return AGCVal;
    }
};

// MSEcalc: DSP.cpp:434-463
struct MSEcalc
{
    MovingAverage pointmean, msema; double mse;
    void init(int n) { pointmean.init(n); msema.init(n); mse = 0; }
    double Update(cpx pt)
    {
        pointmean.Update(std::abs(pt));
        double mu = pointmean.Val; if (mu < 0.000001) mu = 0.000001;
        cpx t = std::sqrt(2) * pt / mu;
        double tda = (std::fabs(t.real()) - 1.0), tdb = (std::fabs(t.imag()) - 1.0);
        mse = msema.Update((tda * tda) + (tdb * tdb));
        return mse;
    }
};

// OQPSKEbNoMeasure: DSP.cpp:715-744 ; MSKEbNoMeasure: DSP.cpp:487-505 (EbNo starts at 0: zero-filled object)
struct EbNoMeasure
{
    MovingAverage E, E2; double EbNo, Var, Mean, Fs, fb; bool oqpsk;
    void init(int n, bool oq, double Fs_, double fb_) { E.init(n); E2.init(n); EbNo = Var = Mean = 0; Fs = Fs_; fb = fb_; oqpsk = oq; }
    double Update(double sig)
    {
        E2.Update(sig * sig); Mean = E.Update(sig);
        double tebno;
        if (oqpsk) {
            double MeanSquared = Mean * Mean;
            Var = (E2.Val) - (E.Val * E.Val);
            Var -= (0.024709 * MeanSquared);
            double mvr = (((Fs * MeanSquared / (2.0 * fb * Var))) * 0.13743);
            if (mvr < 0.000000001) mvr = 0.000000001;
            tebno = 10.0 * std::log10(mvr);
            if (std::isnan(tebno)) tebno = 50;
            if (tebno > 50.0) tebno = 50;
            if (tebno < 0.0) tebno = 0;
        } else {
            Var = (E2.Val) - (E.Val * E.Val);
            double alpha = std::sqrt(2) / Mean;
            tebno = 10.0 * (std::log10(2.0) - std::log10(((Var * alpha * alpha) - 0.0085))) - 5.0;
            if (std::isnan(tebno)) tebno = 50;
            if (tebno > 50.0) tebno = 50;
        }
        EbNo = EbNo * 0.8 + 0.2 * tebno;
        return EbNo;
    }
};

// Delay<T> (fractional, linear interpolation): DSP.h:341-379
template <class T> struct Delay
{
    std::vector<T> buff; int buffptr; double fractdelay;
    void setdelay(double fd) { fractdelay = fd; buff.assign((int)std::ceil(fd) + 1, T(0)); buffptr = 0; }
    T update(T sig)
    {
        buff[buffptr] = sig;
        double dptr = ((double)buffptr) - fractdelay;
        buffptr++; buffptr %= (int)buff.size();
        while (std::floor(dptr) < 0) dptr += ((double)buff.size());
        int iptr = (int)std::floor(dptr);
        double weighting = dptr - ((double)iptr);
        T older = buff[iptr]; iptr++; iptr %= (int)buff.size();
        T newer = buff[iptr];
        return (weighting * newer + (1.0 - weighting) * older);
    }
};

// DelayThing<T> (integer): DSP.h:439-486
template <class T> struct DelayThing
{
    std::vector<T> buffer; int ptr;
    void setLength(int length) { buffer.assign(length + 1, T(0)); ptr = 0; }
    void update(T &data) { buffer[ptr] = data; ptr++; ptr %= (int)buffer.size(); data = buffer[ptr]; }
    T update_dont_touch(T data) { buffer[ptr] = data; ptr++; ptr %= (int)buffer.size(); return buffer[ptr]; }
};

// IIR (direct form, ring buffers): DSP.cpp:634-709
struct IIR
{
    double a[3], b[3], bx[3], by[2]; int xp, yp; double y;
    void init() { bx[0] = bx[1] = bx[2] = 0; by[0] = by[1] = 0; xp = yp = 0; y = 0; }
    double update(double sig)
    {
        bx[xp] = sig; xp++; xp %= 3;
        y = 0;
        for (int i = 2; i >= 0; i--) { y += bx[xp] * b[i]; xp++; xp %= 3; }
        for (int i = 2; i >= 1; i--) { y -= by[yp] * a[i]; yp++; yp %= 2; }
        y /= a[0];
        by[yp] = y; yp++; yp %= 2;
        return y;
    }
};

// DiffDecode::UpdateSoft: DSP.cpp:531-563
struct DiffDecode
{
    double lastsoftstate; DiffDecode() : lastsoftstate(-1) {}
    double UpdateSoft(double soft)
    {
        double r;
        if (soft < 0 && lastsoftstate < 0) { r = lastsoftstate; lastsoftstate = soft; }
        else if (soft > 0 && lastsoftstate > 0) { r = -lastsoftstate; lastsoftstate = soft; }
        else { r = std::fabs(lastsoftstate); lastsoftstate = soft; }
        return r;
    }
};

// RootRaisedCosine::design: DSP.h:316-338
std::vector<double> rrc_design(double alpha, int firsize, double samplerate, double symbol_freq);
// qRound (Qt5 qglobal.h), used at oqpskdemodulator.cpp:569,575 / mskdemodulator.cpp:453,463
inline int qRound(double d) { return d >= 0.0 ? int(d + 0.5) : int(d - double(int(d - 1)) + 0.5) + int(d - 1); }

// unnormalised forward / inverse radix-2 FFT, the oracle's stand-in for JFFT (see oracle/shim/jfft.h)
void fft_pow2(cpx *x, int n, bool inverse_unnormalised);
} // namespace jor
#endif

// Host-side C++ mirror of the reference's demodulator classes over the C ABI (include/jaero_b200.h).
//
// Same class names, Settings structs (fields and defaults), methods and signal names as
//   OqpskDemodulator       JAERO/oqpskdemodulator.h:15-70        MskDemodulator       JAERO/mskdemodulator.h:18-167
//   BurstOqpskDemodulator  JAERO/burstoqpskdemodulator.h:18-90   BurstMskDemodulator  JAERO/burstmskdemodulator.h:21-100
//   JConvolutionalCodec    JAERO/jconvolutionalcodec.h:17-50
// so that code written against the reference (AeroL, the reference's own tests) reads the same. Qt is not required:
// signals are std::function members with the signal's name and argument list (QVector<short> -> std::vector<short>,
// QString -> std::string); inside JAERO the QIODevice subclass of INTEGRATION.md section 2 forwards to these. Each object is
// a batch of one channel on one GPU; a many-channel front end uses the C ABI (or jaero_b200::Batch below) directly.
// There is no CPU path: if the CUDA library cannot create the object, setSettings() raises WarningTextSignal with
// jaero_last_error() and writeData() consumes its input without producing anything, as a reference object without
// settings would.
#ifndef JAERO_B200_HOST_HPP
#define JAERO_B200_HOST_HPP
#include <cstdint>
#include <cstring>
#include <functional>
#include <string>
#include <complex>
#include <vector>

#include "jaero_b200.h"

namespace jaero_b200 {

typedef long long qint64;

namespace detail {

// what differs between the four classes
struct ContinuousApi {
    typedef jaero_batch handle;
    typedef jaero_status status;
    static int write(handle *h, const int16_t *p, size_t n) { return jaero_batch_write(h, p, n, n); }
    static int read(handle *h, int16_t *o, size_t cap, int32_t *c) { return jaero_batch_read_softbits(h, o, cap, c); }
    static int get_status(handle *h, status *s) { return jaero_batch_get_status(h, 0, s); }
    static int set_dcd(handle *h, bool d) { return jaero_batch_set_dcd(h, 0, d ? 1 : 0); }
    static int set_afc(handle *h, bool s) { return jaero_batch_set_afc(h, s ? 1 : 0); }
    static int set_sql(handle *h, bool s) { return jaero_batch_set_sql(h, s ? 1 : 0); }
    static int set_cpu_reduce(handle *h, bool s) { return jaero_batch_set_cpu_reduce(h, s ? 1 : 0); }
    static void destroy(handle *h) { jaero_batch_destroy(h); }
    static bool telemetry(const status &s, double &peak, std::vector<std::complex<double> > &pts)
    { peak = s.peak_volume; pts.clear(); pts.push_back(std::complex<double>(s.scatter[2], s.scatter[3])); pts.push_back(std::complex<double>(s.scatter[0], s.scatter[1])); return true; }
};
struct BurstApi {
    typedef jaero_burst handle;
    typedef jaero_burst_status status;
    static int write(handle *h, const int16_t *p, size_t n) { return jaero_burst_write(h, p, n, n); }
    static int read(handle *h, int16_t *o, size_t cap, int32_t *c) { return jaero_burst_read_softbits(h, o, cap, c); }
    static int get_status(handle *h, status *s) { return jaero_burst_get_status_all(h, s); }   // one channel: one record
    static int set_dcd(handle *h, bool d) { return jaero_burst_set_dcd(h, 0, d ? 1 : 0); }
    static int set_afc(handle *h, bool s) { return jaero_burst_set_afc(h, s ? 1 : 0); }
    static int set_sql(handle *h, bool s) { return jaero_burst_set_sql(h, s ? 1 : 0); }
    static int set_cpu_reduce(handle *, bool) { return JAERO_OK; }   // the burst classes never read cpuReduce on this path
    static void destroy(handle *h) { jaero_burst_destroy(h); }
    static bool telemetry(const status &, double &, std::vector<std::complex<double> > &) { return false; }
};

template <class Api> class DemodulatorBase
{
public:
    // ---- signals (same names and argument lists as the reference's) ----
    std::function<void(const std::vector<short> &soft_bits)> processDemodulatedSoftBits;
    std::function<void(double mse)> MSESignal;
    std::function<void(bool gotasignal)> SignalStatus;
    std::function<void(double EbNo)> EbNoMeasurmentSignal;
    std::function<void(double freq_est, double freq_center, double bandwidth)> Plottables;
    std::function<void(const std::string &str)> WarningTextSignal;
    std::function<void(double Fs)> SampleRateChanged;
    std::function<void(double fb, bool burstmode)> BitRateChanged;
    // GUI telemetry of the reference (oqpskdemodulator.cpp:403-404,549; mskdemodulator.cpp:343-344,443). The reference paces these
    // with a wall-clock timer (150 ms); here they fire once per writeData. PeakVolume = max |sample| since the previous emit,
    // ScatterPoints = the most recent constellation points (decimated pointbuff), OrgOverlapedBuffer = the last 2^13 input samples.
    std::function<void(double maxval)> PeakVolume;
    std::function<void(const std::vector<std::complex<double> > &points)> ScatterPoints;
    std::function<void(const std::vector<double> &buffer)> OrgOverlapedBuffer;

    void setAFC(bool state) { afc = state; if (h) check(Api::set_afc(h, state)); }
    void setSQL(bool state) { sql = state; if (h) check(Api::set_sql(h, state)); }
    void setCPUReduce(bool state) { cpuReduce = state; if (h) check(Api::set_cpu_reduce(h, state)); }
    void invalidatesettings() { Fs = -1; fb = -1; }
    void start() {}                                   // QIODevice::open in the reference
    void stop() {}
    qint64 readData(char *, qint64) { return 0; }

    // int16 mono PCM, len bytes; everything is processed before the call returns, soft bits are emitted in the reference's
    // group size (32 values OQPSK / 12 MSK). Always returns len, like the reference.
    qint64 writeData(const char *data, qint64 len)
    {
        if (!h || len <= 0 || !data) return len;
        const size_t n = (size_t)len / sizeof(int16_t);
        const int16_t *pcm = reinterpret_cast<const int16_t *>(data);
        std::vector<int16_t> aligned;
        if (reinterpret_cast<uintptr_t>(data) % sizeof(int16_t)) { aligned.resize(n); memcpy(aligned.data(), data, n * sizeof(int16_t)); pcm = aligned.data(); }
        if (OrgOverlapedBuffer) {                      // spectrumcycbuff (oqpskdemodulator.cpp:395-396): the last 8192 input samples
            if (spectrum.size() != 8192) { spectrum.assign(8192, 0.0); spectrum_ptr = 0; }
            for (size_t k = n > 8192 ? n - 8192 : 0; k < n; k++) { spectrum[spectrum_ptr] = ((double)pcm[k]) / 32768.0; spectrum_ptr = (spectrum_ptr + 1) % 8192; }
        }
        for (size_t at = 0; at < n;) {                // at most one second per call into the library: the soft-bit ring holds two
            const size_t take = (n - at < max_chunk) ? (n - at) : max_chunk;
            if (!check(Api::write(h, pcm + at, take))) return len;
            at += take;
            int32_t count = 0;
            if (!check(Api::read(h, drain.data(), drain.size(), &count))) return len;
            for (int32_t k = 0; k < count; k++) {
                pending.push_back(drain[(size_t)k]);
                if ((int)pending.size() == emit_group) { if (processDemodulatedSoftBits) processDemodulatedSoftBits(pending); pending.clear(); }
            }
        }
        report();
        return len;
    }
    void dataReceived(const char *audio, size_t bytes, uint32_t /*sampleRate*/) { writeData(audio, (qint64)bytes); }   // ZMQ feed

    double getCurrentFreq()
    {
        typename Api::status st;
        return (h && Api::get_status(h, &st) == JAERO_OK) ? st.center_freq : 0.0;
    }
    void DCDstatSlot(bool _dcd) { dcd = _dcd; if (h) check(Api::set_dcd(h, _dcd)); }
    bool ok() const { return h != 0; }
    typename Api::handle *handle() const { return h; }   // for the device-resident consumers (jaero_pchannel_process_batch, ...)

protected:
    explicit DemodulatorBase(int device_ordinal) : h(0), device(device_ordinal), afc(false), sql(false), cpuReduce(false), dcd(false),
        Fs(-1), fb(-1), lockingbw(0), signalthreshold(0), last_status(-1), emit_group(32), max_chunk(48000) {}
    ~DemodulatorBase() { if (h) Api::destroy(h); }
    DemodulatorBase(const DemodulatorBase &);            // not copyable
    DemodulatorBase &operator=(const DemodulatorBase &);

    bool check(int rc)
    {
        if (rc == JAERO_OK) return true;
        if (WarningTextSignal) WarningTextSignal(std::string(jaero_last_error()));
        return false;
    }
    void fill(jaero_settings &s, int kind, int fft_power, double freq_center)
    {
        memset(&s, 0, sizeof s);
        s.kind = kind; s.coarsefreqest_fft_power = fft_power; s.freq_center = freq_center; s.lockingbw = lockingbw; s.fb = fb; s.Fs = Fs;
        s.signalthreshold = signalthreshold; s.afc = afc; s.sql = sql; s.cpu_reduce = cpuReduce; s.report_ebno = 1;
    }
    // the part of setSettings every class shares: change signals, new device object, dcd carried over
    template <class Create> void apply(double newFs, double newfb, bool burstmode, int group, Create create)
    {
        if (newFs != Fs && SampleRateChanged) SampleRateChanged(newFs);
        if (newfb != fb && BitRateChanged) BitRateChanged(newfb, burstmode);
        Fs = newFs; fb = newfb;
        if (h) { Api::destroy(h); h = 0; }
        pending.clear(); last_status = -1;
        emit_group = group;
        max_chunk = (size_t)(Fs > 0 ? Fs : 48000);
        drain.assign((size_t)(2 * fb + 64) > 4096 ? (size_t)(2 * fb + 64) : 4096, 0);   // the library's ring capacity
        typename Api::handle *nh = 0;
        if (!check(create(&nh))) return;
        h = nh;
        if (dcd) Api::set_dcd(h, true);
        report();
    }
    void report()
    {
        typename Api::status st;
        if (!h || Api::get_status(h, &st) != JAERO_OK) return;
        if (Plottables) Plottables(st.mixer2_freq, st.center_freq, lockingbw);
        if (EbNoMeasurmentSignal) EbNoMeasurmentSignal(st.ebno);
        if (MSESignal) MSESignal(st.mse);
        const int now = st.mse <= signalthreshold ? 1 : 0;      // the reference raises SignalStatus when the gate changes
        if (now != last_status) { last_status = now; if (SignalStatus) SignalStatus(now != 0); }
        double peak = 0; std::vector<std::complex<double> > pts;
        if (Api::telemetry(st, peak, pts)) { if (PeakVolume) PeakVolume(peak); if (ScatterPoints) ScatterPoints(pts); }
        if (OrgOverlapedBuffer && spectrum.size() == 8192) OrgOverlapedBuffer(spectrum);
    }

    typename Api::handle *h;
    int device;
    bool afc, sql, cpuReduce, dcd;
    double Fs, fb, lockingbw, signalthreshold;
    int last_status, emit_group;
    size_t max_chunk;
    std::vector<short> pending;
    std::vector<int16_t> drain;
    std::vector<double> spectrum; size_t spectrum_ptr = 0;
};

}  // namespace detail

class OqpskDemodulator : public detail::DemodulatorBase<detail::ContinuousApi>
{
public:
    struct Settings {
        int coarsefreqest_fft_power; double freq_center, lockingbw, fb, Fs, signalthreshold; bool zmqAudio;
        Settings() : coarsefreqest_fft_power(14), freq_center(8000), lockingbw(10500), fb(10500), Fs(48000), signalthreshold(0.65), zmqAudio(false) {}
    };
    explicit OqpskDemodulator(int device_ordinal = 0) : DemodulatorBase(device_ordinal) {}
    void setSettings(Settings s)                        // JAERO/oqpskdemodulator.cpp:175-333
    {
        lockingbw = s.lockingbw; signalthreshold = s.signalthreshold;
        apply(s.Fs, s.fb, false, 32, [&](jaero_batch **out) {
            jaero_settings js; fill(js, JAERO_KIND_OQPSK, s.coarsefreqest_fft_power, s.freq_center);
            return jaero_batch_create(&js, 1, 0, device, out);
        });
    }
    void CenterFreqChangedSlot(double freq_center) { if (h) check(jaero_batch_set_center_freq(h, 0, freq_center)); }
};

class MskDemodulator : public detail::DemodulatorBase<detail::ContinuousApi>
{
public:
    struct Settings {
        int coarsefreqest_fft_power; double freq_center, lockingbw, fb, Fs; int symbolspercycle; double signalthreshold; bool zmqAudio;
        Settings() : coarsefreqest_fft_power(13), freq_center(1000), lockingbw(900), fb(600), Fs(48000), symbolspercycle(16), signalthreshold(0.5), zmqAudio(false) {}
    };
    explicit MskDemodulator(int device_ordinal = 0) : DemodulatorBase(device_ordinal) {}
    void setSettings(Settings s)                        // JAERO/mskdemodulator.cpp:133-311; 12-value emits (:472)
    {
        lockingbw = s.lockingbw; signalthreshold = s.signalthreshold;
        apply(s.Fs, s.fb, false, 12, [&](jaero_batch **out) {
            jaero_settings js; fill(js, JAERO_KIND_MSK, s.coarsefreqest_fft_power, s.freq_center);
            return jaero_batch_create(&js, 1, 0, device, out);
        });
    }
    void CenterFreqChangedSlot(double freq_center) { if (h) check(jaero_batch_set_center_freq(h, 0, freq_center)); }
};

class BurstOqpskDemodulator : public detail::DemodulatorBase<detail::BurstApi>
{
public:
    struct Settings {
        int coarsefreqest_fft_power; double freq_center, lockingbw, fb, Fs, signalthreshold; bool channel_stereo, zmqAudio;
        Settings() : coarsefreqest_fft_power(13), freq_center(8000), lockingbw(10500), fb(10500), Fs(48000), signalthreshold(0.6), channel_stereo(false), zmqAudio(false) {}
    };
    explicit BurstOqpskDemodulator(int device_ordinal = 0) : DemodulatorBase(device_ordinal) { afc = true; }   // ctor: afc=true
    void setSettings(Settings s)                        // JAERO/burstoqpskdemodulator.cpp:203-277
    {
        lockingbw = s.lockingbw; signalthreshold = s.signalthreshold;
        apply(s.Fs, s.fb, true, 32, [&](jaero_burst **out) {
            jaero_settings js; fill(js, JAERO_KIND_OQPSK, s.coarsefreqest_fft_power, s.freq_center);
            int rc = jaero_burst_oqpsk_create(&js, 1, device, out);
            if (rc == JAERO_OK && !afc) jaero_burst_set_afc(*out, 0);
            return rc;
        });
    }
};

class BurstMskDemodulator : public detail::DemodulatorBase<detail::BurstApi>
{
public:
    struct Settings {
        int coarsefreqest_fft_power; double freq_center, lockingbw, fb, Fs; int symbolspercycle; double signalthreshold; bool zmqAudio;
        Settings() : coarsefreqest_fft_power(13), freq_center(1000), lockingbw(500), fb(125), Fs(8000), symbolspercycle(16), signalthreshold(0.6), zmqAudio(false) {}
    };
    explicit BurstMskDemodulator(int device_ordinal = 0) : DemodulatorBase(device_ordinal) { afc = true; }
    void setSettings(Settings s)                        // JAERO/burstmskdemodulator.cpp:154-323; 12-value emits
    {
        lockingbw = s.lockingbw; signalthreshold = s.signalthreshold;
        apply(s.Fs, s.fb, true, 12, [&](jaero_burst **out) {
            jaero_settings js; fill(js, JAERO_KIND_MSK, s.coarsefreqest_fft_power, s.freq_center);
            int rc = jaero_burst_msk_create(&js, 1, device, out);
            if (rc == JAERO_OK && !afc) jaero_burst_set_afc(*out, 0);
            return rc;
        });
    }
};

// JConvolutionalCodec (K=7, rate 1/2, polynomials 109 / 79 only: the code JAERO uses, JAERO/aerol.cpp:940)
class JConvolutionalCodec
{
public:
    explicit JConvolutionalCodec(int device_ordinal = 0) : v(0), device(device_ordinal), paddinglength(0) {}
    ~JConvolutionalCodec() { if (v) jaero_viterbi_destroy(v); }
    std::function<void(const std::string &str)> WarningTextSignal;
    // JAERO/jconvolutionalcodec.cpp:27-49; false when the code is not the one this library implements or no device exists
    bool SetCode(int inv_rate, int order, const std::vector<uint16_t> &poly, int _paddinglength = 24 * 4)
    {
        if (v) { jaero_viterbi_destroy(v); v = 0; }
        if (inv_rate != 2 || order != 7 || poly.size() != 2 || poly[0] != 109 || poly[1] != 79) {
            if (WarningTextSignal) WarningTextSignal("JConvolutionalCodec: only the K=7 rate-1/2 code (109, 79) is implemented");
            return false;
        }
        paddinglength = _paddinglength;
        if (jaero_viterbi_create(1, paddinglength, device, &v) != JAERO_OK) { v = 0; if (WarningTextSignal) WarningTextSignal(jaero_last_error()); return false; }
        return true;
    }
    int getPaddinglength() const { return paddinglength; }
    // soft values 0..255 (128 = erasure), one per byte; returns the decoded bits (JAERO/jconvolutionalcodec.cpp:151-200)
    std::vector<int> &Decode_Continuous(const std::vector<uint8_t> &soft_bits_in)
    {
        decoded_bits.clear();
        if (!v || soft_bits_in.empty()) return decoded_bits;
        std::vector<uint8_t> bits(soft_bits_in.size() / 2);
        int32_t n_valid = 0;
        if (jaero_viterbi_decode_continuous(v, soft_bits_in.data(), soft_bits_in.size(), 0, bits.data(), &n_valid) != JAERO_OK) {
            if (WarningTextSignal) WarningTextSignal(jaero_last_error());
            return decoded_bits;
        }
        decoded_bits.assign(bits.begin(), bits.begin() + n_valid);
        return decoded_bits;
    }
    // one self-contained block (Decode_soft, JAERO/jconvolutionalcodec.cpp:96-125)
    std::vector<int> &Decode_soft(const std::vector<uint8_t> &soft_bits_in, int size)
    {
        decoded_bits.clear();
        if (!v || size <= 0 || (size_t)size > soft_bits_in.size()) return decoded_bits;
        std::vector<uint8_t> bits((size_t)size / 2);
        if (jaero_viterbi_decode_block(v, soft_bits_in.data(), (size_t)size, bits.data()) != JAERO_OK) {
            if (WarningTextSignal) WarningTextSignal(jaero_last_error());
            return decoded_bits;
        }
        decoded_bits.assign(bits.begin(), bits.end());
        return decoded_bits;
    }
private:
    JConvolutionalCodec(const JConvolutionalCodec &);
    JConvolutionalCodec &operator=(const JConvolutionalCodec &);
    jaero_viterbi *v;
    int device, paddinglength;
    std::vector<int> decoded_bits;
};

}  // namespace jaero_b200
#endif

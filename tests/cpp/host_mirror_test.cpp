// Host-mirror plumbing test (runs on CPU). Built twice by tests/test_host_mirror.py:
//   -DMOCK  linked against tests/cpp/mock_capi.cpp  -> checks chunking, emit grouping, setters, failure path
//   (none)  linked against the real libjaero_b200.so -> on a machine without a GPU: setSettings must fail loudly
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "jaero_b200_host.hpp"

#define CHECK(c) do { if (!(c)) { printf("FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)
#ifdef MOCK
extern "C" { void mock_set_fail(int); int *mock_get_log(void); }
#endif

int main()
{
    using namespace jaero_b200;
#ifdef MOCK
    int *log = mock_get_log();
    {   // ---- continuous OQPSK: odd-sized writes, 32-value emits, nothing lost, nothing reordered
        OqpskDemodulator d;
        std::vector<short> got; int emits = 0, bad_group = 0, status_emits = 0; bool last_status = false; double fs_changed = 0, fb_changed = 0;
        std::vector<std::string> warnings;
        d.processDemodulatedSoftBits = [&](const std::vector<short> &v) { emits++; if (v.size() != 32) bad_group++; got.insert(got.end(), v.begin(), v.end()); };
        d.SignalStatus = [&](bool s) { status_emits++; last_status = s; };
        d.SampleRateChanged = [&](double f) { fs_changed = f; };
        d.BitRateChanged = [&](double f, bool burst) { fb_changed = f; CHECK(!burst); return 0; };
        d.WarningTextSignal = [&](const std::string &s) { warnings.push_back(s); };
        CHECK(d.writeData("abcd", 4) == 4 && emits == 0);                   // no settings yet: consumed, nothing produced
        d.setSQL(false); d.setAFC(true);
        OqpskDemodulator::Settings st; st.freq_center = 5760;
        d.setSettings(st);
        d.setCPUReduce(true);                                               // JAERO's order (mainwindow.cpp:281-283)
        CHECK(d.ok() && warnings.empty() && fs_changed == 48000 && fb_changed == 10500 && log[0] == 1 && log[6] == 1);
        CHECK(d.getCurrentFreq() == 5760);
        std::vector<int16_t> pcm(200003);
        for (size_t i = 0; i < pcm.size(); i++) pcm[i] = (int16_t)(i * 7 + 3);
        const size_t cuts[] = {1, 4799, 4800, 1, 100000, 33, 90369};       // sums to 200003; one write is longer than a second
        size_t at = 0;
        for (size_t k = 0; k < sizeof cuts / sizeof cuts[0]; k++) { CHECK(d.writeData((const char *)(pcm.data() + at), (qint64)(cuts[k] * 2)) == (qint64)(cuts[k] * 2)); at += cuts[k]; }
        CHECK(at == pcm.size());
        CHECK(log[3] <= 48000);                                             // never more than one second per library call
        const size_t produced = (pcm.size() + 3) / 4;
        CHECK(bad_group == 0 && got.size() == produced / 32 * 32);
        for (size_t k = 0; k < got.size(); k++) CHECK(got[k] == (short)(pcm[4 * k] & 0xff));
        CHECK(status_emits == 2 && last_status);                            // false after setSettings, true once the mock "locks"
        d.DCDstatSlot(true); CHECK(log[7] == 1);
        d.CenterFreqChangedSlot(6000); CHECK(d.getCurrentFreq() == 6000);
        d.setSettings(st);                                                  // new settings: new device object, DCD carried over
        CHECK(log[0] == 2 && log[1] == 1 && log[7] == 2);
    }
    CHECK(log[1] == 2);                                                     // destructor released the handle
    {   // ---- MSK: 12-value emits
        MskDemodulator d; int bad = 0; size_t n = 0;
        d.processDemodulatedSoftBits = [&](const std::vector<short> &v) { if (v.size() != 12) bad++; n += v.size(); };
        d.setSettings(MskDemodulator::Settings());
        std::vector<int16_t> pcm(48000 * 3 + 17, 5);
        d.writeData((const char *)pcm.data(), (qint64)(pcm.size() * 2));
        CHECK(bad == 0 && n == ((pcm.size() + 79) / 80) / 12 * 12);                // the mock emits one value per (int)(Fs/fb) = 80 samples
    }
    {   // ---- burst classes: the -1 marker travels inside the stream; AFC starts on and setAFC(false) reaches the library
        BurstOqpskDemodulator d; std::vector<short> got; bool burstmode = false;
        d.processDemodulatedSoftBits = [&](const std::vector<short> &v) { got.insert(got.end(), v.begin(), v.end()); };
        d.BitRateChanged = [&](double, bool b) { burstmode = b; };
        const int afc_calls = log[4];
        d.setAFC(false);
        d.setSettings(BurstOqpskDemodulator::Settings());
        CHECK(burstmode && log[4] == afc_calls + 1);
        std::vector<int16_t> pcm(4000, 9);
        d.writeData((const char *)pcm.data(), 8000);
        CHECK(got.size() >= 32 && got[3] == -1);
        BurstMskDemodulator m; m.setSettings(BurstMskDemodulator::Settings()); CHECK(m.ok());
    }
    {   // ---- codec
        JConvolutionalCodec c; std::vector<uint16_t> poly; poly.push_back(109); poly.push_back(79);
        CHECK(c.SetCode(2, 7, poly, 24) && c.getPaddinglength() == 24);
        std::vector<uint8_t> soft(128, 200);
        CHECK(c.Decode_Continuous(soft).size() == 64 - 13);
        CHECK(c.Decode_soft(soft, 64).size() == 32);
        poly[0] = 91; CHECK(!c.SetCode(2, 7, poly));
    }
    {   // ---- failure path: the library cannot create the object
        mock_set_fail(1);
        OqpskDemodulator d; std::vector<std::string> warnings; int emits = 0;
        d.WarningTextSignal = [&](const std::string &s) { warnings.push_back(s); };
        d.processDemodulatedSoftBits = [&](const std::vector<short> &) { emits++; };
        d.setSettings(OqpskDemodulator::Settings());
        CHECK(!d.ok() && warnings.size() == 1 && warnings[0].find("CUDA") != std::string::npos);
        std::vector<int16_t> pcm(1000, 1);
        CHECK(d.writeData((const char *)pcm.data(), 2000) == 2000 && emits == 0 && d.getCurrentFreq() == 0);
        mock_set_fail(0);
    }
    printf("MOCK OK\n");
#else
    // real library, no CUDA device in this process: every create must fail loudly, nothing may be produced
    if (jaero_device_count() > 0) { printf("SKIP (a CUDA device is present)\n"); return 0; }
    std::vector<std::string> warnings; int emits = 0;
    OqpskDemodulator d;
    d.WarningTextSignal = [&](const std::string &s) { warnings.push_back(s); };
    d.processDemodulatedSoftBits = [&](const std::vector<short> &) { emits++; };
    d.setSettings(OqpskDemodulator::Settings());
    CHECK(!d.ok() && warnings.size() == 1 && !warnings[0].empty());
    std::vector<int16_t> pcm(9600, 100);
    CHECK(d.writeData((const char *)pcm.data(), 19200) == 19200 && emits == 0);
    BurstMskDemodulator b; b.WarningTextSignal = [&](const std::string &s) { warnings.push_back(s); };
    BurstMskDemodulator::Settings bs; bs.fb = 1200; bs.Fs = 48000; bs.lockingbw = 1800;
    b.setSettings(bs);
    CHECK(!b.ok() && warnings.size() == 2);
    JConvolutionalCodec c; std::vector<uint16_t> poly; poly.push_back(109); poly.push_back(79);
    CHECK(!c.SetCode(2, 7, poly, 24));
    printf("REAL-NO-GPU OK: %s\n", warnings[0].c_str());
#endif
    return 0;
}

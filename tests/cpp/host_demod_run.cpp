// Runs one of the host-mirror demodulator classes over a raw int16 PCM file the way the reference's tests drive the
// reference classes (setAFC / setSettings, writeData in fixed chunks, soft bits collected from the signal), and writes the
// soft bits as raw int16. Used by tests/test_host_mirror.py on the GPU box.
//   host_demod_run <oqpsk|msk|burst_oqpsk|burst_msk> <pcm.raw> <soft.raw> <chunk samples> <fb> <freq_center> <lockingbw> <signalthreshold> <afc>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "jaero_b200_host.hpp"

template <class D, class S> static int run(D &d, S st, const std::vector<int16_t> &pcm, size_t chunk, bool afc, const char *out)
{
    std::vector<short> soft; std::string warn; int status_true = 0;
    d.processDemodulatedSoftBits = [&](const std::vector<short> &v) { soft.insert(soft.end(), v.begin(), v.end()); };
    d.WarningTextSignal = [&](const std::string &s) { warn = s; };
    d.SignalStatus = [&](bool s) { status_true += s ? 1 : 0; };
    d.setAFC(afc);
    d.setSQL(false);
    d.setSettings(st);
    if (!d.ok()) { fprintf(stderr, "setSettings failed: %s\n", warn.c_str()); return 2; }
    for (size_t at = 0; at < pcm.size(); at += chunk) {
        const size_t n = pcm.size() - at < chunk ? pcm.size() - at : chunk;
        if (d.writeData((const char *)(pcm.data() + at), (jaero_b200::qint64)(n * 2)) != (jaero_b200::qint64)(n * 2)) return 3;
    }
    if (!warn.empty()) { fprintf(stderr, "warning raised: %s\n", warn.c_str()); return 4; }
    FILE *f = fopen(out, "wb"); if (!f) return 5;
    fwrite(soft.data(), sizeof(short), soft.size(), f); fclose(f);
    printf("%zu soft bits, SignalStatus(true) x%d, freq %.3f\n", soft.size(), status_true, d.getCurrentFreq());
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 10) { fprintf(stderr, "usage\n"); return 1; }
    const std::string kind = argv[1];
    FILE *f = fopen(argv[2], "rb"); if (!f) { perror("pcm"); return 1; }
    fseek(f, 0, SEEK_END); const long bytes = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<int16_t> pcm((size_t)bytes / 2);
    if (fread(pcm.data(), 2, pcm.size(), f) != pcm.size()) return 1;
    fclose(f);
    const size_t chunk = (size_t)atol(argv[4]);
    const double fb = atof(argv[5]), fc = atof(argv[6]), bw = atof(argv[7]), thr = atof(argv[8]);
    const bool afc = atoi(argv[9]) != 0;
    using namespace jaero_b200;
    if (kind == "oqpsk") { OqpskDemodulator d; OqpskDemodulator::Settings s; s.fb = fb; s.freq_center = fc; s.lockingbw = bw; s.signalthreshold = thr; return run(d, s, pcm, chunk, afc, argv[3]); }
    if (kind == "msk") { MskDemodulator d; MskDemodulator::Settings s; s.fb = fb; s.freq_center = fc; s.lockingbw = bw; s.signalthreshold = thr; return run(d, s, pcm, chunk, afc, argv[3]); }
    if (kind == "burst_oqpsk") { BurstOqpskDemodulator d; BurstOqpskDemodulator::Settings s; s.fb = fb; s.freq_center = fc; s.lockingbw = bw; s.signalthreshold = thr; return run(d, s, pcm, chunk, afc, argv[3]); }
    if (kind == "burst_msk") { BurstMskDemodulator d; BurstMskDemodulator::Settings s; s.fb = fb; s.Fs = 48000; s.freq_center = fc; s.lockingbw = bw; s.signalthreshold = thr; return run(d, s, pcm, chunk, afc, argv[3]); }
    fprintf(stderr, "unknown kind\n");
    return 1;
}

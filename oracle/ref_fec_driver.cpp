// TEST INFRASTRUCTURE ONLY. Driver around the reference's OWN frame-layer helper classes: AeroLcrc16, AeroLScrambler,
// PuncturedCode, DelayLine, AeroLInterleaver, PreambleDetector, PreambleDetectorPhaseInvariant,
// OQPSKPreambleDetectorAndAmbiguityCorrection and RTChannelDeleaveFECScram (JAERO/aerol.h:283-895, bodies
// JAERO/aerol.cpp:523-902 and :2505-2524), compiled VERBATIM. aerol.h / aerol.cpp as a whole need Qt GUI/SQL and cannot be built
// here, so oracle/Makefile slices these line ranges out of the files where they lie under /root/reference into a scratch directory
// (oracle/_ref/gen/, removed again after the compile; nothing of it is ever committed) and this file includes them.
// JConvolutionalCodec is the reference's own jconvolutionalcodec.cpp over the restated libcorrect, as in libjaero_ref.so.
// The restated frame-layer oracle (oracle/restated/fec_oracle.cpp) is pinned against this library by tests/test_fec_pinning.py.
#include "qt_shim.h"
#include <assert.h>
#include <math.h>
#include "jconvolutionalcodec.h"

#define private public
#include "gen/fec_types.h"
#undef private
#include "gen/fec_impl.inc"

extern "C" {

unsigned jfec_crc_bytes(const unsigned char *bytes, int n)
{
    AeroLcrc16 c; std::vector<char> b(bytes, bytes + n);
    return c.calcusingbytes(b.data(), n);
}
int jfec_crc_bits_check(const int *bits, int n)
{
    AeroLcrc16 c; std::vector<int> b(bits, bits + n);
    return c.calcusingbitsandcheck(b.data(), n) ? 1 : 0;
}
// AeroLScrambler: reset() then update() over successive pieces (sizes[]), as AeroL::Decode uses it
void jfec_scramble(int *bits, const int *sizes, int npieces)
{
    AeroLScrambler s; s.reset();
    int at = 0;
    for (int k = 0; k < npieces; k++) {
        QVector<int> v; for (int i = 0; i < sizes[k]; i++) v.push_back(bits[at + i]);
        s.update(v);
        for (int i = 0; i < sizes[k]; i++) bits[at + i] = v[i];
        at += sizes[k];
    }
}
void jfec_delayline(int length, int *data, const int *sizes, int npieces)
{
    DelayLine d; d.setLength(length);
    int at = 0;
    for (int k = 0; k < npieces; k++) {
        QVector<int> v; for (int i = 0; i < sizes[k]; i++) v.push_back(data[at + i]);
        d.update(v);
        for (int i = 0; i < sizes[k]; i++) data[at + i] = v[i];
        at += sizes[k];
    }
}
int jfec_deinterleave_ba(const int *block, int n, int setsize, int cols, unsigned char *out)
{
    AeroLInterleaver l; l.setSize(setsize);
    QVector<int> v; for (int i = 0; i < n; i++) v.push_back(block[i]);
    QByteArray &r = l.deinterleave_ba(v, cols);
    for (int i = 0; i < r.size(); i++) out[i] = (unsigned char)r.at(i);
    return r.size();
}
int jfec_deinterleave_msk_ba(const int *block, int n, int setsize, int blocks, unsigned char *out)
{
    AeroLInterleaver l; l.setSize(setsize);
    QVector<int> v; for (int i = 0; i < n; i++) v.push_back(block[i]);
    QByteArray &r = l.deinterleaveMSK_ba(v, blocks);
    for (int i = 0; i < r.size(); i++) out[i] = (unsigned char)r.at(i);
    return r.size();
}
int jfec_interleave(const int *block, int n, int setsize, int *out)
{
    AeroLInterleaver l; l.setSize(setsize);
    QVector<int> v; for (int i = 0; i < n; i++) v.push_back(block[i]);
    QVector<int> &r = l.interleave(v);
    for (int i = 0; i < r.size(); i++) out[i] = r[i];
    return r.size();
}
// PuncturedCode::depunture_soft_block over successive source blocks (reset only on the first, as DecodeC does at aerol.cpp:2306-2327)
int jfec_depuncture(const unsigned char *src, const int *sizes, int npieces, int pattern, unsigned char *out)
{
    PuncturedCode pc; QByteArray target; int at = 0;
    for (int k = 0; k < npieces; k++) {
        QByteArray s; for (int i = 0; i < sizes[k]; i++) s.push_back((char)src[at + i]);
        pc.depunture_soft_block(s, target, pattern, k == 0);
        at += sizes[k];
    }
    for (int i = 0; i < target.size(); i++) out[i] = (unsigned char)target.at(i);
    return target.size();
}
// kind 0: PreambleDetector (32-bit word), 1: PreambleDetectorPhaseInvariant (32-bit word, tolerance), 2: OQPSKPreambleDetectorAndAmbiguityCorrection
// (two 52-bit words, tolerance). out[i] = Update() return value, inv[i] = `inverted` after the update (kinds 1, 2).
void jfec_detect(int kind, unsigned long long w1, unsigned long long w2, int len, int tol, const int *bits, int n, int *out, int *inv)
{
    if (kind == 0) { PreambleDetector d; d.setPreamble(w1, len); for (int i = 0; i < n; i++) { out[i] = d.Update(bits[i]) ? 1 : 0; inv[i] = 0; } }
    else if (kind == 1) { PreambleDetectorPhaseInvariant d; d.setPreamble(w1, len); d.setTollerence(tol); for (int i = 0; i < n; i++) { out[i] = d.Update(bits[i]); inv[i] = d.inverted ? 1 : 0; } }
    else { OQPSKPreambleDetectorAndAmbiguityCorrection d; d.setPreamble(w1, w2, len); d.setTollerence(tol); for (int i = 0; i < n; i++) { out[i] = d.Update(bits[i]); inv[i] = d.inverted ? 1 : 0; } }
}
void *jfec_rt_new() { return new RTChannelDeleaveFECScram(); }
void jfec_rt_free(void *h) { delete (RTChannelDeleaveFECScram *)h; }
int jfec_rt_reset(void *h) { return (int)((RTChannelDeleaveFECScram *)h)->resetblockptr(); }
int jfec_rt_update(void *h, int msk, int soft) { RTChannelDeleaveFECScram *r = (RTChannelDeleaveFECScram *)h; return (int)(msk ? r->updateMSK(soft) : r->update(soft)); }
int jfec_rt_info(void *h, unsigned char *out, int cap, int *numberofsus)
{
    RTChannelDeleaveFECScram *r = (RTChannelDeleaveFECScram *)h;
    int n = r->infofield.size() < cap ? r->infofield.size() : cap;
    for (int i = 0; i < n; i++) out[i] = (unsigned char)r->infofield.at(i);
    *numberofsus = r->numberofsus;
    return r->infofield.size();
}

}

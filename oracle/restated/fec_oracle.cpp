// TEST INFRASTRUCTURE ONLY (oracle): CPU restatement, never linked into the product.
#include "fec_oracle.h"
#include <cstring>
#include <cmath>

// JAERO/aerol.cpp:531-535: interleaverowdepermute[i]=(i*27)%64 ; :613-621 out[k++]=block[depermute[i]*cols+j]
void oracle_deinterleave(const int *block, int cols, uint8_t *out)
{
    int k = 0;
    for (int j = 0; j < cols; j++)
        for (int i = 0; i < 64; i++)
            out[k++] = (uint8_t)block[((i * 27) % 64) * cols + j];
}

ContinuousViterbiOracle::ContinuousViterbiOracle(int pad) : paddinglength(pad)
{
    correct_convolutional_polynomial_t poly[2] = {109, 79};    // jconvolutionalcodec.cpp:13-14
    conv = correct_convolutional_create(2, 7, poly);
}
ContinuousViterbiOracle::~ContinuousViterbiOracle() { correct_convolutional_destroy(conv); }

std::vector<int> ContinuousViterbiOracle::decode(const uint8_t *soft, int n)
{
    const int k = 62;                                           // jconvolutionalcodec.cpp:153
    std::vector<uint8_t> buf(overlap);                          // :155 overlap ‖ block
    buf.insert(buf.end(), soft, soft + n);
    buf.insert(buf.end(), paddinglength, (uint8_t)128);         // :158-160 erasure padding
    std::vector<uint8_t> decoded(buf.size() / 2 + 1, 0);        // :167
    correct_convolutional_decode_soft(conv, buf.data(), buf.size(), decoded.data());   // :169
    int dbits = (int)buf.size() / 2;                            // :172
    std::vector<int> bits(dbits, 0);
    for (int i = 0; i < dbits; i++) bits[i] = (decoded[i >> 3] >> (7 - (i & 7))) & 1;  // :177-190 MSB first
    std::vector<int> out;                                       // :194 mid(paddinglength+1, n/2)
    int pos = paddinglength + 1;
    for (int i = 0; i < n / 2 && pos + i < dbits; i++) out.push_back(bits[pos + i]);
    int kk = k < n ? k : n;                                     // :197-198 right(k) of the *new* block, resized to k
    overlap.assign(soft + n - kk, soft + n);
    overlap.resize(k, 0);
    return out;
}

uint16_t oracle_crc16(const uint8_t *bytes, int n)
{
    uint16_t crc = 0xFFFF;
    for (int i = 0; i < n; i++) {
        int b = (int8_t)bytes[i];                               // `message_byte=bytes[i]` with char bytes (sign irrelevant for 8 shifts)
        for (int k = 0; k < 8; k++) {
            int mb = b & 1; b >>= 1;
            int cb = crc & 1; crc >>= 1;
            if (cb ^ mb) crc ^= 0x8408;
        }
    }
    return (uint16_t)~crc;
}

PChannelOracle::PChannelOracle(int fb) : codec(24)
{
    ifb = fb;
    switch (ifb) {                                              // aerol.cpp:1013-1052
    case 600:  cols = 6;  NumberOfBits = 1152; BitsInHeader = 16; TotalNumberOfBits = 16 + 1152 + 32; useingOQPSK = false; break;
    case 10500: cols = 78; NumberOfBits = 4992; BitsInHeader = 16 + 178; TotalNumberOfBits = 16 + 178 + 4992 + 64; useingOQPSK = true; break;
    default:   cols = 9;  NumberOfBits = 1152; BitsInHeader = 16; TotalNumberOfBits = 16 + 1152 + 32; useingOQPSK = false; break;
    }
    block.assign(cols * 64, 0);
    int dl2len = (ifb == 10500) ? 4992 - 6 : 576 - 6;           // aerol.cpp:1018,1026,1047
    dl2.assign(dl2len + 1, 0); dl2_ptr = 0;
    // scrambler (aerol.h:397-419)
    int st[15] = {1, 1, 0, 1, 0, 0, 1, 0, 1, 0, 1, 1, 0, 0, 1};
    scr.resize(5000);
    for (int a = 0; a < 5000; a++) {
        int v = st[0] ^ st[14];
        scr[a] = v;
        for (int i = 14; i > 0; i--) st[i] = st[i - 1];
        st[0] = v;
    }
    scr_pos = 0;
    preamble.clear();                                           // 0xE15AE893 MSB first (aerol.cpp:730-743,947)
    for (int i = 31; i >= 0; i--) preamble.push_back((3780831379ULL >> i) & 1);
    buf_plain.assign(32, 0); buf_imag.assign(32, 0); buf_real.assign(32, 0);
    inv_imag = inv_real = false;
    realimag = 0; gotsync_last = 0;
    cntr = 1000000000; blockcnt = -1;
    frameinfo = lastframeinfo = 0; formatid = supfrmaker = framecounter1 = framecounter2 = 0;
    datacdcountdown = 0; datacd = false;
    nframes = 0; bits_seen = 0;
}

// PreambleDetectorPhaseInvariant::Update with tollerence 0 (aerol.cpp:781-804, non-burst :1003-1005)
static int uw_update_invariant(std::vector<int> &buf, const std::vector<int> &pre, int val, bool &inverted)
{
    int n = (int)buf.size(), xorsum = 0;
    for (int i = 0; i < n - 1; i++) { buf[i] = buf[i + 1]; xorsum += buf[i] ^ pre[i]; }
    xorsum += val ^ pre[n - 1];
    buf[n - 1] = val;
    if (xorsum >= n) { inverted = true; return 1; }
    if (xorsum <= 0) { inverted = false; return 1; }
    return 0;
}
// PreambleDetector::Update (aerol.cpp:744-750): exact match, buffer cleared on hit
static int uw_update_exact(std::vector<int> &buf, const std::vector<int> &pre, int val)
{
    int n = (int)buf.size();
    for (int i = 0; i < n - 1; i++) buf[i] = buf[i + 1];
    buf[n - 1] = val;
    if (buf == pre) { std::fill(buf.begin(), buf.end(), 0); return 1; }
    return 0;
}

void PChannelOracle::updateDCD()
{
    if (datacdcountdown > 0) datacdcountdown -= 3;
    else if (datacdcountdown < 0) datacdcountdown = 0;
    if (datacd && !datacdcountdown) { datacd = false; dcd_events.push_back((bits_seen << 1) | 0); }
}
void PChannelOracle::lostSignal()
{
    cntr = 1000000000; datacdcountdown = 0; datacd = false; dcd_events.push_back((bits_seen << 1) | 0);
}

void PChannelOracle::process(const short *bits, int n)
{
    for (int i = 0; i < n; i++) {
        bits_seen++;
        int bit = (((uint8_t)bits[i]) >= 128) ? 1 : 0;          // aerol.cpp:1136-1139
        int soft_bit = (uint16_t)bits[i];
        if (bits[i] < 0) continue;                              // burst marker (never in continuous modes)
        int gotsync;
        if (useingOQPSK) {                                      // aerol.cpp:1156-1233
            realimag++; realimag %= 2;
            std::vector<int> &buf = realimag ? buf_imag : buf_real;
            bool &inv = realimag ? inv_imag : inv_real;
            if (cntr > NumberOfBits - 68 || cntr <= 0 || !datacd) {
                gotsync = uw_update_invariant(buf, preamble, bit, inv);
                if (!gotsync_last) { gotsync_last = gotsync; gotsync = 0; } else gotsync_last = 0;
            } else { gotsync = 0; gotsync_last = 0; }
            if (inv) { bit = 1 - bit; if (soft_bit != 128) soft_bit = 255 - soft_bit; }
        } else gotsync = uw_update_exact(buf_plain, preamble, bit);   // aerol.cpp:1269-1272

        if (cntr < 1000000000) cntr++;
        if (cntr < 16) {                                        // :1275-1300 header
            if (cntr == 0) { frameinfo = (uint16_t)bit; infofield.clear(); }
            else { frameinfo = (uint16_t)((frameinfo << 1) | bit); }
        }
        if (cntr == 15) {                                       // :1301-1319 (delayed by one frame)
            uint16_t t = frameinfo; frameinfo = lastframeinfo; lastframeinfo = t;
            formatid = (frameinfo >> 12) & 15; supfrmaker = (frameinfo >> 8) & 15;
            framecounter1 = (frameinfo >> 4) & 15; framecounter2 = frameinfo & 15;
        }
        if (cntr >= 16) {                                       // :1540-1610
            if (cntr == 16) blockcnt = -1;
            int idx = (int)((cntr - BitsInHeader) % (long)block.size());
            if (idx < 0) idx = 0;
            block[idx] = soft_bit;
            if (idx == (int)block.size() - 1) {
                blockcnt++;
                std::vector<uint8_t> dl(block.size());
                oracle_deinterleave(block.data(), cols, dl.data());
                std::vector<int> dec = codec.decode(dl.data(), (int)dl.size());
                for (size_t h = 0; h < dec.size(); h++) {       // dl2.update (aerol.h:465-473)
                    dl2[dl2_ptr] = dec[h]; dl2_ptr++; dl2_ptr %= (int)dl2.size(); dec[h] = dl2[dl2_ptr];
                }
                for (size_t h = 0; h < dec.size(); h++) { dec[h] ^= scr[scr_pos]; scr_pos++; }   // scrambler.update
                int charptr = 0; uint8_t ch = 0;                // :1568-1580 LSB-first packing
                for (size_t h = 0; h < dec.size(); h++) {
                    ch |= (uint8_t)(dec[h] * 128);
                    charptr++; charptr %= 8;
                    if (charptr == 0) { infofield.push_back(ch); ch = 0; } else ch >>= 1;
                }
                if ((cntr - BitsInHeader) == (NumberOfBits - 1)) {   // :1582 frame done
                    for (int k = 0; k < (int)infofield.size() / 12; k++) {
                        uint16_t crc_calc = oracle_crc16(&infofield[k * 12], 10);
                        uint16_t crc_rec = (uint16_t)((infofield[k * 12 + 11] << 8) | infofield[k * 12 + 10]);
                        if ((!crc_rec) && (crc_calc != crc_rec)) {
                            int tsum = 0; for (int ii = 0; ii < 10; ii++) tsum += infofield[k * 12 + ii];
                            if (tsum == 0) crc_calc = 0;
                        }
                        if (crc_calc == crc_rec) { if (datacdcountdown < 12) datacdcountdown += 2; }
                        else { if (datacdcountdown > 0) datacdcountdown -= 3; }
                        if (!datacd && datacdcountdown > 2) { datacd = true; dcd_events.push_back((bits_seen << 1) | 1); }
                        SignalUnit su; memcpy(su.bytes, &infofield[k * 12], 12);
                        su.crc_ok = (crc_calc == crc_rec); su.frame = nframes;
                        sus.push_back(su);
                    }
                    nframes++;
                }
            }
        }
        if (gotsync) {                                          // :1990-2011
            cntr = -1; datacd = true; datacdcountdown = 12; dcd_events.push_back((bits_seen << 1) | 1);
            scr_pos = 0;
        }
        if (cntr + 1 == TotalNumberOfBits) { scr_pos = 0; cntr = -1; }   // :2013-2016
    }
}

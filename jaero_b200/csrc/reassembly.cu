// ISU / SSU reassembly and ACARS block parsing (host side). Pure host code: this is the
// per-aircraft message layer that follows the CRC-checked signal units the device produces; it handles a few
// hundred bytes per second per channel and stays on the CPU in the reference as well.
//
// Behaviour follows the reference's AeroL helpers, including their quirks (each noted where it applies):
//   P/T-channel initial signal unit (0x71) + subsequent signal units   ISUData::update      JAERO/aerol.cpp:151-214
//   R-channel 1..3-SU user-data sequences                             RISUData::update     JAERO/aerol.cpp:27-112
//   ACARS framing inside the user data, parity checks                 ParserISU::parse     JAERO/aerol.cpp:340-487
//   multi-block ACARS messages                                        ACARSDefragmenter    JAERO/aerol.cpp:221-329
//   which signal units are passed on                                  AeroL::Decode        JAERO/aerol.cpp:1357-1399,1497-1513,1900-1925
// Layout is this library's own: flat entries in small vectors (at most ~11 sequences are open per channel).
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <new>
#include <string>
#include <vector>

#include "../../include/jaero_b200.h"

namespace {

typedef std::vector<uint8_t> Bytes;

struct Sequence                      // one partly received ISU
{
    uint32_t aes = 0;
    uint8_t ges = 0, qno = 0, seqno = 0, refno = 0, last_octets = 0;
    int age = 0;
    Bytes data;
    // R channel only
    int seq_indicator = 0, su_type = 0, filled = 0;
};

// writing one past the end grows the array (the reference relies on QByteArray doing so); new bytes are zero
inline void put(Bytes &b, int at, uint8_t v)
{
    if (at >= (int)b.size()) b.resize((size_t)at + 1, 0);
    b[(size_t)at] = v;
}

template <class T> void age_out(std::vector<T> &v, int limit)
{
    size_t w = 0;
    for (size_t i = 0; i < v.size(); i++)
        if (++v[i].age <= limit) { if (w != i) v[w] = v[i]; w++; }
    v.resize(w);
}

// ---- P / T channel: 0x71 + SSUs -------------------------------------------------------------------------------------
struct PtAssembler
{
    std::vector<Sequence> open;
    Sequence head;                   // header of the LAST initial SU seen: subsequent SUs are matched against its AES / GES
    Sequence done;
    bool missing = false;

    bool push(const uint8_t *su)     // 10 bytes
    {
        missing = false;
        const uint8_t kind = su[0];
        if (kind == 0x71)
        {
            age_out(open, 10);
            head.aes = (uint32_t)su[1] << 16 | (uint32_t)su[2] << 8 | su[3];
            head.ges = su[4];
            head.qno = su[5] >> 4; head.refno = su[5] & 15;
            head.seqno = su[6] & 0x3F;
            head.last_octets = su[7] >> 4;
            head.age = 0;
            head.data.assign(su + 8, su + 10);
            int at = -1;
            if (head.last_octets <= 8)
                for (size_t i = 0; i < open.size() && at < 0; i++)
                    if (open[i].aes == head.aes && open[i].ges == head.ges && open[i].qno == head.qno && open[i].refno == head.refno) at = (int)i;
            if (at < 0) open.push_back(head); else open[(size_t)at] = head;
            return false;
        }
        if ((kind & 0xC0) != 0xC0) return false;
        head.seqno = kind & 0x3F;
        head.qno = su[1] >> 4; head.refno = su[1] & 15;
        int at = -1;
        if (head.last_octets <= 8)
            for (size_t i = 0; i < open.size() && at < 0; i++)
                if (open[i].aes == head.aes && open[i].ges == head.ges && (int)open[i].seqno == (int)head.seqno + 1 &&
                    open[i].qno == head.qno && open[i].refno == head.refno) at = (int)i;
        if (at < 0) { missing = true; return false; }
        Sequence &s = open[(size_t)at];
        s.seqno--;
        if (s.seqno == 0)
        {
            for (int i = 2; i <= s.last_octets + 1; i++) s.data.push_back(i < 10 ? su[i] : 0);   // reads past the SU yield 0
            done = s;                // the finished sequence stays in the list until it ages out
            return true;
        }
        s.data.insert(s.data.end(), su + 2, su + 10);
        return false;
    }
};

// ---- R channel: 1..3 SUs of up to 11 bytes ---------------------------------------------------------------------------
struct RAssembler
{
    std::vector<Sequence> open;
    Sequence done;

    bool push(const uint8_t *p)      // 17 bytes
    {
        age_out(open, 10);
        Sequence s;
        s.seq_indicator = p[0] >> 4; s.su_type = p[0] & 15;
        s.qno = p[1] >> 4; s.refno = p[1] & 7;
        s.aes = (uint32_t)p[2] << 16 | (uint32_t)p[3] << 8 | p[4];
        s.ges = p[5];
        int at = -1;
        if (s.su_type >= 1 && s.su_type <= 11)
            for (size_t i = 0; i < open.size() && at < 0; i++)
                if (open[i].ges == s.ges && open[i].aes == s.aes && open[i].qno == s.qno && open[i].refno == s.refno) at = (int)i;
        if (at < 0) { open.push_back(s); at = (int)open.size() - 1; }
        Sequence &e = open[(size_t)at];
        e.age = 0;
        static const int total_of[7] = {0, 1, 2, 2, 3, 3, 3}, index_of[7] = {0, 0, 0, 1, 0, 1, 2};
        const int total = s.seq_indicator <= 6 ? total_of[s.seq_indicator] : 0;
        const int index = s.seq_indicator <= 6 ? index_of[s.seq_indicator] : 0;
        const int nbytes = (s.su_type >= 1 && s.su_type <= 11) ? s.su_type : 0;
        const bool signalling = s.su_type == 15;
        const int want = 11 * total - 11 + nbytes;
        if (want > 0)
        {
            if (e.data.empty()) e.data.resize((size_t)want, 0);
            if (want < (int)e.data.size()) e.data.resize((size_t)want);
        }
        if (!signalling)
        {
            for (int i = 0; i < nbytes; i++) put(e.data, 11 * index + i, p[6 + i]);
            e.filled |= 1 << index;
        }
        else e.data.clear();
        if (signalling || (e.filled == 7 && total == 3) || (e.filled == 3 && total == 2) || (e.filled == 1 && total == 1))
        {
            done = e;
            open.erase(open.begin() + at);
            return true;
        }
        return false;
    }
};

// ---- ACARS ---------------------------------------------------------------------------------------------------------
struct Block
{
    Sequence isu;
    uint8_t mode = 0, tak = 0, bi = 0;
    std::string label, reg, text;
    bool nonacars = false, downlink = false, valid = false, hastext = false, more = false;
    int age = 0;                     // defragmenter bookkeeping
};

struct Output { int kind; Block b; std::string error; };

struct Session
{
    PtAssembler pt;
    RAssembler r;
    std::vector<Block> frags;
    std::deque<Output> out;
    uint64_t n_isu = 0, n_missing = 0, n_errors = 0, n_acars = 0;

    void emit(Block b)               // the aircraft-database look-up is not part of this library: only its dot removal
    {
        size_t i = 0; while (i < b.reg.size() && b.reg[i] == '.') i++;
        b.reg.erase(0, i);
        Output o; o.kind = JAERO_REASM_ACARS; o.b = b;
        out.push_back(o); n_acars++;
    }
    void fail(const Sequence &isu, const std::string &what)
    {
        Output o; o.kind = JAERO_REASM_ERROR; o.b.isu = isu; o.error = what;
        out.push_back(o); n_errors++;
    }
    bool parity_fail(const Sequence &isu)
    {
        char t[128];
        snprintf(t, sizeof t, "ISU: AESID = %X GESID = %X QNO = %02X REFNO = %02X : Parity error", isu.aes, isu.ges, isu.qno, isu.refno);
        fail(isu, t);
        return false;
    }

    // true when the block completes a message (which then replaces b)
    bool defragment(Block &b)
    {
        age_out(frags, 30);
        int at = -1;
        for (size_t i = 0; i < frags.size() && at < 0; i++)
        {
            const Block &f = frags[i];
            if (!(b.reg == f.reg && b.label == f.label && b.mode == f.mode && b.isu.aes == f.isu.aes && b.isu.ges == f.isu.ges && f.more)) continue;
            if (b.tak != f.tak) continue;
            const uint8_t next = (uint8_t)((((int)f.bi + 1 - 'A') % 26) + 'A');
            if (next == b.bi) at = (int)i;
        }
        if (at < 0)
        {
            if (!b.more) return true;
            b.age = 0; frags.push_back(b);
            return false;
        }
        Block &f = frags[(size_t)at];
        f.age = 0; f.bi = b.bi; f.text += b.text; f.more = b.more;
        if (b.more) return false;
        b = f;
        frags.erase(frags.begin() + at);
        return true;
    }

    bool parse(const Sequence &isu, bool downlink)
    {
        n_isu++;
        if (isu.aes == 0) { fail(isu, "Error: AESID == 0"); return false; }
        const Bytes &u = isu.data;
        const size_t n = u.size();
        Block b; b.isu = isu; b.downlink = downlink;
        const bool acars = n > 16 && u[0] == 0xFF && u[1] == 0xFF && (u[15] == 0x83 || u[15] == 0x02);
        if (!acars)
        {
            static const char hex[] = "0123456789ABCDEF";
            b.nonacars = true; b.valid = true;
            for (size_t i = 0; i < n; i++) { b.text += hex[u[i] >> 4]; b.text += hex[u[i] & 15]; }
            emit(b);
            return true;
        }
        auto odd = [&](size_t k) { return (__builtin_popcount(u[k]) & 1) != 0; };
        b.mode = u[3] & 0x7F; b.tak = u[11] & 0x7F;
        b.label.push_back((char)(u[12] & 0x7F)); b.label.push_back((char)(u[13] & 0x7F));
        b.bi = u[14] & 0x7F;
        b.hastext = u[15] == 0x02;
        b.more = u[n - 4] == 0x97;
        for (size_t k = 4; k < 11; k++)
        {
            if (!odd(k)) return parity_fail(isu);
            b.reg.push_back((char)(u[k] & 0x7F));
        }
        if (b.hastext)
            for (size_t k = 16; k + 4 < n; k++)
            {
                if (!odd(k)) return parity_fail(isu);
                const uint8_t c = u[k] & 0x7F;
                if (c == 0x7F) b.text += "<DEL>"; else b.text.push_back((char)c);
            }
        b.valid = true;
        if (defragment(b)) emit(b);
        return true;
    }
};

}  // namespace

struct jaero_reasm { Session s; };

extern "C" {

int jaero_reasm_create(jaero_reasm **out)
{
    if (!out) return JAERO_E_ARG;
    *out = new (std::nothrow) jaero_reasm();
    return *out ? JAERO_OK : JAERO_E_ARG;
}
void jaero_reasm_destroy(jaero_reasm *h) { delete h; }

int jaero_reasm_reset(jaero_reasm *h)
{
    if (!h) return JAERO_E_ARG;
    h->s.pt.open.clear(); h->s.r.open.clear();
    return JAERO_OK;
}
int jaero_reasm_short_frame(jaero_reasm *h)
{
    if (!h) return JAERO_E_ARG;
    h->s.pt.open.clear();
    return JAERO_OK;
}

int jaero_reasm_push_su(jaero_reasm *h, const uint8_t *su, int downlink)
{
    if (!h || !su) return JAERO_E_ARG;
    Session &s = h->s;
    if (su[0] == 0x71) { s.pt.push(su); return 0; }
    if ((su[0] & 0xC0) != 0xC0) return 0;
    int rc = 0;
    if (s.pt.push(su))
    {
        rc |= JAERO_REASM_COMPLETE;
        if (s.parse(s.pt.done, downlink != 0)) rc |= JAERO_REASM_PARSED;
    }
    else if (s.pt.missing) { rc |= JAERO_REASM_MISSING; s.n_missing++; }
    return rc;
}

int jaero_reasm_push_r(jaero_reasm *h, const uint8_t *info, int downlink)
{
    if (!h || !info) return JAERO_E_ARG;
    Session &s = h->s;
    if ((info[1] & 0x08) != 0x08) return 0;
    int rc = 0;
    if (s.r.push(info))
    {
        rc |= JAERO_REASM_COMPLETE;
        if (s.parse(s.r.done, downlink != 0)) rc |= JAERO_REASM_PARSED;
    }
    return rc;
}

int jaero_reasm_push_t_packet(jaero_reasm *h, const uint8_t *info, int n_sus)
{
    if (!h || !info || n_sus < 0) return JAERO_E_ARG;
    int rc = 0;
    for (int k = 0; k < n_sus; k++)
    {
        const uint8_t *su = info + 6 + 12 * k;
        if (su[0] == 0x01) continue;                     // fill-in signal unit
        rc |= jaero_reasm_push_su(h, su, 1);
    }
    return rc;
}

int jaero_reasm_pending(const jaero_reasm *h) { return h ? (int)h->s.out.size() : JAERO_E_ARG; }

long jaero_reasm_pop(jaero_reasm *h, jaero_acars_record *rec, char *text, size_t cap)
{
    if (!h || !rec) return JAERO_E_ARG;
    Session &s = h->s;
    if (s.out.empty()) return -1;
    const Output &o = s.out.front();
    const std::string &t = o.kind == JAERO_REASM_ERROR ? o.error : o.b.text;
    memset(rec, 0, sizeof *rec);
    rec->text_len = (uint32_t)t.size();
    if (t.size() > cap || (t.size() && !text)) return -2;   // rec->text_len says how much room is needed; nothing is popped
    rec->kind = o.kind;
    rec->aes_id = o.b.isu.aes; rec->ges_id = o.b.isu.ges; rec->qno = o.b.isu.qno; rec->refno = o.b.isu.refno;
    rec->seqno = o.b.isu.seqno; rec->last_octets = o.b.isu.last_octets;
    rec->mode = o.b.mode; rec->tak = o.b.tak; rec->block_id = o.b.bi;
    rec->label_len = (uint8_t)o.b.label.size(); memcpy(rec->label, o.b.label.data(), o.b.label.size() < 2 ? o.b.label.size() : 2);
    rec->reg_len = (uint8_t)o.b.reg.size(); memcpy(rec->reg, o.b.reg.data(), o.b.reg.size() < 7 ? o.b.reg.size() : 7);
    rec->flags = (o.b.nonacars ? JAERO_ACARS_NONACARS : 0) | (o.b.downlink ? JAERO_ACARS_DOWNLINK : 0) | (o.b.valid ? JAERO_ACARS_VALID : 0) |
                 (o.b.hastext ? JAERO_ACARS_HASTEXT : 0) | (o.b.more ? JAERO_ACARS_MORE : 0);
    rec->userdata_len = (uint32_t)o.b.isu.data.size();
    if (!t.empty()) memcpy(text, t.data(), t.size());
    s.out.pop_front();
    return (long)t.size();
}

int jaero_reasm_get_stats(const jaero_reasm *h, uint64_t *isus, uint64_t *messages, uint64_t *errors, uint64_t *missing)
{
    if (!h) return JAERO_E_ARG;
    if (isus) *isus = h->s.n_isu;
    if (messages) *messages = h->s.n_acars;
    if (errors) *errors = h->s.n_errors;
    if (missing) *missing = h->s.n_missing;
    return JAERO_OK;
}

}  // extern "C"

// TEST INFRASTRUCTURE ONLY. Driver around the reference's OWN ISU/SSU reassembly code: RISUData, ISUData,
// ACARSDefragmenter (JAERO/aerol.cpp:4-329) and ParserISU::parse (JAERO/aerol.cpp:340-487), compiled VERBATIM.
// aerol.cpp / aerol.h as a whole need Qt GUI/SQL/network and cannot be built here, so the Makefile slices the
// two line ranges out of the files where they lie under /root/reference into a scratch directory (oracle/_ref/gen/,
// removed again after the compile; nothing of it is ever committed) and this file includes them. Hand-written here: the ParserISU constructor, the two signal bodies, a
// stand-in for the aircraft-database look-up (no database: the look-up result is empty, JAERO/aerol.cpp:493-520)
// and the three-line SU dispatch of AeroL::Decode (JAERO/aerol.cpp:1357-1399, 1497-1513, 1900-1925).
#include "qt_shim.h"
#include <deque>

class DBase {};
class DataBaseTextUser : public QObject
{
public:
    explicit DataBaseTextUser(QObject *p = 0) : QObject(p) {}
    void request(const QString &dir, const QString &aesid, DBase *userdata);
};

#define private public
#include "gen/reasm_types.h"
#undef private
#include "gen/reasm_impl.inc"

namespace {
struct Event
{
    int kind;            // 0 ACARS item, 1 error string
    ACARSItem item;
    std::string text;    // error text for kind 1
};
struct RefReasm : QObject
{
    ISUData isudata;
    RISUData risudata;
    ParserISU *parserisu;
    std::deque<Event> q;
    RefReasm() { parserisu = new ParserISU(this); }
    ~RefReasm() { delete parserisu->dbtu; delete parserisu; }
};
}

ParserISU::ParserISU(QObject *parent) : QObject(parent)       // JAERO/aerol.cpp:331-338 without the signal hookup
{
    downlink = false;
    dbtu = new DataBaseTextUser(this);
}
void ParserISU::ACARSsignal(ACARSItem &acarsitem)
{
    Event e; e.kind = 0; e.item = acarsitem;
    static_cast<RefReasm *>(parent())->q.push_back(e);
}
void ParserISU::Errorsignal(QString &error)
{
    Event e; e.kind = 1; e.text = error.s;
    static_cast<RefReasm *>(parent())->q.push_back(e);
}
// The look-up answers at once with an empty result: what is left of acarslookupresult (JAERO/aerol.cpp:493-520)
// is the removal of the leading dots of the registration and the emit.
void DataBaseTextUser::request(const QString &, const QString &, DBase *userdata)
{
    ParserISU *ps = static_cast<ParserISU *>(parent());
    ACARSItem *it = static_cast<ACARSItem *>(userdata);
    int i = 0; while ((i < it->PLANEREG.size()) && (it->PLANEREG[i] == '.')) i++;
    it->PLANEREG = it->PLANEREG.right(it->PLANEREG.size() - i);
    ps->ACARSsignal(*it);
    delete it;
}

extern "C" {
void *jref_reasm_new(void) { return new RefReasm(); }
void jref_reasm_free(void *h) { delete static_cast<RefReasm *>(h); }
void jref_reasm_reset(void *h)                                 // AeroL::setSettings (JAERO/aerol.cpp:992-993)
{
    RefReasm *r = static_cast<RefReasm *>(h);
    r->isudata.reset(); r->risudata.reset();
}
void jref_reasm_short_frame(void *h) { static_cast<RefReasm *>(h)->isudata.reset(); }   // JAERO/aerol.cpp:1997

// one CRC-valid P- or T-channel SU (10 bytes without the CRC). bit0: an ISU completed, bit1: missing SSU,
// bit2: parse() returned true
int jref_reasm_su(void *h, const unsigned char *su, int burstmode)
{
    RefReasm *r = static_cast<RefReasm *>(h);
    QByteArray d((const char *)su, 10);
    int message = su[0], rc = 0;
    if (message == AEROTypeP::User_data_ISU_RLS_P_T_channel) { r->isudata.update(d); return 0; }
    if ((message & 0xC0) != 0xC0) return 0;
    if (r->isudata.update(d))
    {
        rc |= 1;
        r->parserisu->downlink = burstmode;
        if (r->parserisu->parse(r->isudata.lastvalidisuitem)) rc |= 4;
    }
    else if (r->isudata.missingssu) rc |= 2;
    return rc;
}
// one CRC-valid R-channel packet (the 17 information bytes); ignored unless it is a user-data SU (bit 3 of byte 2)
int jref_reasm_r(void *h, const unsigned char *info, int burstmode)
{
    RefReasm *r = static_cast<RefReasm *>(h);
    if ((info[1] & 0x08) != 0x08) return 0;
    int rc = 0;
    if (r->risudata.update(QByteArray((const char *)info, 17)))
    {
        rc |= 1;
        r->parserisu->downlink = burstmode;
        if (r->parserisu->parse(r->risudata.lastvalidisuitem)) rc |= 4;
    }
    return rc;
}
int jref_reasm_pending(void *h) { return (int)static_cast<RefReasm *>(h)->q.size(); }
// meta: kind, AESID, GESID, QNO, REFNO, SEQNO, NOOCT, MODE, TAK, BI, flags(nonacars|downlink<<1|valid<<2|hastext<<3|
// moretocome<<4), label_len, reg_len, message_len, userdata_len. text: label | reg | message | userdata.
// Returns the bytes written to text, -1 when the queue is empty, -2 when cap is too small (nothing popped).
int jref_reasm_pop(void *h, unsigned *meta, unsigned char *text, int cap)
{
    RefReasm *r = static_cast<RefReasm *>(h);
    if (r->q.empty()) return -1;
    const Event &e = r->q.front();
    const ACARSItem &a = e.item;
    const std::string &msg = e.kind ? e.text : a.message.s;
    int need = a.LABEL.size() + a.PLANEREG.size() + (int)msg.size() + a.isuitem.userdata.size();
    if (need > cap) return -2;
    meta[0] = e.kind; meta[1] = a.isuitem.AESID; meta[2] = a.isuitem.GESID; meta[3] = a.isuitem.QNO; meta[4] = a.isuitem.REFNO;
    meta[5] = a.isuitem.SEQNO; meta[6] = a.isuitem.NOOCTLESTINLASTSSU; meta[7] = (unsigned char)a.MODE; meta[8] = a.TAK; meta[9] = a.BI;
    meta[10] = (a.nonacars ? 1 : 0) | (a.downlink ? 2 : 0) | (a.valid ? 4 : 0) | (a.hastext ? 8 : 0) | (a.moretocome ? 16 : 0);
    meta[11] = a.LABEL.size(); meta[12] = a.PLANEREG.size(); meta[13] = (unsigned)msg.size(); meta[14] = a.isuitem.userdata.size();
    int o = 0;
    for (int i = 0; i < a.LABEL.size(); i++) text[o++] = (unsigned char)a.LABEL[i];
    for (int i = 0; i < a.PLANEREG.size(); i++) text[o++] = (unsigned char)a.PLANEREG[i];
    for (size_t i = 0; i < msg.size(); i++) text[o++] = (unsigned char)msg[i];
    for (int i = 0; i < a.isuitem.userdata.size(); i++) text[o++] = (unsigned char)a.isuitem.userdata[i];
    r->q.pop_front();
    return o;
}
}

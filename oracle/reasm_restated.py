"""TEST INFRASTRUCTURE ONLY. CPU restatement (plain Python) of the reference's ISU/SSU reassembly and ACARS block
parsing, each function citing the reference lines it follows. Pinned against the reference's own code compiled verbatim
(oracle/_ref/libjaero_ref_reasm.so, oracle/ref_reasm_driver.cpp) through tests/golden/reasm_golden.json: identical
records and return codes on every committed stream (tests/test_reassembly.py). Only tests may import this."""


class _Item:
    """ISUItem / RISUItem (JAERO/aerol.h:113-145)."""

    def __init__(self):
        self.AESID = 0; self.GESID = 0; self.QNO = 0; self.SEQNO = 0; self.REFNO = 0; self.NOOCT = 0
        self.userdata = bytearray(); self.count = 0
        self.SEQINDICATOR = 0; self.SUTYPE = 0; self.filledarray = 0

    def copy(self):
        c = _Item(); c.__dict__.update(self.__dict__); c.userdata = bytearray(self.userdata); return c


def _put(b, i, v):
    """QByteArray's non-const operator[] grows the array when assigned past the end (zero fill here, as in oracle/shim)."""
    if i >= len(b):
        b.extend(bytes(i + 1 - len(b)))
    b[i] = v


def _age(items, limit):
    """deleteoldisuitems (JAERO/aerol.cpp:16-26, 139-149): count++, drop above the limit."""
    i = 0
    while i < len(items):
        items[i].count += 1
        if items[i].count > limit:
            del items[i]
        else:
            i += 1


class RISUData:
    """JAERO/aerol.cpp:6-112."""

    def __init__(self):
        self.isuitems = []; self.lastvalidisuitem = _Item()

    def reset(self):
        self.isuitems = []

    def _find(self, a):                                    # :6-15
        if a.SUTYPE > 11 or a.SUTYPE < 1:
            return -1
        for i, it in enumerate(self.isuitems):
            if a.GESID == it.GESID and a.AESID == it.AESID and a.QNO == it.QNO and a.REFNO == it.REFNO:
                return i
        return -1

    def update(self, data):                                # :27-112
        _age(self.isuitems, 10)
        b = data
        a = _Item()
        a.SEQINDICATOR = (b[0] & 0xF0) >> 4; a.SUTYPE = b[0] & 0x0F
        a.QNO = (b[1] & 0xF0) >> 4; a.REFNO = b[1] & 0x07
        a.AESID = b[2] << 16 | b[3] << 8 | b[4]; a.GESID = b[5]
        idx = self._find(a)
        if idx < 0:
            self.isuitems.append(a.copy()); idx = len(self.isuitems) - 1
        p = self.isuitems[idx]
        p.count = 0
        total, index = {1: (1, 0), 2: (2, 0), 3: (2, 1), 4: (3, 0), 5: (3, 1), 6: (3, 2)}.get(a.SEQINDICATOR, (0, 0))
        nbytes = a.SUTYPE if 1 <= a.SUTYPE <= 11 else 0
        signalling = a.SUTYPE == 15
        thisnum = 11 * total - 11 + nbytes
        if thisnum > 0:
            if len(p.userdata) == 0:
                p.userdata = bytearray(thisnum)
            if thisnum < len(p.userdata):
                del p.userdata[thisnum:]
        if not signalling:
            for i in range(6, nbytes + 6):
                _put(p.userdata, i + 11 * index - 6, b[i])
            p.filledarray |= 1 << index
        else:
            p.userdata = bytearray()
        if signalling or (p.filledarray == 7 and total == 3) or (p.filledarray == 3 and total == 2) or (p.filledarray == 1 and total == 1):
            self.lastvalidisuitem = p.copy()
            del self.isuitems[idx]
            return True
        return False


class ISUData:
    """JAERO/aerol.cpp:116-214."""

    def __init__(self):
        self.isuitems = []; self.an = _Item(); self.lastvalidisuitem = _Item(); self.missingssu = False

    def reset(self):
        self.isuitems = []

    def update(self, data):
        self.missingssu = False
        d = bytes(data) + bytes(8)                          # reads past the 10 bytes yield 0 (QByteRef)
        m = d[0]
        an = self.an
        if m == 0x71:                                      # :158-182
            _age(self.isuitems, 10)
            an.AESID = d[1] << 16 | d[2] << 8 | d[3]; an.GESID = d[4]
            an.QNO = (d[5] >> 4) & 15; an.REFNO = d[5] & 15
            an.SEQNO = d[6] & 0x3F; an.NOOCT = (d[7] >> 4) & 15
            an.count = 0; an.userdata = bytearray(d[8:10])
            idx = -1
            if an.NOOCT <= 8:                              # findisuitem71 :116-124
                for i, it in enumerate(self.isuitems):
                    if an.AESID == it.AESID and an.GESID == it.GESID and an.QNO == it.QNO and an.REFNO == it.REFNO:
                        idx = i; break
            if idx < 0:
                self.isuitems.append(an.copy())
            else:
                self.isuitems[idx] = an.copy()
            return False
        if (m & 0xC0) != 0xC0:                             # :186
            return False
        an.SEQNO = m & 0x3F; an.QNO = (d[1] >> 4) & 15; an.REFNO = d[1] & 15
        idx = -1
        if an.NOOCT <= 8:                                  # findisuitemC0 :125-138 (AES/GES of the last 0x71 seen)
            for i, it in enumerate(self.isuitems):
                if an.AESID == it.AESID and an.GESID == it.GESID and an.SEQNO + 1 == it.SEQNO and an.QNO == it.QNO and an.REFNO == it.REFNO:
                    idx = i; break
        if idx < 0:
            self.missingssu = True
            return False
        p = self.isuitems[idx]
        p.SEQNO = (p.SEQNO - 1) & 0xFF
        if p.SEQNO == 0:
            p.userdata += d[2:p.NOOCT + 2]
            self.lastvalidisuitem = p.copy()
            return True
        p.userdata += d[2:10]
        return False


class _Acars:
    """ACARSItem (JAERO/aerol.h:176-211)."""

    def __init__(self):
        self.isuitem = _Item(); self.MODE = 0; self.TAK = 0; self.LABEL = b""; self.BI = 0; self.PLANEREG = b""
        self.nonacars = False; self.downlink = False; self.valid = False; self.hastext = False; self.moretocome = False
        self.message = b""; self.count = 0


class Parser:
    """ParserISU::parse (JAERO/aerol.cpp:340-487) + ACARSDefragmenter (:221-329); the database look-up answers empty, which
    leaves the removal of the registration's leading dots (:499-502)."""

    def __init__(self):
        self.frags = []; self.out = []

    def _emit(self, a):
        a.PLANEREG = a.PLANEREG.lstrip(b".")
        self.out.append(("acars", a))

    def _defragment(self, a):                              # :291-329
        _age(self.frags, 30)
        idx = -1
        for i, f in enumerate(self.frags):                 # findfragment :221-289
            if (a.PLANEREG == f.PLANEREG and a.LABEL == f.LABEL and a.MODE == f.MODE and a.isuitem.AESID == f.isuitem.AESID
                    and a.isuitem.GESID == f.isuitem.GESID and f.moretocome):
                if a.TAK != f.TAK:
                    continue
                t = f.BI + 1 - 65
                exp = ((t % 26 if t >= 0 else -((-t) % 26)) + 65) & 0xFF      # C remainder, then uchar
                if exp == a.BI:
                    idx = i; break
        if idx < 0:
            if not a.moretocome:
                return a
            a.count = 0; self.frags.append(a)
            return None
        f = self.frags[idx]
        f.count = 0; f.BI = a.BI; f.message += a.message; f.moretocome = a.moretocome
        if a.moretocome:
            return None
        del self.frags[idx]
        return f

    def parse(self, isu, downlink):
        if isu.AESID == 0:
            self.out.append(("error", b"Error: AESID == 0")); return False
        u = bytes(isu.userdata); n = len(u)
        odd = [bin(x).count("1") & 1 for x in u]
        a = _Acars(); a.downlink = bool(downlink); a.isuitem = isu.copy()
        if n > 16 and u[0] == 0xFF and u[1] == 0xFF and u[15] in (0x83, 0x02):
            a.MODE = u[3] & 0x7F; a.TAK = u[11] & 0x7F; a.LABEL = bytes([u[12] & 0x7F, u[13] & 0x7F]); a.BI = u[14] & 0x7F
            a.hastext = u[15] == 0x02
            a.moretocome = u[n - 4] == 0x97
            err = ("ISU: AESID = %X GESID = %X QNO = %02X REFNO = %02X : Parity error" % (isu.AESID, isu.GESID, isu.QNO, isu.REFNO)).encode()
            reg = bytearray()
            for k in range(4, 11):
                if not odd[k]:
                    self.out.append(("error", err)); return False
                reg.append(u[k] & 0x7F)
            a.PLANEREG = bytes(reg)
            msg = bytearray()
            if a.hastext:
                for k in range(16, n - 4):
                    if not odd[k]:
                        self.out.append(("error", err)); return False
                    c = u[k] & 0x7F
                    msg += b"<DEL>" if c == 0x7F else bytes([c])
            a.message = bytes(msg); a.valid = True
            done = self._defragment(a)
            if done is not None:
                self._emit(done)
            return True
        a.nonacars = True; a.valid = True
        a.message = u.hex().upper().encode()
        self._emit(a)
        return True


class Reassembly:
    """The three call sites of AeroL::Decode (JAERO/aerol.cpp:1357-1399 R, :1497-1513 T, :1900-1925 P)."""

    def __init__(self):
        self.isudata = ISUData(); self.risudata = RISUData(); self.parser = Parser()

    def reset(self):                                       # :992-993
        self.isudata.reset(); self.risudata.reset()

    def short_frame(self):                                 # :1997
        self.isudata.reset()

    def push_su(self, su, burstmode=False):
        m = su[0]
        if m == 0x71:
            self.isudata.update(su[:10]); return 0
        if (m & 0xC0) != 0xC0:
            return 0
        rc = 0
        if self.isudata.update(su[:10]):
            rc |= 1
            if self.parser.parse(self.isudata.lastvalidisuitem, burstmode):
                rc |= 4
        elif self.isudata.missingssu:
            rc |= 2
        return rc

    def push_r(self, info, burstmode=True):
        if (info[1] & 0x08) != 0x08:
            return 0
        rc = 0
        if self.risudata.update(bytes(info[:17])):
            rc |= 1
            if self.parser.parse(self.risudata.lastvalidisuitem, burstmode):
                rc |= 4
        return rc

    def pop_all(self):
        """records in the layout of oracle/ref.py reasm_record (hex strings)"""
        res = []
        for kind, x in self.parser.out:
            if kind == "error":
                res.append(dict(kind=1, message=x.hex()))
            else:
                i = x.isuitem
                res.append(dict(kind=0, aesid=i.AESID, gesid=i.GESID, qno=i.QNO, refno=i.REFNO, seqno=i.SEQNO, nooct=i.NOOCT,
                                mode=x.MODE, tak=x.TAK, bi=x.BI, nonacars=x.nonacars, downlink=x.downlink, valid=x.valid,
                                hastext=x.hastext, moretocome=x.moretocome, label=x.LABEL.hex(), reg=x.PLANEREG.hex(),
                                message=x.message.hex(), userdata=bytes(i.userdata).hex()))
        self.parser.out = []
        return res

"""Synthetic signal-unit streams for the ISU/SSU reassembly layer (test infrastructure).

Encoders are the inverse of what the reference decodes (JAERO/aerol.cpp:27-112 R-channel SUs, :151-214 P/T-channel
0x71 + SSUs, :340-487 ACARS framing inside the ISU user data). Stream events:
  ("su", 10 bytes, burstmode)   one CRC-valid P/T-channel signal unit
  ("r", 17 bytes, burstmode)    one CRC-valid R-channel packet
  ("reset",) / ("short",)       AeroL::setSettings / the short-frame reset (JAERO/aerol.cpp:992-993, 1997)"""
import numpy as np


def odd_parity(b):
    b &= 0x7F
    return b | (0x80 if bin(b).count("1") % 2 == 0 else 0)


def acars_userdata(reg, label, bi, text=None, more=False, mode="2", tak=0x15, bad_parity_at=None):
    """FF FF SOH mode reg*7 TAK label*2 BI STX text ETX/ETB BSC*2 DEL (JAERO/aerol.cpp:363-375)."""
    ud = [0xFF, 0xFF, 0x01, odd_parity(ord(mode))]
    ud += [odd_parity(ord(c)) for c in reg.rjust(7, ".")[:7]]
    ud += [odd_parity(tak), odd_parity(ord(label[0])), odd_parity(ord(label[1])), odd_parity(bi)]
    if text is None:
        ud += [0x83]
    else:
        ud += [0x02] + [odd_parity(ord(c)) for c in text]
        ud += [0x97 if more else 0x83]
    ud += [0x93, 0xAB, 0x7F]
    if bad_parity_at is not None:
        ud[bad_parity_at] ^= 0x80
    return bytes(ud)


def isu_to_sus(aes, ges, qno, refno, userdata):
    """0x71 initial SU + SSUs (JAERO/aerol.cpp:159-211): 2 bytes in the initial SU, 8 per SSU, 1..8 in the last."""
    ud = bytes(userdata)
    assert len(ud) >= 3
    rest = ud[2:]
    n_ssu = (len(rest) + 7) // 8
    last = len(rest) - 8 * (n_ssu - 1)
    assert 1 <= n_ssu <= 63
    sus = [bytes([0x71, (aes >> 16) & 255, (aes >> 8) & 255, aes & 255, ges, (qno << 4) | refno, n_ssu, last << 4, ud[0], ud[1]])]
    for k in range(n_ssu):
        chunk = rest[8 * k:8 * k + 8].ljust(8, b"\0")
        sus.append(bytes([0xC0 | (n_ssu - 1 - k), (qno << 4) | refno]) + chunk)
    return sus


def risu_to_packets(aes, ges, qno, refno, userdata):
    """R-channel user-data SUs (JAERO/aerol.cpp:27-112): up to 3 SUs of 11 bytes, the SU type of each = bytes it carries."""
    ud = bytes(userdata)
    total = (len(ud) + 10) // 11
    assert 1 <= total <= 3
    seq0 = {1: 1, 2: 2, 3: 4}[total]
    out = []
    for k in range(total):
        chunk = ud[11 * k:11 * k + 11]
        b = bytes([((seq0 + k) << 4) | len(chunk), (qno << 4) | 0x08 | (refno & 7), (aes >> 16) & 255, (aes >> 8) & 255, aes & 255, ges])
        out.append(b + chunk.ljust(11, b"\0"))
    return out


def _text(rng, n):
    alphabet = "ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 /.,-\r\n"
    return "".join(alphabet[i] for i in rng.integers(0, len(alphabet), n))


def p_stream(seed, burst=False, n_msgs=60):
    rng = np.random.default_rng(seed)
    ev = []
    fleet = [(0x400000 + int(rng.integers(1, 1 << 20)), "G-%s" % "".join("ABCDEFGHJK"[i] for i in rng.integers(0, 10, 4))) for _ in range(6)]
    bi = {a: ord("A") + int(rng.integers(0, 26)) for a, _ in fleet}

    def nextbi(a):
        bi[a] = (bi[a] + 1 - ord("A")) % 26 + ord("A")
        return bi[a]

    for m in range(n_msgs):
        aes, reg = fleet[int(rng.integers(0, len(fleet)))]
        ges = 0x90 + (aes & 3)
        qno, refno = int(rng.integers(0, 16)), int(rng.integers(0, 16))
        label = "".join("HQ15AB_"[i] for i in rng.integers(0, 7, 2))
        mode = rng.integers(0, 12)
        if mode == 0:      # no text
            uds = [acars_userdata(reg, label, nextbi(aes))]
        elif mode == 1:    # three-block message (ETB, ETB, ETX)
            uds = [acars_userdata(reg, label, nextbi(aes), _text(rng, 200), more=True),
                   acars_userdata(reg, label, nextbi(aes), _text(rng, 180), more=True),
                   acars_userdata(reg, label, nextbi(aes), _text(rng, 33))]
        elif mode == 2:    # parity error inside the text
            uds = [acars_userdata(reg, label, nextbi(aes), _text(rng, 40), bad_parity_at=20)]
        elif mode == 3:    # parity error inside the registration
            uds = [acars_userdata(reg, label, nextbi(aes), _text(rng, 40), bad_parity_at=6)]
        elif mode == 4:    # not ACARS: reported as hex
            uds = [bytes(rng.integers(0, 256, int(rng.integers(3, 60)), dtype=np.uint8))]
        elif mode == 5:    # second block with a different TAK: no defragmentation match
            uds = [acars_userdata(reg, label, nextbi(aes), _text(rng, 50), more=True),
                   acars_userdata(reg, label, nextbi(aes), _text(rng, 20), tak=0x06)]
        elif mode == 6:    # block index does not follow on
            b0 = nextbi(aes); nextbi(aes)
            uds = [acars_userdata(reg, label, b0, _text(rng, 50), more=True), acars_userdata(reg, label, nextbi(aes), _text(rng, 20))]
        elif mode == 7:    # DEL inside the text
            uds = [acars_userdata(reg, label, nextbi(aes), "AB\x7fCD")]
        else:
            uds = [acars_userdata(reg, label, nextbi(aes), _text(rng, int(rng.integers(1, 221))))]
        for ud in uds:
            sus = isu_to_sus(0 if (mode == 8 and m % 5 == 0) else aes, ges, qno, refno, ud)
            r = rng.integers(0, 10)
            if r == 0 and len(sus) > 3:      # lose one SSU
                del sus[int(rng.integers(1, len(sus)))]
            elif r == 1 and len(sus) > 3:    # duplicate one SSU
                k = int(rng.integers(1, len(sus))); sus.insert(k, sus[k])
            elif r == 2:                     # another aircraft's initial SU in the middle (SSUs follow the LAST 0x71 seen)
                other = isu_to_sus(fleet[0][0] ^ 1, 0x91, 3, 4, acars_userdata("N123AB", "H1", ord("C"), "X" * 10))
                sus.insert(min(2, len(sus)), other[0])
            for s in sus:
                ev.append(("su", s, burst))
            for _ in range(int(rng.integers(0, 4))):   # other SU types in between are not passed on
                ev.append(("su", bytes([int(rng.choice([0x01, 0x20, 0x40, 0x61]))]) + bytes(rng.integers(0, 256, 9, dtype=np.uint8)), burst))
        if m == n_msgs // 2:
            ev.append(("short",))
        if m == 3 * n_msgs // 4:
            ev.append(("reset",))
    return ev


def r_stream(seed, n_msgs=80):
    rng = np.random.default_rng(seed)
    ev = []
    for m in range(n_msgs):
        aes = 0x500000 + int(rng.integers(1, 1 << 16)); ges = 0xC1
        qno, refno = int(rng.integers(0, 16)), int(rng.integers(0, 8))
        mode = rng.integers(0, 8)
        if mode == 0:
            ud = acars_userdata("VH-OQA", "Q0", ord("A") + m % 26)            # 19 bytes: 2 SUs
        elif mode == 1:
            ud = acars_userdata("VH-OQA", "5Z", ord("A") + m % 26, _text(rng, int(rng.integers(1, 13))))   # <= 33 bytes
        else:
            ud = bytes(rng.integers(0, 256, int(rng.integers(1, 34)), dtype=np.uint8))
        pk = risu_to_packets(aes, ges, qno, refno, ud)
        order = list(range(len(pk)))
        if mode == 2:
            order.reverse()
        if mode == 3 and len(pk) > 1:
            order = order[:-1]                                                   # never completes
        for k in order:
            ev.append(("r", pk[k], True))
        if mode == 4:                                                            # signalling SU type 15 / reserved types
            ev.append(("r", bytes([0x1F, (qno << 4) | 0x08 | refno]) + bytes(rng.integers(0, 256, 15, dtype=np.uint8)), True))
            ev.append(("r", bytes([0x1C, (qno << 4) | 0x08 | refno]) + bytes(rng.integers(0, 256, 15, dtype=np.uint8)), True))
            ev.append(("r", bytes([0x10, (qno << 4) | 0x08 | refno]) + bytes(rng.integers(0, 256, 15, dtype=np.uint8)), True))
        if mode == 5:                                                            # not a user-data SU: ignored
            ev.append(("r", bytes([0x11, (qno << 4) | refno, 0x22]) + bytes(rng.integers(0, 256, 14, dtype=np.uint8)), True))
    return ev


def garbage_stream(seed, n=3000):
    """Dense random SUs: every byte pattern must be handled exactly as the reference handles it."""
    rng = np.random.default_rng(seed)
    ev = []
    for i in range(n):
        k = rng.integers(0, 10)
        if k < 3:
            b = bytearray(rng.integers(0, 256, 10, dtype=np.uint8)); b[0] = 0x71
            b[1] = 0; b[2] = 0; b[3] = int(rng.integers(0, 3)); b[4] = int(rng.integers(0, 2)); b[5] &= 0x11; b[6] &= 0x03
            ev.append(("su", bytes(b), bool(i & 1)))
        elif k < 8:
            b = bytearray(rng.integers(0, 256, 10, dtype=np.uint8)); b[0] = 0xC0 | int(rng.integers(0, 3)); b[1] &= 0x11
            ev.append(("su", bytes(b), bool(i & 1)))
        else:
            b = bytearray(rng.integers(0, 256, 17, dtype=np.uint8))
            b[1] = (b[1] & 0x11) | 0x08; b[2] = 0; b[3] = 0; b[4] = int(rng.integers(0, 2)); b[5] = int(rng.integers(0, 2))
            ev.append(("r", bytes(b), True))
    return ev


def synthetic_streams():
    return {
        "p_synthetic": p_stream(11),
        "t_synthetic": p_stream(12, burst=True, n_msgs=40),
        "r_synthetic": r_stream(13),
        "garbage": garbage_stream(14),
    }


_KIND = {"su": 0, "r": 1, "reset": 2, "short": 3}


def pack_stream(ev):
    a = np.zeros((len(ev), 20), dtype=np.uint8)
    for i, e in enumerate(ev):
        a[i, 0] = _KIND[e[0]]
        if e[0] in ("su", "r"):
            a[i, 1] = int(e[2])
            p = np.frombuffer(bytes(e[1]), dtype=np.uint8)
            a[i, 2:2 + len(p)] = p
    return a


def unpack_stream(a):
    ev = []
    for row in a:
        k = int(row[0])
        if k == 0:
            ev.append(("su", bytes(row[2:12]), bool(row[1])))
        elif k == 1:
            ev.append(("r", bytes(row[2:19]), bool(row[1])))
        elif k == 2:
            ev.append(("reset",))
        else:
            ev.append(("short",))
    return ev

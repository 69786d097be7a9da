"""First-contact GPU check (development aid, not a test): CUDA path vs the CPU oracle with verbose diffs."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jaero_b200  # noqa: E402
from oracle import restated  # noqa: E402


def check_viterbi():
    rng = np.random.default_rng(1)
    C, n = 8, 4992
    vb = jaero_b200.ViterbiBatch(C, 24)
    orc = [restated.OracleViterbi(24) for _ in range(C)]
    ok = True
    for it in range(3):
        soft = np.zeros((C, n), dtype=np.uint8)
        for c in range(C):
            msg = rng.integers(0, 256, size=n // 16, dtype=np.uint8)
            enc = restated.conv_encode(msg)[:n].astype(np.float64)
            noisy = (enc * 2 - 1) * 60 + 128 + rng.normal(0, 45 + 10 * c, size=n)
            soft[c] = np.clip(np.round(noisy), 0, 255).astype(np.uint8)
        # interleave so that cols=78 de-interleaves back to `soft`
        inter = np.zeros_like(soft)
        for c in range(C):
            k = np.arange(n); i = k % 64; j = k // 64
            inter[c, ((i * 27) % 64) * 78 + j] = soft[c, k]
        t = time.time()
        got = vb.decode_continuous(inter, 78)
        dt = time.time() - t
        for c in range(C):
            ref = orc[c].decode_continuous(soft[c])
            same = np.array_equal(ref.astype(np.uint8), got[c][:len(ref)]) and vb.last_valid[c] == len(ref)
            ok &= same
            if not same:
                bad = np.nonzero(ref.astype(np.uint8) != got[c][:len(ref)])[0]
                print("  viterbi it", it, "ch", c, "mismatch at", bad[:10], "count", len(bad))
        print("viterbi iteration", it, "ok" if ok else "FAIL", "%.1f ms" % (dt * 1e3))
    blk = rng.integers(0, 256, size=(C, 600), dtype=np.uint8)
    got = vb.decode_block(blk)
    for c in range(C):
        ref = restated.conv_decode_soft(blk[c])[:294]
        if not np.array_equal(ref, got[c][:294]):
            ok = False; print("  decode_block mismatch ch", c)
    print("VITERBI", "PASS" if ok else "FAIL")
    return ok


def check_demod(kind, name, kw, nch=3, seconds=None):
    pcm = np.load(os.path.join(ROOT, "tests", "golden", name + "_excerpt.npz"))["pcm"]
    if seconds:
        pcm = pcm[:int(seconds * 48000)]
    rng = np.random.default_rng(7)
    chans = [pcm]
    for c in range(1, nch):
        x = pcm.astype(np.float64) * (0.5 + 0.4 * c / nch) + rng.normal(0, 150 * c, size=len(pcm))
        chans.append(np.clip(np.round(x), -32768, 32767).astype(np.int16))
    pcm2 = np.stack(chans)
    b = jaero_b200.DemodBatch(kind, nch, device=0, report_ebno=True, **kw)
    orc = [restated.OracleDemod(kind, **kw) for _ in range(nch)]
    chunk = 4800
    t = time.time()
    acc = [[] for _ in range(nch)]
    for a in range(0, pcm2.shape[1], chunk):
        b.write(pcm2[:, a:a + chunk])
        if (a // chunk) % 8 == 7:
            for c, s_ in enumerate(b.read_softbits()):
                acc[c].append(s_)
    b.sync()
    tg = time.time() - t
    for c, s_ in enumerate(b.read_softbits()):
        acc[c].append(s_)
    soft = [np.concatenate(x) for x in acc]
    st = b.status()
    ok = True
    for c in range(nch):
        t = time.time()
        for a in range(0, pcm2.shape[1], chunk):
            orc[c].write(pcm2[c, a:a + chunk])
        tc = time.time() - t
        s_ref = orc[c].take_soft(); o = orc[c].state()
        n = min(len(s_ref), len(soft[c]))
        hard_same = np.array_equal(s_ref[:n] >= 128, soft[c][:n] >= 128) and len(s_ref) == len(soft[c])
        exact = np.array_equal(s_ref, soft[c])
        maxd = int(np.abs(s_ref[:n].astype(int) - soft[c][:n].astype(int)).max()) if n else -1
        rel = {k: abs(st[c][k] - o[k]) / max(abs(o[k]), 1e-12) for k in o if k in st[c]}
        worst = max(rel, key=rel.get)
        print(f"{kind} ch{c}: soft n={len(soft[c])}/{len(s_ref)} exact={exact} hard_same={hard_same} max|d|={maxd} "
              f"worst state {worst} rel={rel[worst]:.3e}  mse={st[c]['mse']:.5f}/{o['mse']:.5f} f={st[c]['mixer2_freq']:.4f}/{o['mixer2_freq']:.4f} "
              f"ebno={st[c]['ebno']:.3f}/{o['ebno']:.3f} cpu={tc:.2f}s")
        ok &= hard_same and rel[worst] < 1e-4
    print(kind.upper(), "PASS" if ok else "FAIL", "gpu wall %.2fs for %d ch x %d samples, launches=%d" % (tg, nch, pcm2.shape[1], b.launches))
    b.close()
    return ok


def check_pchannel():
    """demod -> device framing/Viterbi/CRC vs oracle P-channel on the same soft bits, incl. DCD feedback."""
    from jaero_b200 import synth
    ok = True
    cases = [("oqpsk", 10500, np.load(os.path.join(ROOT, "tests", "golden", "oqpsk_10500_excerpt.npz"))["pcm"],
              dict(fb=10500, freq_center=5760, lockingbw=10500, afc=True)),
             ("oqpsk", 10500, synth.oqpsk_pchannel_pcm(10, fc=8000.0, seed=5, ebn0_db=9.0), dict(fb=10500, freq_center=8000, lockingbw=10500)),
             ("msk", 600, np.load(os.path.join(ROOT, "tests", "golden", "msk_600_excerpt.npz"))["pcm"],
              dict(fb=600, freq_center=1000, lockingbw=900, afc=True))]
    for kind, fb, pcm, kw in cases:
        b = jaero_b200.DemodBatch(kind, 2, **kw)
        pc = jaero_b200.PChannelBatch(2, fb)
        od = [restated.OracleDemod(kind, fft_power=14 if kind == "oqpsk" else 13, signalthreshold=0.65 if kind == "oqpsk" else 0.5, **kw) for _ in range(2)]
        op = [restated.OraclePChannel(fb) for _ in range(2)]
        pcm2 = np.stack([pcm, (pcm.astype(np.int32) * 3 // 4).astype(np.int16)])
        chunk = 4096
        got = [[], []]
        for a in range(0, pcm2.shape[1], chunk):
            b.write(pcm2[:, a:a + chunk])
            pc.process_batch(b)
            for c in range(2):
                od[c].set_dcd(op[c].dcd)          # the DCD the GPU pipeline sees: state at the end of the previous chunk
                od[c].write(pcm2[c, a:a + chunk])
                op[c].process(od[c].take_soft())
            if (a // chunk) % 11 == 10:
                for c, r in enumerate(pc.read_sus()):
                    got[c].append(r)
            if (a + chunk) % 48000 < chunk:       # ~1 s tick on both sides at the same chunk boundary
                pc.tick(b)
                for c in range(2):
                    op[c].update_dcd()
        for c, r in enumerate(pc.read_sus()):
            got[c].append(r)
        dcd, tot, okc = pc.stats()
        for c in range(2):
            gb = np.concatenate([g[0] for g in got[c]]); gok = np.concatenate([g[1] for g in got[c]])
            rb, rok, rfr = op[c].take_sus()
            same = gb.shape == rb.shape and np.array_equal(gb, rb) and np.array_equal(gok, rok)
            st = b.status()[c]; o = od[c].state()
            print(f"pchannel {kind}{fb} ch{c}: SUs gpu={len(gok)} oracle={len(rok)} crc_ok gpu={int(gok.sum())} oracle={int(rok.sum())} identical={same} "
                  f"dcd gpu={dcd[c]} oracle={int(op[c].dcd)} f={st['mixer2_freq']:.4f}/{o['mixer2_freq']:.4f}")
            ok &= same and dcd[c] == int(op[c].dcd)
        b.close(); pc.close()
    print("PCHANNEL", "PASS" if ok else "FAIL")
    return ok


def check_burst(names=("burst_msk_1200_a", "burst_msk_1200_b")):
    ok = True
    for name in names:
        pcm = np.load(os.path.join(ROOT, "tests", "golden", name + "_excerpt.npz"))["pcm"]
        pcm2 = np.stack([pcm, (pcm.astype(np.int32) * 3 // 5).astype(np.int16), np.roll(pcm, 12345)])
        oq = name.startswith("burst_oqpsk")
        okind = "burst_oqpsk" if oq else "burst_msk"
        kw = dict(fb=10500.0, freq_center=8000.0, lockingbw=10500.0, signalthreshold=0.6) if oq else dict(fb=1200.0, freq_center=1000.0, lockingbw=1800.0, signalthreshold=0.6)
        b = (jaero_b200.BurstOqpskBatch if oq else jaero_b200.BurstMskBatch)(3, **kw)
        acc = [[] for _ in range(3)]
        t = time.time()
        for a in range(0, pcm2.shape[1], 4800):
            b.write(pcm2[:, a:a + 4800])
            for c, s_ in enumerate(b.read_softbits()):
                acc[c].append(s_)
        tg = time.time() - t
        st = b.status()
        for c in range(3):
            o = restated.OracleDemod(okind, **kw)
            for a in range(0, pcm2.shape[1], 4800):
                o.write(pcm2[c, a:a + 4800])
            so = o.take_soft(); sg = np.concatenate(acc[c]); os_ = o.state()
            n = min(len(so), len(sg))
            same_len = len(so) == len(sg)
            hard = same_len and np.array_equal(so >= 128, sg >= 128) and np.array_equal(so < 0, sg < 0)
            maxd = int(np.abs(so[:n].astype(int) - sg[:n].astype(int)).max()) if n else -1
            keys = ["mixer2_freq", "vol_gain", "mse", "n_sig_true", "n_sig_false", "cntr", "startstop", "agc", "st_wtptr", "rotator_freq"]
            rel = {k: abs(st[c][k] - os_[k]) / max(abs(os_[k]), 1e-9) for k in keys}
            worst = max(rel, key=rel.get)
            eb = o.take_aux(0)
            print(f"burst {name} ch{c}: soft {len(sg)}/{len(so)} markers {int((sg<0).sum())}/{int((so<0).sum())} hard_same={hard} max|d|={maxd} worst {worst} rel={rel[worst]:.2e} "
                  f"f={st[c]['mixer2_freq']:.3f}/{os_['mixer2_freq']:.3f} gain={st[c]['vol_gain']:.6f}/{os_['vol_gain']:.6f} ebno_last={st[c]['last_burst_ebno']:.4f}/{eb[-1] if len(eb) else float('nan'):.4f} emits={st[c]['n_ebno_emits']:.0f}/{len(eb)}")
            ok &= hard and rel[worst] < 1e-6
        print("gpu wall %.2fs launches=%d" % (tg, b.launches))
        b.close()
    print("BURST", "PASS" if ok else "FAIL")
    return ok


def check_cchannel():
    """C-channel (8400 bps): GPU demodulator -> GPU frame layer (DCD fed back on the device) vs the restated oracle chain."""
    pcm = np.load(os.path.join(ROOT, "tests", "golden", "oqpsk_8400_excerpt.npz"))["pcm"]
    pcm2 = np.stack([pcm, (pcm.astype(np.int32) * 2 // 3).astype(np.int16)])
    kw = dict(fb=8400, freq_center=8000, lockingbw=10500, fft_power=14, signalthreshold=0.65, afc=True)
    b = jaero_b200.DemodBatch("oqpsk", 2, **kw)
    cc = jaero_b200.CChannelBatch(2)
    got = [[], []]
    for k, a in enumerate(range(0, pcm2.shape[1], 4800)):
        b.write(pcm2[:, a:a + 4800])
        cc.process_batch(b)
        if k % 10 == 9:
            cc.tick(b)
        for c, fr in enumerate(cc.read_frames()):
            got[c].append(fr)
    dcd, tot, okc = cc.stats()
    ok = True
    for c in range(2):
        o = restated.OracleDemod("oqpsk", **kw); oc = restated.OracleCChannel()
        for k, a in enumerate(range(0, pcm2.shape[1], 4800)):
            o.write(pcm2[c, a:a + 4800])
            oc.process(o.take_soft())
            o.set_dcd(int(oc.dcd))
            if k % 10 == 9:
                oc.update_dcd(); o.set_dcd(int(oc.dcd))
        su, cok, voice = oc.take_frames()
        gsu = np.concatenate([g[0] for g in got[c]]); gok = np.concatenate([g[1] for g in got[c]]); gv = np.concatenate([g[2] for g in got[c]])
        same = gsu.shape == su.shape and np.array_equal(gsu, su) and np.array_equal(gok, cok) and np.array_equal(gv, voice)
        print(f"cchannel ch{c}: frames gpu={len(gsu)} oracle={len(su)} SU crc_ok gpu={int(gok.sum())} oracle={int(cok.sum())} identical={same} dcd gpu={dcd[c]} oracle={int(oc.dcd)}")
        ok &= same and dcd[c] == int(oc.dcd) and (c != 0 or (len(su) > 5 and int(cok.sum()) > 10))
    b.close(); cc.close()
    print("CCHANNEL", "PASS" if ok else "FAIL")
    return ok


def check_rt():
    """R/T burst channel layer: GPU burst demodulator -> GPU packet decoder vs the restated oracle chain."""
    ok = True
    for name in ("burst_msk_1200_a", "burst_msk_1200_b", "burst_oqpsk_10500"):
        pcm = np.load(os.path.join(ROOT, "tests", "golden", name + "_excerpt.npz"))["pcm"]
        pcm2 = np.stack([pcm, np.roll(pcm, 7001)])
        oq = name.startswith("burst_oqpsk")
        fb = 10500 if oq else 1200
        kw = dict(fb=10500.0, freq_center=8000.0, lockingbw=10500.0, signalthreshold=0.6) if oq else dict(fb=1200.0, freq_center=1000.0, lockingbw=1800.0, signalthreshold=0.6)
        b = (jaero_b200.BurstOqpskBatch if oq else jaero_b200.BurstMskBatch)(2, **kw)
        rt = jaero_b200.RTChannelBatch(2, fb)
        got = [[], []]
        for k, a in enumerate(range(0, pcm2.shape[1], 4800)):
            b.write(pcm2[:, a:a + 4800])
            rt.process_burst(b)
            if k % 10 == 9:
                rt.tick()
            for c, pk in enumerate(rt.read_packets()):
                got[c] += pk
        tr, bad, dcd = rt.stats()
        for c in range(2):
            o = restated.OracleDemod("burst_oqpsk" if oq else "burst_msk", **kw)
            ort = restated.OracleRTChannel(fb)
            for k, a in enumerate(range(0, pcm2.shape[1], 4800)):
                o.write(pcm2[c, a:a + 4800])
                ort.process(o.take_soft())
                if k % 10 == 9:
                    ort.update_dcd()
            ref = ort.packets()
            same = len(ref) == len(got[c]) and all(r["type"] == g["type"] and r["nsus"] == g["nsus"] and np.array_equal(r["bytes"], g["bytes"]) for r, g in zip(ref, got[c]))
            print(f"rt {name} ch{c}: packets gpu={len(got[c])} oracle={len(ref)} identical={same} trials gpu={tr[c]} oracle={ort.trials} "
                  f"types={[g['type'] for g in got[c]]} nsus={[g['nsus'] for g in got[c]]}")
            ok &= same and tr[c] == ort.trials and len(ref) > 0
        b.close(); rt.close()
    print("RT", "PASS" if ok else "FAIL")
    return ok


if __name__ == "__main__":
    which = sys.argv[1:] or ["viterbi", "oqpsk", "msk", "pchannel"]
    ok = True
    if "viterbi" in which:
        ok &= check_viterbi()
    if "oqpsk" in which:
        ok &= check_demod("oqpsk", "oqpsk_10500", dict(fb=10500, freq_center=5760, lockingbw=10500, fft_power=14, signalthreshold=0.65, afc=True))
    if "cchannel" in which:
        ok &= check_cchannel()
    if "rt" in which:
        ok &= check_rt()
    if "burst_oqpsk" in which:
        ok &= check_burst(("burst_oqpsk_10500",))
    if "oqpsk8400" in which:
        ok &= check_demod("oqpsk", "oqpsk_8400", dict(fb=8400, freq_center=8000, lockingbw=10500, fft_power=14, signalthreshold=0.65, afc=True))
    if "msk" in which:
        ok &= check_demod("msk", "msk_600", dict(fb=600, freq_center=1000, lockingbw=900, fft_power=13, signalthreshold=0.5, afc=True))
    if "burst" in which:
        ok &= check_burst()
    if "pchannel" in which:
        ok &= check_pchannel()
    print("ALL", "PASS" if ok else "FAIL")
    sys.exit(0 if ok else 1)

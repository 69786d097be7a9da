"""CPU checks of the arithmetic identities the round-2 kernels lean on (no GPU needed): each is a property of the reference's
algorithm that the CUDA code uses to do less work, restated here in numpy / C so that a change of the reference constants
(polynomials, unique word, frame sizes) would be caught before it silently breaks a kernel."""
import os
import subprocess

import numpy as np

from conftest import ROOT

POLYS = (109, 79)                      # JAERO/jconvolutionalcodec.cpp:13-14
UWORD = 0xE15AE893                     # JAERO/aerol.cpp:947


def _table(sr):
    return (bin(sr & POLYS[0]).count("1") & 1) | ((bin(sr & POLYS[1]).count("1") & 1) << 1)


def test_viterbi_branch_labels_are_one_label_and_its_complement():
    """viterbi_core.cuh: lane l needs table[2l], table[2l | 64], table[2l + 1], table[(2l + 1) | 64]; both polynomials have their
    first and last taps set, so these are t, t ^ 3, t ^ 3, t — and with soft values a, b the four distances are d and 510 - d."""
    for lane in range(32):
        s0 = 2 * lane
        t = _table(s0)
        assert _table(s0 | 64) == t ^ 3 and _table(s0 | 1) == t ^ 3 and _table(s0 | 1 | 64) == t
    rng = np.random.default_rng(1)
    for a0, b0 in rng.integers(0, 256, size=(200, 2)):
        a1, b1 = 255 - a0, 255 - b0
        for t in range(4):
            d = (a1 if t & 1 else a0) + (b1 if t & 2 else b0)
            x = (int(a0) | (int(b0) << 8)) ^ ((0x00ff if t & 1 else 0) | (0xff00 if t & 2 else 0))
            assert (x & 0xff) + (x >> 8) == d                       # the kernel's one-XOR form of the distance
            tc = t ^ 3
            assert (a1 if tc & 1 else a0) + (b1 if tc & 2 else b0) == 510 - d


def _even_bits(x):
    x &= 0x55555555
    x = (x | (x >> 1)) & 0x33333333
    x = (x | (x >> 2)) & 0x0f0f0f0f
    x = (x | (x >> 4)) & 0x00ff00ff
    x = (x | (x >> 8)) & 0x0000ffff
    return x


def _brev(x):
    return int("{:032b}".format(x & 0xffffffff)[::-1], 2)


def test_frame_kernel_ballot_windows_equal_the_serial_shift_registers():
    """pchannel.cu fast path: the unique-word registers after each of 32 interleaved real / imaginary bits, formed from one
    ballot word, equal the reference's bit-serial update of the two shift registers (aerol.cpp:781-804, :1156-1233)."""
    rng = np.random.default_rng(2)
    for trial in range(300):
        r = int(rng.integers(1, 33))
        bits = rng.integers(0, 2, size=32)
        sr = {"imag": int(rng.integers(0, 2 ** 32)), "real": int(rng.integers(0, 2 ** 32))}
        realimag = int(rng.integers(0, 2))
        # serial
        s_im, s_re, ri = sr["imag"], sr["real"], realimag
        serial = []
        for j in range(r):
            ri = (ri + 1) % 2
            if ri:
                s_im = ((s_im << 1) | int(bits[j])) & 0xffffffff; serial.append(s_im)
            else:
                s_re = ((s_re << 1) | int(bits[j])) & 0xffffffff; serial.append(s_re)
        # ballot form
        valid = 0xffffffff if r == 32 else (1 << r) - 1
        B = sum(int(bits[j]) << j for j in range(32)) & valid
        a_imag = 0 if ((realimag + 1) & 1) else 1
        E0, E1 = _even_bits(B), _even_bits(B >> 1)
        for lane in range(r):
            mine_imag = (lane & 1) == a_imag
            Em = E1 if lane & 1 else E0
            srp = sr["imag"] if mine_imag else sr["real"]
            k = lane >> 1
            W = ((srp << (k + 1)) | (_brev(Em) >> (31 - k))) & 0xffffffff
            assert W == serial[lane], (trial, lane)
        n0, n1 = (r + 1) >> 1, r >> 1
        imag_E, real_E = (E1, E0) if a_imag else (E0, E1)
        imag_n, real_n = (n1, n0) if a_imag else (n0, n1)
        f_im = ((sr["imag"] << imag_n) | (_brev(imag_E) >> (32 - imag_n))) & 0xffffffff if imag_n else sr["imag"]
        f_re = ((sr["real"] << real_n) | (_brev(real_E) >> (32 - real_n))) & 0xffffffff if real_n else sr["real"]
        assert (f_im, f_re) == (s_im, s_re), trial


def test_su_kernel_ballot_bytes_equal_the_lsb_first_packing():
    """pchannel.cu SU kernel: byte t of a round = bits 8t..8t+7 of the ballot word == the reference's `ch8 |= b*128; ch8 >>= 1`
    packing (aerol.cpp:1568-1580)."""
    rng = np.random.default_rng(3)
    for _ in range(100):
        b = rng.integers(0, 2, size=32)
        word = sum(int(b[k]) << k for k in range(32))
        ch8, charptr, out = 0, 0, []
        for k in range(32):
            ch8 |= int(b[k]) * 128
            charptr = (charptr + 1) % 8
            if charptr == 0:
                out.append(ch8 & 0xff); ch8 = 0
            else:
                ch8 >>= 1
        assert out == [(word >> (8 * t)) & 0xff for t in range(4)]


def test_estimator_log10_host_twin_is_within_two_ulp_of_libm(tmp_path):
    """cfe.cu log10_ge1: the host twin under tools/micro (same constants and operation order, `/` for div_fast)."""
    exe = str(tmp_path / "log10_test")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tools", "micro", "log10_test.c"), "-lm"], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    line = [l for l in out.splitlines() if l.startswith("max ulp")][0]
    assert float(line.split()[2]) <= 2.0, out
    src = open(os.path.join(ROOT, "jaero_b200", "csrc", "cfe.cu")).read()
    twin = open(os.path.join(ROOT, "tools", "micro", "log10_test.c")).read()
    for const in ("6.666666666666735130e-01", "1.479819860511658591e-01", "4.34294481903251816668e-01", "3.69423907715893078616e-13", "0x95f64"):
        assert const in src and const in twin, const


def test_idle_block_index_never_completes_a_block():
    """pchannel.cu: while the frame counter idles at 1e9 every soft bit lands on one constant block position; the fast path (and
    the removal of the per-block copy-forward) rely on that position never being the last one of a block."""
    for bits_in_header, block_len in ((16, 64 * 6), (16, 64 * 9), (16 + 178, 64 * 78)):
        assert (1000000000 - bits_in_header) % block_len != block_len - 1

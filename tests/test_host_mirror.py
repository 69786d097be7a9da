"""C++ host mirror of the reference's demodulator classes (include/jaero_b200_host.hpp) over the C ABI.

CPU: the mirror's plumbing against a mock of the C ABI (tests/cpp/mock_capi.cpp), and the loud-failure path against the
real library on a machine without a GPU. GPU: the classes driven like the reference's own (setAFC / setSettings /
writeData in chunks / processDemodulatedSoftBits) on the reference recordings must give exactly the soft bits of the
batch API, whose parity with the oracle the other GPU tests establish."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_cuda, load_excerpt

INC = os.path.join(ROOT, "include")
CPP = os.path.join(ROOT, "tests", "cpp")
LIBDIR = os.path.join(ROOT, "jaero_b200")


def _build(tmp_path, name, sources, mock):
    exe = str(tmp_path / name)
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-I" + INC] + (["-DMOCK"] if mock else []) + [os.path.join(CPP, s) for s in sources]
    if not mock:
        cmd += ["-L" + LIBDIR, "-ljaero_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd + ["-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_host_mirror_plumbing_against_mock_abi(tmp_path):
    exe = _build(tmp_path, "hm_mock", ["host_mirror_test.cpp", "mock_capi.cpp"], mock=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "MOCK OK" in r.stdout, r.stdout + r.stderr


def test_host_mirror_fails_loudly_without_a_device(tmp_path):
    exe = _build(tmp_path, "hm_real", ["host_mirror_test.cpp"], mock=False)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ("REAL-NO-GPU OK" in r.stdout) or ("SKIP" in r.stdout)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("name", ["oqpsk_10500", "msk_600", "burst_msk_1200_a", "burst_oqpsk_10500"])
def test_host_mirror_classes_on_reference_recordings(tmp_path, golden, name):
    import jaero_b200 as jb
    case = golden[name]
    kind, kw = case["kind"], dict(case["kw"])
    pcm = load_excerpt(case["excerpt"])
    exe = _build(tmp_path, "host_demod_run", ["host_demod_run.cpp"], mock=False)
    raw, out = str(tmp_path / "pcm.raw"), str(tmp_path / "soft.raw")
    pcm.astype("<i2").tofile(raw)
    afc = int(kw.get("afc", True))
    r = subprocess.run([exe, kind, raw, out, str(case["chunk"]), str(kw["fb"]), str(kw["freq_center"]), str(kw["lockingbw"]),
                        str(kw["signalthreshold"]), str(afc)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    got = np.fromfile(out, dtype="<i2")
    # the same stream through the batch API (one channel)
    if kind == "burst_msk":
        b = jb.BurstMskBatch(1, **kw)
    elif kind == "burst_oqpsk":
        b = jb.BurstOqpskBatch(1, **kw)
    else:
        b = jb.DemodBatch(kind, 1, **kw)
    acc = []
    for a in range(0, len(pcm), case["chunk"]):
        b.write(pcm[None, a:a + case["chunk"]])
        acc.append(b.read_softbits()[0])
    want = np.concatenate(acc)
    b.close()
    group = 12 if "msk" in kind else 32
    assert len(want) - len(got) in range(0, group)                      # a last partial group stays pending, as in the reference
    assert np.array_equal(got, want[:len(got)])
    if not kind.startswith("burst"):
        assert len(got) == case["n_soft"]                                # what the verbatim reference emits on this excerpt

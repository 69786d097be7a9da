import numpy as np
"""CPU checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/jaero_b200.h declares; without a GPU the product fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, has_cuda


def _declared():
    src = open(os.path.join(ROOT, "include", "jaero_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jaero_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from jaero_b200 import build
    so = build.build()
    L = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "missing export " + n
    import jaero_b200
    assert set(jaero_b200.EXPORTS) <= set(names)


def test_python_structs_match_header_layout(tmp_path):
    """sizeof / offsetof as the C compiler sees include/jaero_b200.h == the ctypes mirrors in jaero_b200/__init__.py"""
    import subprocess
    import jaero_b200
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "jaero_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(jaero_settings), sizeof(jaero_status), offsetof(jaero_status, samples), offsetof(jaero_status, dcd),'
                   'offsetof(jaero_status, peak_volume), offsetof(jaero_status, scatter), sizeof(jaero_acars_record));return 0;}\n')
    exe = str(tmp_path / "layout")
    subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), str(src), "-o", exe], check=True)
    got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    S = jaero_b200.Status
    assert got == [ctypes.sizeof(jaero_b200.Settings), ctypes.sizeof(S), S.samples.offset, S.dcd.offset, S.peak_volume.offset, S.scatter.offset,
                   ctypes.sizeof(jaero_b200.AcarsRecord)]


@pytest.mark.skipif(has_cuda(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    import jaero_b200
    with pytest.raises(jaero_b200.JaeroError):
        jaero_b200.DemodBatch("oqpsk", 4, fb=10500)
    with pytest.raises(jaero_b200.JaeroError):
        jaero_b200.ViterbiBatch(4)
    with pytest.raises(jaero_b200.JaeroError):
        jaero_b200.PChannelBatch(4, 10500)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under jaero_b200/ may include, import or link it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "jaero_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r'#include\s*[<"][^>"]*oracle', txt), f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f
                assert "libjaero_oracle" not in txt and "libjaero_ref" not in txt, f


def test_ingest_router_files_messages_by_topic_prefix():
    """SURVEY 8(f)4 (host side): [topic][uint32 rate][int16 PCM] messages (zmq_audioreceiver.cpp:37-87) are filed under the
    channel whose 5-byte subscription prefix matches; a foreign rate or topic is rejected; no GPU involved."""
    import struct
    import jaero_b200
    r = jaero_b200.IngestRouter(["VFO01-long-name", "VFO02", "abc"], sample_rate=48000, capacity_samples=1000)
    rate = struct.pack("<I", 48000)
    pcm = np.arange(300, dtype=np.int16).tobytes()
    assert r.message(b"VFO02", rate, pcm) == 1
    assert r.message(b"VFO01 whatever follows", rate, pcm[:200]) == 0          # prefix of 5 bytes, as ZMQ_SUBSCRIBE with length 5
    assert r.available == 0                                                    # channel 2 has nothing yet
    assert r.message(b"abc", rate, pcm[:101]) == 2                             # odd byte count: 50 samples
    assert r.available == 50
    with pytest.raises(jaero_b200.JaeroError, match="no channel"):
        r.message(b"other", rate, pcm)
    with pytest.raises(jaero_b200.JaeroError, match="sample rate"):
        r.message(b"abc", struct.pack("<I", 44100), pcm)
    with pytest.raises(jaero_b200.JaeroError, match="4 bytes"):
        r.message(b"abc", b"\x00\x00", pcm)
    r.close()

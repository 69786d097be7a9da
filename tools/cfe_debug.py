"""Compare the coarse estimates of every epoch, GPU vs oracle, on the three test variants of an excerpt."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jaero_b200  # noqa: E402
from oracle import restated  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "oqpsk_8400"
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "expected_outputs.json")))
case = golden[name]
pcm = np.load(os.path.join(ROOT, "tests", "golden", case["excerpt"] + "_excerpt.npz"))["pcm"]
pcm2 = np.stack([pcm, (pcm.astype(np.int32) * 2 // 3).astype(np.int16), pcm[::-1].copy()])
kw = dict(case["kw"])
b = jaero_b200.DemodBatch(case["kind"], 3, **kw)
g = [[] for _ in range(3)]
nsoft = [0, 0, 0]
for a in range(0, pcm2.shape[1], 4096):
    b.write(pcm2[:, a:a + 4096])
    st = b.status()
    for c in range(3):
        g[c].append(st[c]["cfe_est"])
    for c, s in enumerate(b.read_softbits()):
        nsoft[c] += len(s)
for c in range(3):
    o = restated.OracleDemod(case["kind"], **kw)
    for a in range(0, pcm2.shape[1], 4096):
        o.write(pcm2[c, a:a + 4096])
    log = np.asarray(o.take_cfe_log())
    gg = np.asarray(g[c])[:len(log)]
    bad = np.nonzero(gg != log[:len(gg)])[0]
    print(f"ch{c}: soft gpu {nsoft[c]} oracle {len(o.take_soft())}; epochs {len(log)} / gpu {len(g[c])}; mismatching epochs {bad[:10]} "
          f"gpu {gg[bad[:5]]} oracle {log[bad[:5]]}")

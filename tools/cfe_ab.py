"""A/B the two coarse-estimator implementations (cluster-resident vs four-pass) on noise: per-epoch estimates must agree."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jaero_b200  # noqa: E402

C = 512
rng = np.random.default_rng(3)
pcm = (rng.standard_normal((C, 4096 * 8)) * 2500).astype(np.int16)
pcm[C // 2:] = (pcm[C // 2:].astype(np.int32) * 2 // 3).astype(np.int16)
for fb in (10500.0, 8400.0):
    out = {}
    for mode in ("1", "0"):
        os.environ["JAERO_CFE_CLUSTER"] = mode
        b = jaero_b200.DemodBatch("oqpsk", C, fb=fb, freq_center=8000.0, lockingbw=10500.0, afc=True, fft_power=14)
        ests = []
        for a in range(0, pcm.shape[1], 4096):
            b.write(pcm[:, a:a + 4096])
            ests.append([s["cfe_est"] for s in b.status()])
            b.read_softbits()
        out[mode] = np.asarray(ests)
        b.close()
    d = out["1"] != out["0"]
    print(f"fb={fb}: {d.sum()} of {d.size} estimates differ; epochs with differences: {np.nonzero(d.any(axis=1))[0]}")
    if d.any():
        e, c = np.argwhere(d)[0]
        print("  first:", e, c, out["1"][e, c], out["0"][e, c])

"""Time the coarse-estimator kernels (K2) and the segment kernel alone: N channels of noise, a few estimator epochs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import jaero_b200  # noqa: E402

if __name__ == "__main__":
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    fb = float(sys.argv[2]) if len(sys.argv) > 2 else 10500.0
    kind = sys.argv[3] if len(sys.argv) > 3 else "oqpsk"
    rng = np.random.default_rng(1)
    n = 4096 * 5
    pcm = (rng.standard_normal((64, n)) * 3000).astype(np.int16)
    pcm = np.ascontiguousarray(np.tile(pcm, (C // 64 + 1, 1))[:C])
    b = jaero_b200.DemodBatch(kind, C, fb=fb, freq_center=8000.0 if kind == "oqpsk" else 2000.0, lockingbw=10500.0 if kind == "oqpsk" else 1800.0,
                              afc=True, fft_power=14 if kind == "oqpsk" else 13)
    b.write(pcm)
    b.read_softbits()
    b.set_profiling(True); b.get_profile()
    b.write(pcm)
    b.sync()
    pr = b.get_profile()
    print("channels %d: cfe %.3f ms/run (%d runs), segment %.3f ms/launch (%d launches, %.2f us/sample-step)" % (
        C, pr["cfe_ms"] / max(pr["cfe_runs"], 1), pr["cfe_runs"], pr["segment_ms"] / max(pr["segment_launches"], 1), pr["segment_launches"],
        1e3 * pr["segment_ms"] / n))

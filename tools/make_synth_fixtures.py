"""Synthetic PCM fixtures (committed): signals the reference has no recording for.

  tests/golden/msk_1200_excerpt.npz   BASELINE cfg 2 signal model: 10 s of continuous 1200 bps MSK P-channel frames
                                      (jaero_b200.synth.msk_pchannel_pcm), carrier 2037 Hz, AWGN at Eb/N0 = 8 dB, RMS 0.2 FS

The generator is deterministic (numpy PCG64), but the array is committed so that the golden digests in
tests/golden/expected_outputs.json (made from it by tools/make_golden_outputs.py with the VERBATIM reference build)
do not depend on the FFT / libm of the machine that runs the tests."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jaero_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    pcm = synth.msk_pchannel_pcm(10, fc=2037.0, seed=0x4A41, ebn0_db=8.0, fb=1200.0, phase=1.1, delay=23)
    assert len(pcm) == 480000
    np.savez_compressed(os.path.join(GOLD, "msk_1200_excerpt.npz"), pcm=pcm)
    mpath = os.path.join(GOLD, "pcm_manifest.json")
    manifest = json.load(open(mpath))
    manifest["msk_1200"] = {"source": "synthetic: synth.msk_pchannel_pcm(10, fc=2037.0, seed=0x4A41, ebn0_db=8.0, fb=1200.0, phase=1.1, delay=23)",
                            "native_rate": 48000, "samples": int(len(pcm)), "sha256": hashlib.sha256(pcm.tobytes()).hexdigest(),
                            "excerpt": [0, int(len(pcm))], "excerpt_sha256": hashlib.sha256(pcm.tobytes()).hexdigest()}
    with open(mpath, "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print("msk_1200", len(pcm), manifest["msk_1200"]["sha256"][:16])


if __name__ == "__main__":
    main()

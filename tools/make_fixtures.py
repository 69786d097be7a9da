"""Build the PCM fixtures from the reference's recordings (/root/reference/samples).

Run in the build container only (the GPU box has no /root/reference). Writes
  tests/golden/pcm_full/<name>.npy   full-length int16 @48 kHz (git-ignored, travels with gpurun)
  tests/golden/<name>_excerpt.npz    short committed excerpt (int16) used by the default tests
  tests/golden/pcm_manifest.json     lengths + SHA-256 of every array (committed)
Pinned recipe: Vorbis/MP3 44.1 kHz -> scipy.signal.resample_poly(x,160,147)
-> round(x*32767) clipped to int16; Opus/WAV are already 48 kHz.
"""
import hashlib
import json
import os
import sys

import numpy as np
from scipy.io import wavfile
from scipy.signal import resample_poly

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from audio_decode import decode_mono  # noqa: E402

REF = "/root/reference/samples"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
FULL = os.path.join(GOLD, "pcm_full")

# name -> (file, native rate, excerpt [start_s, dur_s])
ITEMS = {
    "oqpsk_10500": ("10.5k_sample.ogg", 44100, (0.0, 12.0)),
    "msk_600": ("600bps_sample.ogg", 48000, (0.0, 20.0)),
    "oqpsk_8400": ("8400bps_ambe_sample.ogg", 48000, (0.0, 8.0)),
    "burst_oqpsk_10500": ("10.5k_burst_sample.mp3", 44100, (0.0, 8.0)),
    "burst_msk_1200_a": ("1200bps_burst_sample1.wav", 48000, None),
    "burst_msk_1200_b": ("1200bps_burst_sample2.wav", 48000, None),
}


def to_i16(x):
    return np.clip(np.round(x.astype(np.float64) * 32767.0), -32768, 32767).astype(np.int16)


def main():
    os.makedirs(FULL, exist_ok=True)
    manifest = {}
    for name, (fn, rate, exc) in ITEMS.items():
        path = os.path.join(REF, fn)
        if fn.endswith(".wav"):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r, x = wavfile.read(path)
            assert r == 48000 and x.dtype == np.int16
            pcm = x if x.ndim == 1 else x[:, 0]
        else:
            f, _ = decode_mono(path)
            if rate == 44100:
                f = resample_poly(f.astype(np.float64), 160, 147)
            pcm = to_i16(f)
        pcm = np.ascontiguousarray(pcm)
        np.save(os.path.join(FULL, name + ".npy"), pcm)
        entry = {"source": fn, "native_rate": rate, "samples": int(len(pcm)),
                 "sha256": hashlib.sha256(pcm.tobytes()).hexdigest()}
        if exc is None:
            ex = pcm
            entry["excerpt"] = [0, int(len(pcm))]
        else:
            a = int(exc[0] * 48000); b = a + int(exc[1] * 48000)
            ex = pcm[a:b]
            entry["excerpt"] = [a, b]
        entry["excerpt_sha256"] = hashlib.sha256(np.ascontiguousarray(ex).tobytes()).hexdigest()
        np.savez_compressed(os.path.join(GOLD, name + "_excerpt.npz"), pcm=ex)
        manifest[name] = entry
        print(name, entry["samples"], entry["sha256"][:16])
    with open(os.path.join(GOLD, "pcm_manifest.json"), "w") as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

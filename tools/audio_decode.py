"""Decode the reference's compressed recordings (Ogg/Vorbis, Ogg/Opus, MP3) to float32 mono.

Fixture tooling only (runs in the build container, where /root/reference exists).
There is no ffmpeg binary / libvorbis / torchcodec in the image, but the FFmpeg
shared libraries bundled with opencv-python-headless can be driven through ctypes
(probe P2). Struct offsets are for the bundled FFmpeg 8 build
(libavformat 62 / libavcodec 62 / libavutil 60) on x86-64.
"""
import ctypes
import glob
import os

import numpy as np


def _libs():
    import cv2  # noqa: F401  (only to locate site-packages)
    base = os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs")
    def load(stem):
        path = sorted(glob.glob(os.path.join(base, stem + "-*.so*")))[0]
        return ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    avutil = load("libavutil")
    try:
        load("libswresample")
    except Exception:
        pass
    avcodec = load("libavcodec")
    avformat = load("libavformat")
    return avutil, avcodec, avformat


def decode_mono(path):
    """-> (float32 samples, sample_rate)"""
    avutil, avcodec, avformat = _libs()
    vp = ctypes.c_void_p
    avformat.avformat_open_input.argtypes = [ctypes.POINTER(vp), ctypes.c_char_p, vp, vp]
    avformat.avformat_find_stream_info.argtypes = [vp, vp]
    avformat.av_read_frame.argtypes = [vp, vp]
    avcodec.avcodec_find_decoder.restype = vp
    avcodec.avcodec_find_decoder.argtypes = [ctypes.c_int]
    avcodec.avcodec_alloc_context3.restype = vp
    avcodec.avcodec_alloc_context3.argtypes = [vp]
    avcodec.avcodec_parameters_to_context.argtypes = [vp, vp]
    avcodec.avcodec_open2.argtypes = [vp, vp, vp]
    avcodec.av_packet_alloc.restype = vp
    avcodec.av_packet_unref.argtypes = [vp]
    avcodec.avcodec_send_packet.argtypes = [vp, vp]
    avcodec.avcodec_receive_frame.argtypes = [vp, vp]
    avutil.av_frame_alloc.restype = vp
    avutil.av_frame_unref.argtypes = [vp]

    fmt = vp(None)
    if avformat.avformat_open_input(ctypes.byref(fmt), path.encode(), None, None) < 0:
        raise RuntimeError("cannot open " + path)
    avformat.avformat_find_stream_info(fmt, None)
    nb_streams = ctypes.c_uint.from_address(fmt.value + 44).value
    streams = vp.from_address(fmt.value + 48).value
    assert nb_streams >= 1
    st0 = vp.from_address(streams).value
    codecpar = vp.from_address(st0 + 16).value
    codec_id = ctypes.c_int.from_address(codecpar + 4).value
    dec = avcodec.avcodec_find_decoder(codec_id)
    ctx = avcodec.avcodec_alloc_context3(dec)
    avcodec.avcodec_parameters_to_context(ctx, codecpar)
    if avcodec.avcodec_open2(ctx, dec, None) < 0:
        raise RuntimeError("cannot open decoder")
    pkt = avcodec.av_packet_alloc()
    frame = avutil.av_frame_alloc()
    out = []
    rate = None

    def drain():
        nonlocal rate
        while avcodec.avcodec_receive_frame(ctx, frame) == 0:
            ext = vp.from_address(frame + 96).value
            n = ctypes.c_int.from_address(frame + 112).value
            f = ctypes.c_int.from_address(frame + 116).value
            if rate is None:
                rate = ctypes.c_int.from_address(frame + 120).value  # sample_rate follows format
            p0 = vp.from_address(ext).value
            if f in (3, 8):      # AV_SAMPLE_FMT_FLT / FLTP (mono: identical layout)
                a = np.frombuffer((ctypes.c_float * n).from_address(p0), dtype=np.float32).copy()
            elif f in (1, 6):    # S16 / S16P
                a = np.frombuffer((ctypes.c_int16 * n).from_address(p0), dtype=np.int16).astype(np.float32) / 32768.0
            else:
                raise RuntimeError("unsupported sample format %d" % f)
            out.append(a)
            avutil.av_frame_unref(frame)

    while avformat.av_read_frame(fmt, pkt) >= 0:
        avcodec.avcodec_send_packet(ctx, pkt)
        avcodec.av_packet_unref(pkt)
        drain()
    avcodec.avcodec_send_packet(ctx, None)
    drain()
    return np.concatenate(out), rate


if __name__ == "__main__":
    import sys
    x, r = decode_mono(sys.argv[1])
    print(len(x), r, float(np.abs(x).max()))

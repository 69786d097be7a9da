#!/usr/bin/env python
"""bench.py — headline benchmark of the jaero_b200 hot path.

Workload (BASELINE.json configs[2], the one the metric is quoted on): 4096 concurrent continuous 10.5 kbps
OQPSK P-channels per GPU, synthetic real-passband int16 @48 kHz (Eb/N0 = 10 dB), each step = 1 s of signal per
channel through  demodulator (K1a) + coarse frequency estimator (K2) + P-channel framing + fused
de-interleave/Viterbi (K5) + descramble + CRC  with DCD fed back. Metric: Msamples/s (one sample = one int16 input
sample of one channel); channels@RT = samples/s / 48000.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one process per GPU under torchrun)
  python bench.py --impl reference ...                      the reference's own CPU path on this box's host cores

Prints ONE JSON line (rank 0). `value` = device-timed, input resident in HBM; `e2e` = through the C ABI with HOST
buffers (H2D of the PCM and D2H of the decoded signal units inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48000
FB = 10500
ALG_BYTES_PER_SAMPLE = 194.2          # faithful fp64 state incl. the EbNo observable
STEP_SAMPLES = 48000                  # 1 s = 2 P-channel frames per channel per step



def _ncu_traffic_bytes():
    """dram__bytes_read.sum + dram__bytes_write.sum of one K1a launch, from the committed ncu --set full capture (None if absent)."""
    import csv
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_oqpsk_pipe_v8_full_raw.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr, units, row = rows[0], rows[1], rows[2]
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(k)
            tot += float(row[i].replace(",", "")) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return tot
    except Exception:
        return None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=4096, help="channels per GPU (weak scaling)")
    ap.add_argument("--ebn0", type=float, default=10.0)
    ap.add_argument("--cpu-seconds", type=float, default=2.0, help="seconds of signal per channel for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------- synthetic input
def base_envelopes(n_base, seed):
    """n_base distinct, seamlessly loopable 1 s (2-frame) complex envelopes (numpy, rank 0)."""
    from jaero_b200 import synth
    envs = []
    for k in range(n_base):
        bits = synth.pchannel_bits(FB, 2, seed=seed + k)
        envs.append(synth.oqpsk_envelope(bits, FB, FS))
    return np.stack(envs).astype(np.complex64)


def make_pcm_gpu(envs_t, ch0, n_ch, ebn0_db, device):
    """Per-channel real passband int16 on the GPU: base envelope (c % B) with a circular delay, integer-Hz carrier
    8000 + U(-500,500), random phase, AWGN at Eb/N0, RMS 0.2 FS. Seeds depend on the GLOBAL channel index."""
    import torch
    B, L = envs_t.shape
    out = torch.empty((n_ch, L), dtype=torch.int16, device=device)
    n = torch.arange(L, device=device, dtype=torch.float64)
    fcs = np.zeros(n_ch)
    for a in range(0, n_ch, 256):
        m = min(256, n_ch - a)
        g = torch.Generator(device="cpu"); g.manual_seed(0x4A4145524F + ch0 + a)
        delay = torch.randint(0, L, (m,), generator=g)
        fc = 8000 + torch.randint(-500, 501, (m,), generator=g).to(torch.float64)
        ph = torch.rand((m,), generator=g, dtype=torch.float64) * 2 * np.pi
        fcs[a:a + m] = fc.numpy()
        idx = (torch.arange(L).unsqueeze(0) - delay.unsqueeze(1)) % L
        k = (torch.arange(ch0 + a, ch0 + a + m) % B)
        env = envs_t[k.to(device).unsqueeze(1), idx.to(device)]
        arg = (2 * np.pi * fc.to(device).unsqueeze(1) * n.unsqueeze(0) / FS + ph.to(device).unsqueeze(1))
        x = (env.real.to(torch.float64) * torch.cos(arg) - env.imag.to(torch.float64) * torch.sin(arg)).to(torch.float32)
        ps = (x * x).mean(dim=1, keepdim=True)
        if ebn0_db is not None:
            n0 = ps * (FS / FB) / (10 ** (ebn0_db / 10.0))
            gg = torch.Generator(device=device); gg.manual_seed(12345 + ch0 + a)
            x = x + torch.randn(x.shape, generator=gg, device=device) * torch.sqrt(n0 / 2.0)
        x = x * (0.2 / torch.sqrt((x * x).mean(dim=1, keepdim=True)))
        out[a:a + m] = torch.clamp(torch.round(x * 32767.0), -32768, 32767).to(torch.int16)
    return out, fcs


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi polled every 100 ms from BEFORE the warm-up (its start-up takes longer than a short timed region);
    only the rows received between mark_begin() and stop() - i.e. during the timed region - are reported."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        self.t_begin = None

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t_end = time.perf_counter()
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        for t, r in self.rows:
            if t < t0 or t > t_end + 0.05:
                continue
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference arm
def usable_cores():
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:                                                     # cgroup v2 / v1 CPU quota
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(float(q) / float(per))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read()); per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // per))
        except Exception:
            pass
    return n


def _cpu_worker(args):
    """One host core: the reference's demodulator (oracle/_ref, verbatim build) or the restated port, followed by the
    restated AeroL P-channel layer (frame sync, de-interleave, Viterbi, CRC) with DCD fed back — the same work per
    sample as the GPU pipeline. Returns (samples, seconds, su_total, su_ok)."""
    kind, pcm, n_steps, fc = args
    from oracle import ref, restated
    kw = dict(fb=FB, freq_center=fc, lockingbw=10500, fft_power=14, signalthreshold=0.65, afc=False)
    d = ref.RefDemod("oqpsk", **kw) if kind == "reference" else restated.OracleDemod("oqpsk", **kw)
    p = restated.OraclePChannel(FB)
    t0 = time.perf_counter()
    tot = ok = 0
    for s in range(n_steps):
        for a in range(0, len(pcm), 4800):                   # 100 ms writeData calls
            d.write(pcm[a:a + 4800])
            p.process(d.take_soft())
            d.set_dcd(p.dcd)
        p.update_dcd()
        _, o, _ = p.take_sus()
        tot += len(o); ok += int(o.sum())
    return len(pcm) * n_steps, time.perf_counter() - t0, tot, ok


def cpu_reference_run(pcm_rows, fcs, n_steps, cores):
    import multiprocessing as mp
    from oracle import ref, restated
    kind = "reference" if ref.available() else "port"
    if kind == "port" and not restated.available():
        raise RuntimeError("neither oracle/_ref nor oracle/_build is built")
    jobs = [(kind, pcm_rows[i], n_steps, float(fcs[i])) for i in range(len(pcm_rows))]
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(cores) as pool:        # fresh processes: the reference keeps function-local statics
        res = pool.map(_cpu_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    samples = sum(r[0] for r in res)
    busy = max(r[1] for r in res)                            # slowest worker's writeData time (start-up excluded)
    return kind, samples, busy, wall, sum(r[2] for r in res), sum(r[3] for r in res)


# ----------------------------------------------------------------------------- main
def main():
    a = parse()
    import torch
    from jaero_b200 import shard
    cores = usable_cores()
    if a.impl == "reference":
        # CPU arm: rank 0 alone measures and prints; the other ranks exit without joining any process group
        rank = int(os.environ.get("RANK", "0"))
        n_gpus = max(a.gpus, int(os.environ.get("WORLD_SIZE", "1")))
        if rank != 0:
            return 0
        envs = base_envelopes(16, 7)
        envs_t = torch.from_numpy(envs)
        n_cpu = cores
        pcm, fcs = make_pcm_gpu(envs_t, 0, n_cpu, a.ebn0, "cpu")
        rows = [pcm[i].numpy().copy() for i in range(n_cpu)]
        cpu_reference_run(rows[:min(2, n_cpu)], fcs, 1, min(2, n_cpu))           # warm the page cache / libm
        kind, samples, busy, wall, tot, ok = cpu_reference_run(rows, fcs, max(1, a.steps), cores)
        val = samples / busy / 1e6
        line = {"impl": "reference", "metric": "IQ Msamples/s (10.5k OQPSK demod + Viterbi)", "value": val, "unit": "Msamples/s",
                "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": busy * 1e3 / max(1, a.steps),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "channels_rt": val * 1e6 / FS,
                "config": {"workload": "4096-channel 10.5 kbps continuous OQPSK + Viterbi (BASELINE configs[2]); reference arm = bounded sample",
                           "channels": n_cpu, "seconds_per_step": 1.0, "ebn0_db": a.ebn0},
                "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": cores, "kind": kind,
                                 "sample": "%d channels x %d s, one process per core, 100 ms writeData calls, demod + P-channel decode" % (n_cpu, max(1, a.steps))},
                "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "decode": {"su_total": tot, "su_crc_ok": ok}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    rank, local, world = shard.init_from_env()
    n_gpus = max(a.gpus, world)
    import jaero_b200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    C = a.channels
    ch0 = rank * C

    # shared source: base envelopes built on rank 0 and broadcast over NCCL/NVLink (the only collective besides reporting)
    envs_t = torch.zeros((16, STEP_SAMPLES), dtype=torch.complex64, device=dev)
    if rank == 0:
        envs_t.copy_(torch.from_numpy(base_envelopes(16, 7)))
    ev = torch.view_as_real(envs_t)
    shard.broadcast_(ev, 0)
    pcm, fcs = make_pcm_gpu(envs_t, ch0, C, a.ebn0, dev)
    torch.cuda.synchronize()

    batch = jaero_b200.DemodBatch("oqpsk", C, fb=FB, freq_center=fcs, lockingbw=10500, afc=False, report_ebno=True, device=local)
    pch = jaero_b200.PChannelBatch(C, FB, device=local)
    stream = torch.cuda.Stream(device=dev)          # a real (non-default) stream: handle 0 would mean "library's own stream"
    assert stream.cuda_stream != 0
    batch.set_stream(stream.cuda_stream)            # demod segments, estimator, frame layer, Viterbi all launch here
    stride = pcm.stride(0)

    def step_device():
        batch.write_device(pcm.data_ptr(), STEP_SAMPLES, stride)
        pch.process_batch(batch)
        pch.tick(batch)
        pch.discard_sus()          # results stay on the device for the HBM-resident measurement

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()             # nvidia-smi needs longer to start than a short timed region lasts: start it before the warm-up
    for _ in range(max(3, a.warmup)):
        step_device()
    torch.cuda.synchronize()
    shard.barrier()

    # ---- timed region (device, CUDA events on the launching stream)
    l0 = batch.launches + pch.launches
    batch.set_profiling(True); batch.get_profile()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); shard.barrier()
    clocks.mark_begin()
    e0.record(stream)
    for _ in range(a.steps):
        step_device()
    e1.record(stream)
    torch.cuda.synchronize(); shard.barrier()
    ms = shard.reduce_max(e0.elapsed_time(e1), device=dev)
    launches = batch.launches + pch.launches - l0
    prof = batch.get_profile(); batch.set_profiling(False)
    clk = clocks.stop() if rank == 0 else None
    dcd, su_tot, su_ok = pch.stats()
    total_samples = float(C) * STEP_SAMPLES * a.steps * world
    value = total_samples / (ms * 1e-3) / 1e6

    # ---- end-to-end through the C ABI with HOST buffers
    e2e = None
    if not a.no_e2e:
        host = torch.empty((C, STEP_SAMPLES), dtype=torch.int16).pin_memory()
        host.copy_(pcm.cpu())
        hnp = host.numpy()
        su_buf = np.empty((C, pch.su_cap, 16), dtype=np.uint8); su_cnt = np.zeros(C, dtype=np.int32)
        def step_e2e():
            batch.write(hnp)                       # H2D of the step's PCM (pinned) inside the call
            pch.process_batch(batch)
            pch.tick(batch)
            return pch.read_sus_raw(su_buf, su_cnt)   # D2H of the decoded signal units + CRC flags (bulk records, no per-channel Python objects)
        step_e2e()
        torch.cuda.synchronize(); shard.barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            sus = step_e2e()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        ems = shard.reduce_max((t1 - t0) * 1e3, device=dev)
        d2h = C * pch.su_cap * 16 + C * 120
        e2e = {"value": total_samples / (ems * 1e-3) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(C * STEP_SAMPLES * 2),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": ems / a.steps}

    tot = shard.reduce_sum([float(su_tot.sum()), float(su_ok.sum()), float(dcd.sum())], device=dev)

    # ---- roofline of the dominant kernel (oqpsk_pipe_kernel), measured live with events around every launch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    seg_s = prof["segment_ms"] * 1e-3
    achieved = (ALG_BYTES_PER_SAMPLE * prof["samples"] * C) / seg_s / 1e9 if seg_s > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "oqpsk_pipe_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": _ncu_traffic_bytes(), "traffic_source": "profiles/r01_oqpsk_pipe_v8_full_raw.csv (ncu --set full, one launch = 4096 samples x 4096 channels)",
                "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                "alg_bytes_per_sample": ALG_BYTES_PER_SAMPLE,
                "avg_launch_ms": prof["segment_ms"] / max(1, prof["segment_launches"]), "launches": prof["segment_launches"],
                "share_of_step": prof["segment_ms"] / (e0.elapsed_time(e1)), "cfe_share_of_step": prof["cfe_ms"] / (e0.elapsed_time(e1)),
                "note": "per-channel fp64 feedback loop: bound by dependent-issue latency, not HBM (see DESIGN.md section 5)"}

    cpu_base = None
    if rank == 0 and not a.no_cpu_baseline:
        n_cpu = cores
        secs = max(1, int(round(a.cpu_seconds)))
        rows = [pcm[i].cpu().numpy().copy() for i in range(n_cpu)]
        _, s1, b1, _, _, _ = cpu_reference_run(rows[:1], fcs[:1], 1, 1)          # one core alone (also warms caches)
        kind, samples, busy, wall, ctot, cok = cpu_reference_run(rows, fcs[:n_cpu], secs, cores)
        cpu_base = {"value": samples / busy / 1e6, "unit": "Msamples/s", "cores": cores, "kind": kind,
                    "single_core_value": s1 / b1 / 1e6,
                    "sample": "%d channels x %d s of the same synthetic workload, one process per core, demod + P-channel decode" % (n_cpu, secs),
                    "su_total": ctot, "su_crc_ok": cok}

    if rank == 0:
        line = {"metric": "IQ Msamples/s (10.5k OQPSK demod + Viterbi)", "value": value, "unit": "Msamples/s", "n_gpus": world,
                "steps": a.steps, "warmup": max(3, a.warmup), "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "channels_rt": value * 1e6 / FS,
                "config": {"workload": "4096-channel 10.5 kbps continuous OQPSK + Viterbi, synthetic real-passband int16 @48 kHz (BASELINE configs[2])",
                           "channels_per_gpu": C, "seconds_per_step": 1.0, "ebn0_db": a.ebn0, "parallelism": "channels sharded x%d, no data-path collective" % world,
                           "l2": "inputs (%.0f MB int16 + %.1f GB of ring state per step) exceed the 126 MB L2" % (C * STEP_SAMPLES * 2 / 1e6, C * 3.4e-3)},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu_base,
                "decode": {"su_total": tot[0], "su_crc_ok": tot[1], "channels_with_dcd": tot[2]}}
        print(json.dumps(line))
    batch.close(); pch.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

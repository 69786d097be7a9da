"""Multi-process host logic on CPU (gloo, world_size 2): channel partition, broadcast of the shared source
waveform, counter reduction, max-over-ranks timing. No data-path collective exists."""
import os
import socket

import torch
import torch.multiprocessing as mp

from jaero_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, _, w = shard.init_from_env(backend="gloo")
    lo, hi = shard.channel_range(4097, w, r)
    wave = torch.arange(1000, dtype=torch.float32) if r == 0 else torch.zeros(1000)
    shard.broadcast_(wave, 0)
    tmax = shard.reduce_max(10.0 + r)
    tot = shard.reduce_sum([hi - lo, 1.0])
    rows = shard.gather_rows([r, lo, hi])
    shard.barrier()
    q.put((r, lo, hi, float(wave.sum()), tmax, tot, rows.tolist()))
    torch.distributed.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    (r0, lo0, hi0, s0, t0, tot0, rows0), (r1, lo1, hi1, s1, t1, tot1, rows1) = res
    assert (lo0, hi1) == (0, 4097) and hi0 == lo1              # contiguous, disjoint, complete
    assert s0 == s1 == float(sum(range(1000)))                 # broadcast delivered the rank-0 waveform
    assert t0 == t1 == 11.0                                    # max over ranks
    assert tot0 == tot1 == [4097.0, 2.0]
    assert rows0 == rows1 == [[0, lo0, hi0], [1, lo1, hi1]]


def test_partitions():
    for n in (1, 7, 4096, 16384):
        for w in (1, 2, 4, 8):
            rs = [shard.channel_range(n, w, r) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n and all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
    costs = [1.0] * 12288 + [1.6] * 4096                      # 10.5k OQPSK + 8400 C-channel mix (cfg 5)
    rs = shard.weighted_ranges(costs, 8)
    loads = [sum(costs[a:b]) for a, b in rs]
    assert rs[0][0] == 0 and rs[-1][1] == len(costs) and max(loads) / min(loads) < 1.05

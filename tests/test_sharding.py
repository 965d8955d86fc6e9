"""N>1 path on CPU: two gloo ranks each demodulate their shard of the channels (with the host
warp emulator standing in for the GPU), rank 0 merges; the union must equal the one-rank result
(no collective on the data path, only plumbing)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.pyoracle import Oracle, link_decode
from tests import emul
from tests.parity_util import make_channels
from vdlm2dec_b200 import shard

NCH, NS = 4, 700_000


def _demod_shard(chs, specs, iq):
    out = []
    for c in chs:
        o = Oracle("port", chn=c, Fo=specs[c].Fo).feed(iq[c])
        b, _, _, _ = emul.demod(o.dumps, 2688, chn=c)
        out.append(b)
    return np.concatenate(out) if out else np.zeros(0, dtype=emul.BLOCK_DT)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    specs, iq = make_channels(NCH, NS, seed=8)
    mine = shard.shard_channels(NCH, 1, rank, world)
    blocks = _demod_shard(mine, specs, iq)
    t = shard.max_over_ranks(1.0 + rank)
    merged = shard.gather_blocks(blocks)
    frames = shard.gather_frames(link_decode("port", blocks)[0])     # row f1: every rank runs the block pipeline on its own blocks
    if rank == 0:
        q.put((t, merged.tobytes(), len(merged), frames.tobytes(), len(frames)))
    dist.barrier()
    dist.destroy_process_group()


def test_partition_is_disjoint_and_complete():
    for world in (1, 2, 4, 8):
        for total, cps in ((1024, 1), (128, 8), (10, 1)):
            got = sorted(c for r in range(world) for c in shard.shard_channels(total, cps, r, world))
            assert got == list(range(total * cps))
            assert all(s % world == r for r in range(world) for s in shard.shard_streams(total, r, world))


def test_two_rank_gloo_union_equals_single_rank():
    emul.build()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    t, raw, n, fraw, nf = q.get(timeout=300)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert t == 2.0  # max over ranks
    specs, iq = make_channels(NCH, NS, seed=8)
    single = shard.merge_blocks([_demod_shard(list(range(NCH)), specs, iq)])
    assert n == len(single) > 0 and raw == single.tobytes()
    fsingle = shard.merge_frames([link_decode("port", single)[0]])
    assert nf == len(fsingle) > 0 and fraw == fsingle.tobytes()

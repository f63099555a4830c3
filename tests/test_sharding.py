"""CPU: the N>1 path (rank slicing + the single all_gather) with gloo, world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from generativeimage2text_b200.sharding import shard_range, gather_captions


def test_shard_range_matches_reference_formula():
    # reference inference.py:165-169: ceil(N/W) per rank, last rank short
    for n, w in [(10, 3), (8192, 8), (5, 8), (1, 1), (64, 2)]:
        per = -(-n // w)
        got = [shard_range(n, r, w) for r in range(w)]
        rows = [i for s, e in got for i in range(s, e)]
        assert rows == list(range(n))
        for r, (s, e) in enumerate(got):
            assert s == min(per * r, max(s, 0)) or s == e
            assert e - s <= per


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rows, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    s, e = shard_range(n_rows, rank, world)
    toks = torch.arange(s, e)[:, None].repeat(1, 5) * 10 + torch.arange(5)[None]
    lps = -torch.arange(s, e).float()
    all_t, all_l = gather_captions(toks, lps, n_rows)
    ok = torch.equal(all_t, torch.arange(n_rows)[:, None] * 10 + torch.arange(5)[None]) and \
        torch.equal(all_l, -torch.arange(n_rows).float())
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_gather_world2_gloo():
    world, n_rows = 2, 7      # uneven: shards of 4 and 3
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n_rows, ret), nprocs=world, join=True)
        assert dict(ret) == {0: True, 1: True}

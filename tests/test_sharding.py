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


# ---- merged prediction TSVs of the sharded TSV path (generativeimage2text_b200.inference.write_rows_sharded) ------------
def _rows_of(n_rows, rank, world):
    s, e = shard_range(n_rows, rank, world)
    return [('key%03d' % i, '[{"caption":"c %d"}]' % i) for i in range(s, e)]


def _tsv_worker(rank, world, port, n_rows, out_tsv, ret):
    from generativeimage2text_b200 import inference as inf
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ret[rank] = inf.write_rows_sharded(iter(_rows_of(n_rows, rank, world)), out_tsv, rank, world)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_tsv_rows_one_gather_gloo(tmp_path):
    """world_size 2 with a process group: ONE gather to rank 0, which writes the merged TSV in row order."""
    from generativeimage2text_b200 import tsv_io
    world, n_rows = 2, 9
    out_tsv = str(tmp_path / 'pred.tsv')
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_tsv_worker, args=(world, _free_port(), n_rows, out_tsv, ret), nprocs=world, join=True)
        assert dict(ret) == {0: 5, 1: 4}
    t = tsv_io.TSVFile(out_tsv)
    assert [r for r in t] == [list(r) for r in _rows_of(n_rows, 0, 1)]
    assert not os.path.exists(out_tsv + '.0.2.tsv')      # no part files on this path


def test_sharded_tsv_rows_part_files_without_process_group(tmp_path):
    """Ranks launched without torch.distributed (mpirun, like the reference): part files + rank-0 concat
    (reference inference.py:159-162, 213-225); byte-identical to the single-process file."""
    import threading
    from generativeimage2text_b200 import inference as inf
    world, n_rows = 3, 10
    out_tsv = str(tmp_path / 'pred.tsv')
    single = str(tmp_path / 'single.tsv')
    assert inf.write_rows_sharded(iter(_rows_of(n_rows, 0, 1)), single, 0, 1) == n_rows
    counts = {}

    def run(rank):
        counts[rank] = inf.write_rows_sharded(iter(_rows_of(n_rows, rank, world)), out_tsv, rank, world, poll_s=0.02)
    th0 = threading.Thread(target=run, args=(0,))
    th0.start()                      # rank 0 finishes its own part first, then waits for the others
    for r in (2, 1):
        run(r)
    th0.join(timeout=30)
    assert not th0.is_alive() and counts == {0: 4, 1: 4, 2: 2}
    assert open(out_tsv, 'rb').read() == open(single, 'rb').read()
    base = os.path.splitext(out_tsv)[0]
    assert open(base + '.lineidx.8b', 'rb').read() == open(os.path.splitext(single)[0] + '.lineidx.8b', 'rb').read()

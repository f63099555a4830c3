"""Image-wise data parallelism of the TSV path: one process per GPU, the rank's contiguous slice of the rows
(identical to reference inference.py:152-169), and ONE collective at the end -- a single all_gather of the finished
token ids with the logprobs packed beside them -- replacing the reference's per-rank TSV part files + 5 s filesystem poll + byte concat
(reference inference.py:159-164, 213-225).  Images are independent, so there is no collective on the data path.
"""
import math
import os

import torch
import torch.distributed as dist


def get_mpi_rank():
    """reference common.py:106-110 semantics (torchrun RANK or OpenMPI env)."""
    return int(os.environ.get('RANK', os.environ.get('OMPI_COMM_WORLD_RANK', '0')))


def get_mpi_size():
    return int(os.environ.get('WORLD_SIZE', os.environ.get('OMPI_COMM_WORLD_SIZE', '1')))


def get_mpi_local_rank():
    return int(os.environ.get('LOCAL_RANK', os.environ.get('OMPI_COMM_WORLD_LOCAL_RANK', '0')))


def shard_range(num_rows, rank=None, world=None):
    """[start, end) of this rank: ceil(N/W) rows each, last shard short (reference inference.py:165-169)."""
    rank = get_mpi_rank() if rank is None else rank
    world = get_mpi_size() if world is None else world
    per = int(math.ceil(num_rows / world))
    start = per * rank
    end = min(start + per, num_rows)
    return start, max(start, end)


def gather_captions(tokens, logprobs, num_rows, pad_token=102, group=None):
    """all_gather the per-rank results; returns (tokens [num_rows, T], logprobs [num_rows]) in row order on
    every rank.  `tokens` is this rank's [n_local, T] int64 (n_local <= ceil(N/W)); short shards are padded
    for the collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return tokens[:num_rows], logprobs[:num_rows]
    world = dist.get_world_size(group)
    per = int(math.ceil(num_rows / world))
    T = tokens.shape[1]
    # ONE collective: int32 on the wire (vocabulary ids < 2^31), [per, T + 1] per rank -- the last column carries the fp32
    # logprob's bit pattern
    buf = torch.full((per, T + 1), pad_token, dtype=torch.int32, device=tokens.device)
    n = tokens.shape[0]
    buf[:n, :T] = tokens.to(torch.int32)
    lp = torch.zeros((per,), dtype=torch.float32, device=tokens.device)
    lp[:n] = logprobs.reshape(-1)[:n].float()
    buf[:, T] = lp.view(torch.int32)
    out = torch.empty((world * per, T + 1), dtype=torch.int32, device=tokens.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out[:num_rows, :T].long(), out[:num_rows, T].contiguous().view(torch.float32)

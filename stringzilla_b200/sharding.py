"""Multi-GPU partition of one cross product: every matrix cell is independent, so the batch is split
into contiguous blocks of QUERY ROWS balanced by total query length; every rank needs all candidates. One process per
GPU; the only collectives are one broadcast of the packed inputs and one gather of each rank's score slab.

The reference has no multi-GPU path at all (no NCCL/MPI symbol in the tree); this module sits above
the C ABI and drives one single-GPU scope per rank, which the reference's threading contract already allows
(types.cuh:433-444).
"""

from __future__ import annotations

from typing import Callable

import numpy as np


def partition_rows(query_lengths: np.ndarray, world_size: int) -> list[tuple[int, int]]:
    """Contiguous row blocks with near-equal sums of query length (the cost of a row block is
    `sum(len(q) in block) * sum(len(c))`). Returns `world_size` half-open (begin, end) ranges covering all rows."""
    lengths = np.asarray(query_lengths, dtype=np.int64)
    n = int(lengths.size)
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    # Weigh every row at least 1 so that empty strings still spread over ranks.
    cumulative = np.concatenate([[0], np.cumsum(np.maximum(lengths, 1))])
    total = int(cumulative[-1])
    bounds = [0]
    for rank in range(1, world_size):
        target = total * rank / world_size
        cut = int(np.searchsorted(cumulative, target, side="left"))
        bounds.append(min(max(cut, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def broadcast_tape(tape, src: int = 0, device=None, group=None):
    """Broadcasts a `Tape` (byte blob + offsets) from `src` with `torch.distributed`: two small size messages and
    two payload broadcasts. With the NCCL backend the payloads travel GPU to GPU over NVLink; `device` selects where
    the received blob lives. Returns (data_tensor, offsets_tensor)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    backend_device = device if device is not None else torch.device("cpu")
    header = torch.zeros(2, dtype=torch.int64, device=backend_device)
    if rank == src:
        header[0], header[1] = int(tape.data.size), int(tape.offsets.size)
    dist.broadcast(header, src=src, group=group)
    n_bytes, n_offsets = int(header[0]), int(header[1])
    if rank == src:
        data = torch.from_numpy(tape.data).to(backend_device)
        offsets = torch.from_numpy(tape.offsets.astype(np.int64)).to(backend_device)
    else:
        data = torch.empty(n_bytes, dtype=torch.uint8, device=backend_device)
        offsets = torch.empty(n_offsets, dtype=torch.int64, device=backend_device)
    dist.broadcast(data, src=src, group=group)
    dist.broadcast(offsets, src=src, group=group)
    return data, offsets


def broadcast_tapes(tapes, src: int = 0, device=None, group=None):
    """Broadcasts several `Tape`s from `src` as ONE payload (plus one small header): offsets first (8-byte aligned),
    then the byte blobs. Non-source ranks pass `None` entries. Returns a list of (data_tensor, offsets_numpy_uint64):
    the blobs stay on `device` (views into the received payload), the offsets come back to the host in one copy,
    because the C ABI reads them there."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    backend_device = device if device is not None else torch.device("cpu")
    count = len(tapes)
    header = torch.zeros(2 * count, dtype=torch.int64, device=backend_device)
    if rank == src:
        for i, tape in enumerate(tapes):
            header[2 * i], header[2 * i + 1] = int(tape.offsets.size), int(tape.data.size)
    dist.broadcast(header, src=src, group=group)
    sizes = header.cpu().tolist()
    offsets_bytes = sum(8 * sizes[2 * i] for i in range(count))
    total = offsets_bytes + sum(sizes[2 * i + 1] for i in range(count))
    if rank == src:
        host = np.empty(total, dtype=np.uint8)
        cursor = 0
        for tape in tapes:
            chunk = tape.offsets.astype(np.uint64).view(np.uint8)
            host[cursor : cursor + chunk.size] = chunk
            cursor += chunk.size
        for tape in tapes:
            host[cursor : cursor + tape.data.size] = tape.data
            cursor += tape.data.size
        payload = torch.from_numpy(host).to(backend_device)
    else:
        payload = torch.empty(total, dtype=torch.uint8, device=backend_device)
    dist.broadcast(payload, src=src, group=group)
    all_offsets = payload[:offsets_bytes].cpu().numpy().view(np.uint64)
    out, offsets_cursor, data_cursor = [], 0, offsets_bytes
    for i in range(count):
        n_offsets, n_bytes = sizes[2 * i], sizes[2 * i + 1]
        out.append((payload[data_cursor : data_cursor + n_bytes], all_offsets[offsets_cursor : offsets_cursor + n_offsets].copy()))
        offsets_cursor += n_offsets
        data_cursor += n_bytes
    return out


def gather_rows(local_rows, blocks: list[tuple[int, int]], n_columns: int, dst: int = 0, group=None):
    """Gathers each rank's contiguous (rows_block x n_columns) score slab on `dst`; returns the full matrix there
    (a torch tensor on the slabs' device) and None elsewhere. Slabs are padded to the tallest block so that one
    `gather` moves everything."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    tallest = max(end - begin for begin, end in blocks)
    padded = torch.zeros((tallest, n_columns), dtype=local_rows.dtype, device=local_rows.device)
    padded[: local_rows.shape[0]] = local_rows
    received = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, received, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([received[r][: blocks[r][1] - blocks[r][0]] for r in range(world)], dim=0)


def sharded_scores(score_rows: Callable[[int, int], "np.ndarray"], query_lengths: np.ndarray, n_columns: int,
                   dst: int = 0, device=None, group=None):
    """Row-sharded cross product: rank r calls `score_rows(begin, end)` for its block — a (end - begin) x n_columns
    array — and the slabs are gathered on `dst`. `score_rows` is the engine's shard primitive
    (`_Engine.score_rows` -> `szs_b200_scores_u64tape_rows`)."""
    import torch
    import torch.distributed as dist

    blocks = partition_rows(query_lengths, dist.get_world_size(group))
    begin, end = blocks[dist.get_rank(group)]
    local = score_rows(begin, end)
    local_tensor = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        local_tensor = local_tensor.to(device)
    return gather_rows(local_tensor, blocks, n_columns, dst=dst, group=group)

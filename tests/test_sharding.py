"""CPU suite, part 3: the multi-GPU partition logic, exercised with world_size-2 gloo process groups.
The per-rank scorer is the oracle here (no GPU in this container); on the GPU box the same code path drives the
CUDA engine (tests/test_parity_gpu.py::test_row_shards_reassemble_the_matrix covers the shard primitive itself)."""

import os
import socket

import numpy as np
import pytest

from stringzilla_b200 import sharding
from tests import support


def test_partition_rows_covers_and_balances():
    rng = np.random.default_rng(1)
    lengths = rng.integers(0, 4096, size=1000)
    for world in (1, 2, 3, 8):
        blocks = sharding.partition_rows(lengths, world)
        assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == len(lengths)
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        loads = [int(lengths[b:e].sum()) for b, e in blocks]
        assert max(loads) <= sum(loads) / world + 4096
    assert sharding.partition_rows(np.array([], dtype=np.int64), 4) == [(0, 0)] * 4
    assert sharding.partition_rows(np.array([5]), 3)[-1][1] == 1


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, seed: int):
    import torch.distributed as dist

    import stringzilla_b200 as szb

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        # Only rank 0 owns the inputs; the others receive them through the broadcast.
        queries = support.random_strings(rng, 17, 0, 120, support.DNA)
        candidates = support.random_strings(rng, 9, 0, 120, support.DNA)
        q_tape = szb.Tape(queries) if rank == 0 else None
        c_tape = szb.Tape(candidates) if rank == 0 else None
        if seed % 2:
            (q_data, q_off), (c_data, c_off) = sharding.broadcast_tapes([q_tape, c_tape])
            q_local, c_local = szb.Tape((q_data.numpy(), q_off)), szb.Tape((c_data.numpy(), c_off))
        else:
            q_data, q_off = sharding.broadcast_tape(q_tape)
            c_data, c_off = sharding.broadcast_tape(c_tape)
            q_local = szb.Tape((q_data.numpy(), q_off.numpy().astype(np.uint64)))
            c_local = szb.Tape((c_data.numpy(), c_off.numpy().astype(np.uint64)))

        def score_rows(begin, end):
            rows = [q_local[i] for i in range(begin, end)]
            return support.oracle_levenshtein(rows, c_local).astype(np.int64) if rows else np.zeros((0, 9), np.int64)

        full = sharding.sharded_scores(score_rows, q_local.lengths, len(c_local))
        if rank == 0:
            want = support.oracle_levenshtein(queries, candidates).astype(np.int64)
            assert np.array_equal(full.numpy(), want)
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("seed", [98, 99], ids=["per_tape_broadcasts", "packed_broadcast"])
def test_broadcast_shard_gather_with_gloo_world_size_2(seed):
    import torch.multiprocessing as mp

    mp.spawn(_worker, args=(2, _free_port(), seed), nprocs=2, join=True)

"""The multi-GPU entries of the C ABI (`szs_b200_partition_rows`, `szs_b200_scores_u64tape_sharded`,
`szs_b200_comm_*`): the host-side partition is checked here on CPU; the GPU tests run the threaded orchestrator with
several shards on ONE device (same code path as several devices), a communicator of one rank, and — when the box has
two GPUs — two ranks in two processes with a real NCCL broadcast and gather, always against the CPU oracle."""

from __future__ import annotations

import ctypes as ct
import os
import tempfile

import numpy as np
import pytest

from tests import support


def c_partition(lengths, parts, symmetric=False):
    from stringzilla_b200 import _abi

    lib = _abi.bind()
    lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
    bounds = np.zeros(parts + 1, dtype=np.uint64)
    assert lib.szs_b200_partition_rows(lengths.ctypes.data, lengths.size, parts, int(symmetric), bounds.ctypes.data) == 0
    return [int(b) for b in bounds]


def test_partition_rows_covers_and_balances(library_built):
    rng = np.random.default_rng(3)
    lengths = rng.integers(0, 4096, size=1000)
    for parts in (1, 2, 3, 8):
        bounds = c_partition(lengths, parts)
        assert bounds[0] == 0 and bounds[-1] == len(lengths) and all(a <= b for a, b in zip(bounds, bounds[1:]))
        loads = [int(np.maximum(lengths[a:b], 1).sum()) for a, b in zip(bounds, bounds[1:])]
        assert max(loads) <= sum(loads) / parts + 4096
    # Symmetric calls score the cells left of the diagonal: row i weighs len(i) * (len(0) + ... + len(i)).
    weights = np.maximum(lengths, 1) * np.cumsum(np.maximum(lengths, 1))
    for parts in (2, 4, 8):
        bounds = c_partition(lengths, parts, symmetric=True)
        loads = [int(weights[a:b].sum()) for a, b in zip(bounds, bounds[1:])]
        assert max(loads) <= sum(loads) / parts + int(weights.max())
        assert bounds[1] > len(lengths) // parts  # the first block is taller than an equal-rows split
    assert c_partition(np.zeros(0), 4) == [0, 0, 0, 0, 0]
    assert c_partition(np.array([5]), 3)[-1] == 1
    assert c_partition(np.zeros(10), 5) == [0, 2, 4, 6, 8, 10]  # empty strings still spread


def test_multi_gpu_symbols_refuse_bad_arguments(library_built):
    from stringzilla_b200 import _abi

    lib = _abi.bind()
    message = ct.c_char_p()
    assert lib.szs_b200_partition_rows(None, 3, 2, 0, None) != 0
    assert lib.szs_b200_scores_u64tape_sharded(0, None, None, None, None, None, 0, ct.byref(message)) != 0
    assert b"must not be null" in message.value
    assert lib.szs_b200_comm_scores_u64tape(0, None, None, None, None, None, None, 0, ct.byref(message)) != 0
    handle = ct.c_void_p()
    assert lib.szs_b200_comm_init(None, 2, 0, 0, ct.byref(handle), ct.byref(message)) != 0


# ---------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------


def _engines():
    import stringzilla_b200 as szb
    from stringzilla_b200 import scoring

    byte_to_class, blosum = scoring.blosum62()
    skewed = blosum.copy()
    skewed[3, 7] += 2  # asymmetric: the orientation of every cell matters
    return [
        ("lev_unit", szb.LevenshteinDistances(), support.DNA,
         lambda q, c: support.oracle_levenshtein(q, c)),
        ("lev_weird_affine", szb.LevenshteinDistances(1, 3, 4, 2), support.DNA,
         lambda q, c: support.oracle_levenshtein(q, c, 1, 3, 4, 2)),
        ("nw_affine", szb.NeedlemanWunschScores(byte_to_class, blosum, open=-4, extend=-1), support.PROTEIN,
         lambda q, c: support.oracle_weighted(q, c, byte_to_class, blosum, -4, -1, False)),
        ("sw_skewed", szb.SmithWatermanScores(byte_to_class, skewed, open=-4, extend=-1), support.PROTEIN,
         lambda q, c: support.oracle_weighted(q, c, byte_to_class, skewed, -4, -1, True)),
        ("nw_skewed_linear", szb.NeedlemanWunschScores(byte_to_class, skewed, open=-3, extend=-3), support.PROTEIN,
         lambda q, c: support.oracle_weighted(q, c, byte_to_class, skewed, -3, -3, False)),
    ]


@pytest.mark.gpu
def test_sharded_call_with_three_shards_on_one_device():
    import stringzilla_b200 as szb

    rng = np.random.default_rng(71)
    group = szb.DeviceGroup([0, 0, 0])
    assert len(group) == 3
    for name, engine, alphabet, oracle in _engines():
        queries = support.random_strings(rng, 41, 0, 300, alphabet)
        candidates = support.mutated_copies(rng, (queries * 2)[:29], 0.2, alphabet)
        got = engine.sharded(queries, candidates, group)
        assert np.array_equal(got, oracle(queries, candidates)), name
        # Symmetric self-similarity: shards score the cells left of the diagonal, the orchestrator mirrors them.
        got = engine.sharded(queries, None, group)
        want = oracle(queries, None)
        assert np.array_equal(got, want), name
        assert np.array_equal(got, engine(queries, None, szb.DeviceScope(gpu_device=0))), name
        # Fewer rows than shards, and a strided result matrix.
        wide = np.full((2, 40), 7, dtype=got.dtype)
        engine.sharded(queries[:2], candidates, group, out=wide[:, :29])
        assert np.array_equal(wide[:, :29], oracle(queries[:2], candidates)) and (wide[:, 29:] == 7).all(), name
    stats = _engines()[0][1].last_call_stats
    assert stats.cells == 0  # a fresh engine reports nothing


@pytest.mark.gpu
def test_communicator_of_one_rank():
    import stringzilla_b200 as szb

    rng = np.random.default_rng(72)
    scope = szb.DeviceScope(gpu_device=0)
    communicator = szb.Communicator(szb.Communicator.unique_id(), 1, 0, 0)
    for name, engine, alphabet, oracle in _engines():
        queries = support.random_strings(rng, 19, 0, 200, alphabet)
        candidates = support.random_strings(rng, 11, 0, 200, alphabet)
        assert np.array_equal(engine.communicate(communicator, scope, queries, candidates), oracle(queries, candidates)), name
        assert np.array_equal(engine.communicate(communicator, scope, queries, None), oracle(queries, None)), name
    # What rank 0 rejects travels in the exchange header (every rank returns it instead of waiting in a collective),
    # and the communicator stays usable.
    name, engine, alphabet, oracle = _engines()[0]
    queries = support.random_strings(rng, 5, 1, 50, alphabet)
    candidates = support.random_strings(rng, 4, 1, 50, alphabet)
    too_narrow = np.zeros((5, 3), dtype=np.uint64)
    with pytest.raises(szb.SimilarityError) as refusal:
        engine.communicate(communicator, scope, queries, candidates, out=too_narrow)
    assert refusal.value.status == -15 and "stride" in str(refusal.value)  # sz_unexpected_dimensions_k
    assert np.array_equal(engine.communicate(communicator, scope, queries, candidates), oracle(queries, candidates))


def _two_rank_worker(rank: int, id_path: str, seed: int):
    import time

    import stringzilla_b200 as szb

    if rank == 0:
        with open(id_path + ".tmp", "wb") as handle:
            handle.write(szb.Communicator.unique_id())
        os.replace(id_path + ".tmp", id_path)
    else:
        while not os.path.exists(id_path):
            time.sleep(0.01)
    unique_id = open(id_path, "rb").read()
    communicator = szb.Communicator(unique_id, 2, rank, rank)
    scope = szb.DeviceScope(gpu_device=rank)
    rng = np.random.default_rng(seed)
    for name, engine, alphabet, oracle in _engines():
        # Only rank 0 passes the strings; the draw is repeated on rank 1 just to keep the generators in step.
        queries = support.random_strings(rng, 37, 0, 400, alphabet)
        candidates = support.random_strings(rng, 23, 0, 400, alphabet)
        for symmetric in (False, True):
            got = engine.communicate(communicator, scope, queries if rank == 0 else None,
                                     (None if symmetric else candidates) if rank == 0 else None)
            if rank == 0:
                assert np.array_equal(got, oracle(queries, None if symmetric else candidates)), (name, symmetric)
            else:
                assert got is None
    # Rank 0's rejection of its own arguments reaches rank 1 through the header: both return it, nobody waits.
    engine = _engines()[0][1]
    with pytest.raises(szb.SimilarityError) as refusal:
        if rank == 0:
            engine.communicate(communicator, scope, queries, candidates, out=np.zeros((len(queries), 1), dtype=np.uint64))
        else:
            engine.communicate(communicator, scope)
    assert refusal.value.status == -15  # sz_unexpected_dimensions_k
    # The threaded orchestrator over two real devices, from rank 0's process alone.
    if rank == 0:
        group = szb.DeviceGroup([0, 1])
        for name, engine, alphabet, oracle in _engines():
            queries = support.random_strings(rng, 33, 0, 600, alphabet)
            assert np.array_equal(engine.sharded(queries, None, group), oracle(queries, None)), name


@pytest.mark.gpu
def test_two_ranks_exchange_over_nccl():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under `gpurun --gpus 2`)")
    import time

    with tempfile.TemporaryDirectory() as directory:
        # A rank that fails leaves its peer blocked inside a collective: never wait for the children without a deadline.
        context = mp.spawn(_two_rank_worker, args=(os.path.join(directory, "nccl_id"), 5), nprocs=2, join=False)
        deadline = time.monotonic() + 420
        try:
            while not context.join(timeout=5):
                assert time.monotonic() < deadline, "the two-rank exchange did not finish in time"
        finally:
            for process in context.processes:
                if process.is_alive():
                    process.kill()

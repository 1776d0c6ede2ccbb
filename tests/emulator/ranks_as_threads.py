#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: `szs_b200_comm_scores_u64tape` with every rank a THREAD of this process, on the emulated library and
the thread-backed NCCL stand-in (run by tests/test_emulated_library.py with SZS_B200_LIBRARY / SZS_B200_NCCL_LIBRARY set).
Checks, for worlds of 2 and 3 ranks: rectangular and symmetric calls of every engine of tests/test_multi_gpu.py against the
oracle on rank 0 (asymmetric tables included: the orientation of every cell matters), more ranks than rows, and that
arguments rank 0 rejects come back from EVERY rank instead of leaving the others in the broadcast."""
import os
import sys
import threading
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

import stringzilla_b200 as szb  # noqa: E402
from tests import support  # noqa: E402
from tests.test_multi_gpu import _engines  # noqa: E402


def rank_main(world, rank, unique_id, seed, failures):
    try:
        communicator = szb.Communicator(unique_id, world, rank, 0)
        scope = szb.DeviceScope(gpu_device=0)
        rng = np.random.default_rng(seed)
        for name, engine, alphabet, oracle in _engines():
            queries = support.random_strings(rng, 23, 0, 260, alphabet)
            candidates = support.random_strings(rng, 17, 0, 260, alphabet)
            few = queries[: world - 1]  # fewer rows than ranks: some ranks score nothing
            for q, c in ((queries, candidates), (queries, None), (few, candidates)):
                got = engine.communicate(communicator, scope, q if rank == 0 else None, c if rank == 0 else None)
                if rank == 0:
                    assert np.array_equal(got, oracle(q, c)), (name, world, c is None, len(q))
                else:
                    assert got is None
        # Rank 0's rejection of its own arguments reaches every rank through the header.
        engine = _engines()[0][1]
        try:
            if rank == 0:
                engine.communicate(communicator, scope, queries, candidates, out=np.zeros((len(queries), 1), dtype=np.uint64))
            else:
                engine.communicate(communicator, scope)
            raise AssertionError("the call should have been refused")
        except szb.SimilarityError as refusal:
            assert refusal.status == -15, refusal
        # ... and the communicator is still in step afterwards.
        got = engine.communicate(communicator, scope, queries if rank == 0 else None, candidates if rank == 0 else None)
        if rank == 0:
            assert np.array_equal(got, _engines()[0][3](queries, candidates))
    except BaseException as error:  # noqa: BLE001 - reported by the main thread
        failures.append((world, rank, repr(error)))


def main():
    for world in (2, 3):
        unique_id = szb.Communicator.unique_id()
        failures = []
        threads = [threading.Thread(target=rank_main, args=(world, rank, unique_id, 90 + world, failures), daemon=True)
                   for rank in range(world)]
        for thread in threads:
            thread.start()
        for thread in threads:
            thread.join(timeout=600)
            if thread.is_alive():
                print(f"world {world}: a rank is stuck in a collective; failures so far: {failures}", flush=True)
                os._exit(3)
        if failures:
            print(f"world {world}: {failures}", flush=True)
            os._exit(1)
        print(f"world {world}: ok", flush=True)


if __name__ == "__main__":
    main()

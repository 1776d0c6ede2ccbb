import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import stringzilla_b200 as szb
from tests import support
rng = np.random.default_rng(29)
scope = szb.DeviceScope(gpu_device=0)
for n in [1, 31, 32, 33, 64, 100, 255, 256, 257, 300, 511, 513, 600, 767, 1023, 1024, 1025, 1088, 1279, 1537, 2049, 3073, 4097, 5000, 6145, 8192, 9000, 12000, 16384, 16385]:
    cands = [support.random_strings(rng, 1, n, n, support.DNA)[0]]
    queries = support.random_strings(rng, 9, 0, 1500, support.DNA) + [b""]
    e = szb.LevenshteinDistances()
    got = e(queries, cands, device=scope)
    want = support.oracle_levenshtein(queries, cands)
    bad = np.argwhere(got != want)
    st = e.last_call_stats
    print(n, "bitpar" if st.pairs_bit_parallel else "dp", "OK" if len(bad) == 0 else f"BAD {len(bad)}/{got.size} e.g. q{bad[0][0]} len {len(queries[bad[0][0]])}: got {got[tuple(bad[0])]} want {want[tuple(bad[0])]}")

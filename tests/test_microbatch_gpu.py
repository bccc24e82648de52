"""GPU: the opt-in request combiner behind the stock VecSimIndex_TopKQuery (VECSIM_B200_MICROBATCH_US > 0): many
threads issuing single queries must each get the exact answer, while the library serves them with shared corpus passes.

The combiner itself is unit-tested on the host (tests/test_micro_batcher.py); the feature is off unless the variable
is set."""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import os, sys
import numpy as np
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import oracle_lib as ol
from redisearch_b200 import vecsim as vs

n, dim, k, T = 70_000, 128, 10, 48
rows = ol.synth_rows(ol.F32, 42, 0, n, dim)
g = vs.VecSimIndex(vs.VecSimType_FLOAT32, dim, vs.VecSimMetric_Cosine)
p = ol.PortIndex(ol.F32, dim, ol.COS, tier=ol.TIER_AVX512)
assert g.add_many(rows, label0=1) == n
p.add_many(rows, 1)
qs = ol.synth_rows(ol.F32, 43, 0, T, dim)
expect = [p.topk(q, k) for q in qs]
g.stats(reset=True)
with ThreadPoolExecutor(max_workers=T) as pool:
    for rep in range(3):
        got = list(pool.map(lambda q: g.topk(q, k), qs))
        for (gi, gs, code), (pi, ps) in zip(got, expect):
            assert code == 0 and gi.tolist() == pi.tolist()
            assert gs.astype(np.float32).tobytes() == ps.astype(np.float32).tobytes()
st = g.stats()
assert st.scan_launches < 3 * T, f"{st.scan_launches} corpus passes for {3 * T} queries: nothing was combined"
print("MICROBATCH-OK", st.scan_launches)
'''


def test_concurrent_single_queries_are_combined(tmp_path):
    script = tmp_path / "mb.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + SCRIPT)
    env = dict(os.environ, VECSIM_B200_MICROBATCH_US="3000")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0 and "MICROBATCH-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]

"""GPU: parity at (or near) the sizes BASELINE.json quotes — a different regime from the 66K-300K-row shapes of
test_vecsim_coarse.py (tens of thousands of row tiles per CTA range, the sample pass + fixed-bound main pass at scale,
multi-GB corpora).  The corpus is generated on the device, the answer of the C-ABI batch entry point is checked against the
REFERENCE's own distance kernels + heap (oracle/_ref Ref_ScanTopKChunk, or the C restatement when _ref is absent) fed with the
device's stored rows copied back from HBM (VecSimB200_ReadRows) — the same checker bench.py uses at 10M x 768."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _setup():
    import bench

    return bench, bench.Env()


def _check(bench, env, index, rows, q_stored, k, vtype_code, metric_code, picks):
    chk, _, kind = bench.reference_scan_of_device_rows(env, index, rows, 0, np.ascontiguousarray(q_stored[picks]), k, vtype_code, metric_code,
                                                       bench.usable_cores())
    return [chk.result(i) for i in range(len(picks))], kind


@pytest.mark.parametrize("metric", ["cosine", "l2"])
def test_fp32_batch_256_at_2m_rows(metric):
    bench, env = _setup()
    vs, L, torch = env.vs, env.L, env.torch
    rows, nq, k, dim = 2_000_000, 256, 10, 768
    m = vs.VecSimMetric_Cosine if metric == "cosine" else vs.VecSimMetric_L2
    index, _ = bench.build_shard(env, vs.VecSimType_FLOAT32, m, rows, 0)
    qdev = torch.empty((nq, dim), dtype=torch.float32, device=env.dev)
    assert env.S.Synth_FillRows(qdev.data_ptr(), dim * 4, 0, 43, 0, nq, dim, env.sp) == 0
    if metric == "cosine":
        assert env.S.Synth_NormalizeRowsF32(qdev.data_ptr(), dim * 4, nq, dim, env.sp) == 0
    torch.cuda.synchronize()
    out_l = torch.empty((nq, k), dtype=torch.int64, device=env.dev)
    out_s = torch.empty((nq, k), dtype=torch.float32, device=env.dev)
    for _ in range(2):
        assert L.VecSimB200_TopKQueryBatchDevice(index.h, qdev.data_ptr(), nq, k, out_l.data_ptr(), out_s.data_ptr(), env.sp) == 0
    torch.cuda.synchronize()
    assert L.VecSimB200_LastBatchPath(index.h) == 1, "the batch did not take the tensor-core coarse route"
    flags = np.zeros(nq, dtype=np.uint32)
    assert L.VecSimB200_LastCoarseFlags(index.h, flags.ctypes.data, nq) == 0
    assert (flags != 0).sum() >= nq * 0.95, np.bincount(flags, minlength=3).tolist()
    picks = list(range(0, nq, 16))
    exp, _ = _check(bench, env, index, rows, qdev.cpu().numpy(), k, ol.F32, ol.COS if metric == "cosine" else ol.L2, picks)
    dl, ds = out_l.cpu().numpy(), out_s.cpu().numpy()
    for i, q in enumerate(picks):
        assert dl[q].tolist() == exp[i][0].tolist(), (metric, q, int(flags[q]))
        assert ds[q].tobytes() == exp[i][1].astype(np.float32).tobytes()
    env.close()


def test_config3_real_shard_fp16_k100_batch1024():
    """One GPU's shard of BASELINE configs[2]: 6.25M x 768 fp16, IP, k=100, batch 1024, on the tcgen05 direct route
    (LastBatchPath == 2).  Bar (north_star): distances within 1e-2, ids identical modulo candidates within that tolerance of
    the k-th; the reference's own fp16 tier on this CPU is the checker."""
    bench, env = _setup()
    vs, L, torch = env.vs, env.L, env.torch
    rows, nq, k, dim = 6_250_000, 1024, 100, 768
    index, _ = bench.build_shard(env, vs.VecSimType_FLOAT16, vs.VecSimMetric_IP, rows, 0)
    qdev = torch.empty((nq, dim), dtype=torch.float16, device=env.dev)
    assert env.S.Synth_FillRows(qdev.data_ptr(), dim * 2, vs.VecSimType_FLOAT16, 43, 0, nq, dim, env.sp) == 0
    torch.cuda.synchronize()
    out_l = torch.empty((nq, k), dtype=torch.int64, device=env.dev)
    out_s = torch.empty((nq, k), dtype=torch.float32, device=env.dev)
    assert L.VecSimB200_TopKQueryBatchDevice(index.h, qdev.data_ptr(), nq, k, out_l.data_ptr(), out_s.data_ptr(), env.sp) == 0
    torch.cuda.synchronize()
    assert L.VecSimB200_LastBatchPath(index.h) == 2, "the batch did not take the tensor-core direct route"
    picks = [0, 257, 511, 1023]
    q_host = qdev.cpu().numpy().view(np.uint16)
    exp, _ = _check(bench, env, index, rows, q_host, k, ol.F16, ol.IP, picks)
    dl, ds = out_l.cpu().numpy(), out_s.cpu().numpy()
    for i, q in enumerate(picks):
        ref_ids, ref_sc = exp[i]
        assert np.abs(ds[q] - ref_sc.astype(np.float32)).max() <= 1e-2 * max(1.0, float(np.abs(ref_sc).max()))
        kth = float(ref_sc[-1])
        safe = {int(l) for l, s in zip(ref_ids.tolist(), ref_sc.tolist()) if s < kth - 1e-2 * max(1.0, abs(kth))}
        assert safe <= set(dl[q].tolist())
        assert len(set(dl[q].tolist()) & set(ref_ids.tolist())) >= k - 5
    env.close()

"""Row-range sharding of a FLAT corpus over ranks and the single exchange step of the KNN path.

One process per GPU (torchrun).  Each rank owns rows [lo, hi) of the corpus (labels travel with the
rows), scans its shard, and the per-shard top-k lists are exchanged with ONE all-gather and merged by
`(score asc, label asc)` — the coordinator merge of the reference (src/module.c:3139-3176, comparator
VS/utils/query_result_utils.h:19-23).  Exactness: the global top-k is a subset of the union of the local
top-k lists.  torch.distributed is plumbing only; the merge itself is a CUDA kernel
(VecSimB200_MergeShardTopK).
"""
import ctypes as C


def shard_range(n_total: int, world: int, rank: int):
    """Contiguous ranges; shard g owns rows [g*N/G, (g+1)*N/G)."""
    lo = (n_total * rank) // world
    hi = (n_total * (rank + 1)) // world
    return lo, hi


def allgather_topk(scores, labels, group=None):
    """scores [B,k] float32, labels [B,k] int64 (-1 = empty) on this rank -> ([G,B,k], [G,B,k])."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    gs = torch.empty((world,) + tuple(scores.shape), dtype=scores.dtype, device=scores.device)
    gl = torch.empty((world,) + tuple(labels.shape), dtype=labels.dtype, device=labels.device)
    if scores.is_cuda:
        dist.all_gather_into_tensor(gs, scores.contiguous(), group=group)
        dist.all_gather_into_tensor(gl, labels.contiguous(), group=group)
    else:  # gloo (CPU tests)
        ls = [torch.empty_like(scores) for _ in range(world)]
        ll = [torch.empty_like(labels) for _ in range(world)]
        dist.all_gather(ls, scores.contiguous(), group=group)
        dist.all_gather(ll, labels.contiguous(), group=group)
        gs, gl = torch.stack(ls), torch.stack(ll)
    return gs, gl


def merge_topk_device(gath_scores, gath_labels, stream_ptr=None):
    """[G,B,k] gathered lists on a CUDA device -> merged ([B,k], [B,k]) with the library's kernel."""
    import torch

    from . import vecsim

    G, B, k = gath_scores.shape
    out_s = torch.empty((B, k), dtype=torch.float32, device=gath_scores.device)
    out_l = torch.empty((B, k), dtype=torch.int64, device=gath_scores.device)
    sp = C.c_void_p(stream_ptr if stream_ptr is not None else torch.cuda.current_stream().cuda_stream)
    rc = vecsim.lib().VecSimB200_MergeShardTopK(gath_scores.data_ptr(), gath_labels.data_ptr(), G, B, k,
                                                out_s.data_ptr(), out_l.data_ptr(), sp)
    if rc != 0:
        raise RuntimeError("VecSimB200_MergeShardTopK failed")
    return out_s, out_l

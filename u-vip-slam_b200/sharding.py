"""Multi-GPU partitioning of the hot path.  One process per GPU (torch.distributed).

* frames / sequences: independent units, rank r takes a contiguous block — no collective on the data path.
* database-scale brute-force kNN (config 4): train rows are sharded contiguously, queries are replicated, every rank
  computes its local top-2 with GLOBAL train indices, the per-query (idx, dist) pairs (16 B) are all-gathered (NCCL over
  NVLink on GPUs, gloo in the CPU tests) and merged by (distance, global index) — identical for any shard count."""
import numpy as np


def shard_bounds(n, world):
    """contiguous ranges [b[r], b[r+1]) covering n rows; sizes differ by at most one"""
    return [(r * n) // world for r in range(world + 1)]


def frames_for_rank(nframes, rank, world):
    b = shard_bounds(nframes, world)
    return b[rank], b[rank + 1]


def sharded_knn2(local_top2, merge, gather, queries, train_shard, idx_base, world):
    """local_top2(queries, train_shard, idx_base) -> (idx[nq,2], dist[nq,2]) with global indices;
    gather(x) -> list of `world` arrays/tensors in rank order; merge(idx_parts, dist_parts) -> (idx, dist)."""
    idx, dist = local_top2(queries, train_shard, idx_base)
    if world == 1:
        return idx, dist
    return merge(gather(idx), gather(dist))


def merge_top2_numpy(idx_parts, dist_parts):
    """reference merge for host-side tests: lexicographic min-2 by (dist, global index); -1 marks a missing neighbour"""
    idx = np.concatenate(idx_parts, 1).astype(np.int64); dist = np.concatenate(dist_parts, 1).astype(np.int64)
    key = np.where(idx >= 0, dist * (1 << 32) + idx, np.iinfo(np.int64).max)
    order = np.argsort(key, 1, kind='stable')[:, :2]
    oi = np.take_along_axis(idx, order, 1); od = np.take_along_axis(dist, order, 1)
    miss = np.take_along_axis(key, order, 1) == np.iinfo(np.int64).max
    oi[miss] = -1; od[miss] = 257
    return oi.astype(np.int32), od.astype(np.int32)


CUDA_STREAM_LEGACY = 1      # cudaStreamLegacy: the C-ABI treats a NULL stream as "the handle's own stream", never as the default stream


def gpu_sharded_knn2(pkg, matcher, d_q, d_t_shard, idx_base, world, dist_mod=None, stream_ptr=None):
    """device path: uvip_knn2_device on the local shard, all_gather of the 2 x (nq, 2) int32 results, uvip_knn2_merge_device.
    Every step runs on ONE stream — torch's current stream unless stream_ptr names another (which must then be torch's current
    stream too, since the NCCL collectives are ordered against that one): the kernels write oi / od before the gather reads them
    and the merge runs behind the gather.  Torch's default stream has handle 0; it is passed as cudaStreamLegacy."""
    import ctypes as C
    import torch
    L = pkg.capi.lib()
    nq = d_q.shape[0]
    oi = torch.empty((nq, 2), dtype=torch.int32, device=d_q.device); od = torch.empty_like(oi)
    if stream_ptr is None:
        stream_ptr = torch.cuda.current_stream(d_q.device).cuda_stream
    sp = C.c_void_p(stream_ptr if stream_ptr else CUDA_STREAM_LEGACY)
    pkg.capi.check(L.uvip_knn2_device(matcher.h, C.c_void_p(d_q.data_ptr()), nq, C.c_void_p(d_t_shard.data_ptr()), d_t_shard.shape[0],
                                      int(idx_base), C.c_void_p(oi.data_ptr()), C.c_void_p(od.data_ptr()), sp))
    if world == 1:
        return oi, od
    gi = torch.empty((world, nq, 2), dtype=torch.int32, device=d_q.device); gd = torch.empty_like(gi)
    dist_mod.all_gather_into_tensor(gi, oi); dist_mod.all_gather_into_tensor(gd, od)
    mi = torch.empty_like(oi); md = torch.empty_like(od)
    pkg.capi.check(L.uvip_knn2_merge_device(matcher.h, C.c_void_p(gi.data_ptr()), C.c_void_p(gd.data_ptr()), world, nq * 2, nq,
                                            C.c_void_p(mi.data_ptr()), C.c_void_p(md.data_ptr()), sp))
    return mi, md

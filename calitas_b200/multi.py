"""Host-side gather for the one-process-per-GPU launch: contig-range shards are independent, so the only multi-rank step is
merging the per-rank hit lists.  There is no collective on the compute path; this runs after calitas_search has returned on every rank.

Each rank's list is ordered (guide, contig, coordinate_start, strand, -score) and the shards are contiguous, ascending base ranges, so the global
ReferenceHit.sort order (ReferenceHit.scala:276-287) per guide is ALMOST rank 0's hits of that guide, then rank 1's, ...: consecutive windows overlap
by guide length + d + g - 1 bases, so around a cut the last window of one shard and the first window of the next can both report hits whose starts
interleave (they do when -O is too large for removeOverlaps to collapse them).  The lists are therefore merged by the sort key, stably in shard
order -- which is the single engine's arrival order for equal keys.  No further de-duplication is needed (hits near a cut are resolved identically on
both sides thanks to the halo windows, and reported only by the owner).

Limit: with max_overlap <= 0 every later hit of a (contig, strand) group "overlaps" (>= 0), so the reference's sweep has unbounded reach and no
halo makes a shard's result exact; search with dedup=False then and de-duplicate the gathered hits on one host, as the C++ tool layer does
(calitas_tool_search_reference_batch with several engines)."""
import numpy as np

from ._capi import hit_dtype


def merge_shard_records(per_rank, dedup=True):
    """per_rank: list (shard order) of structured arrays from HitSet.records() -> one array in single-engine order.
    dedup=False: the lists of a search without removeOverlaps/sort (arrival order: guide, window, strand, rank) are concatenated per guide."""
    dt = hit_dtype()
    per_rank = [np.asarray(r, dtype=dt) for r in per_rank]
    if not per_rank:
        return np.zeros(0, dtype=dt)
    allr = np.concatenate(per_rank)
    if allr.size == 0:
        return allr
    rank_of = np.concatenate([np.full(r.size, i, dtype=np.int64) for i, r in enumerate(per_rank)])
    pos = np.concatenate([np.arange(r.size, dtype=np.int64) for r in per_rank])
    if not dedup:
        order = np.lexsort((pos, rank_of, allr["guide_idx"]))      # guide-major, then shard order, then each shard's own order
    else:                                                          # ReferenceHit.sort key, ties in shard order then each shard's own order
        order = np.lexsort((pos, rank_of, -allr["score"].astype(np.int64), allr["strand"], allr["guide_start_offset"], allr["contig_idx"], allr["guide_idx"]))
    return allr[order]


def gather_hits(records, group=None, dst=0):
    """torch.distributed gather of every rank's hit records to `dst` (gloo or nccl process group; call on all ranks).
    Returns the merged array on dst, None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    raw = np.ascontiguousarray(records, dtype=hit_dtype()).view(np.uint8).reshape(-1)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([raw.size], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, mine, group=group)
    sizes = [int(x) for x in sizes.tolist()]
    cap = max(1, max(sizes))
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    if raw.size:
        buf[:raw.size] = torch.from_numpy(raw.copy()).to(dev)
    out = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst, group=group)
    if rank != dst:
        return None
    return merge_shard_records([o[:n].cpu().numpy().view(hit_dtype()).reshape(-1) for o, n in zip(out, sizes)])

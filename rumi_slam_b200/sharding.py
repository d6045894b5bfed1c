"""Multi-GPU sharding of the two hot paths.  One process per GPU; torch.distributed is plumbing only.

* extraction shards by FRAME: no data-path collective, each rank extracts its own frames;
* all-pairs matching shards the TRAIN set into contiguous index ranges; every rank matches all queries against its
  shard (global train indices), the 8-byte {d1, d2, idx} candidates are all-gathered, and the merge kernel folds them
  in rank order with the reference's strict '<' rule -- identical to the single-GPU scan because every index of
  shard s precedes every index of shard s+1.
"""
import numpy as np


def frame_shard(n_frames, rank, world):
    """Contiguous [begin, end) range of the frames rank `rank` extracts."""
    base, rem = divmod(n_frames, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def train_shard(n_train, rank, world):
    return frame_shard(n_train, rank, world)


def sharded_top2(matcher, Q, T_local, t_base, group=None):
    """Q: all queries (CUDA u8 [nq,32]) replicated on every rank; T_local: this rank's train rows; t_base: global
    index of its first row.  Returns the global (idx1, d1, d2) on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    i1, d1, d2 = matcher.top2_device(Q, T_local, t_base=t_base, sync=False)
    packed = matcher.pack_device(i1, d1, d2, sync=True)
    gathered = torch.empty((world, Q.shape[0]), dtype=torch.int64, device=Q.device)
    dist.all_gather_into_tensor(gathered.view(-1), packed, group=group)
    return matcher.merge_device(gathered, world, Q.shape[0])


def merge_top2_host(parts):
    """Host restatement of the merge rule for the gloo (CPU) multi-process tests: parts = list of (idx1, d1, d2)
    numpy triples in ascending train-range order."""
    i1 = np.full_like(parts[0][0], -1)
    b1 = np.full(parts[0][1].shape, 256, np.int32)
    b2 = np.full(parts[0][1].shape, 256, np.int32)
    for (pi, p1, p2) in parts:
        p1 = p1.astype(np.int32)
        p2 = p2.astype(np.int32)
        better = p1 < b1
        b2 = np.where(better, np.minimum(b1, p2), np.minimum(b2, p1))
        i1 = np.where(better, pi, i1)
        b1 = np.where(better, p1, b1)
    return i1, b1.astype(np.uint16), b2.astype(np.uint16)

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


def init_matcher_comm(matcher, group=None):
    """Creates the matcher's own NCCL communicator over the ranks of `group`: rank 0 draws the NCCL unique id,
    torch.distributed (any backend) only carries those 128 bytes to the other ranks."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [matcher.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    matcher.comm_init(box[0], rank, world)


def sharded_top2(matcher, Q, T_local, t_base, out=None, sync=False):
    """Q: all queries (CUDA u8 [nq,32]) replicated on every rank; T_local: this rank's train rows; t_base: global
    index of its first row.  Returns the global (idx1, d1, d2) on every rank.  The whole exchange step lives in the
    C ABI (rumi_hamming_top2_sharded: scan -> packed candidates -> ncclAllGather -> fold, all on the matcher's stream,
    no host synchronisation); call init_matcher_comm once before."""
    return matcher.top2_sharded(Q, T_local, t_base, out=out, sync=sync)


def query_shard(n_query, rank, world):
    """Small problems (a shard's scan is shorter than an exchange step) shard the QUERIES instead: every rank scans
    its own query range against the whole train set with rumi_hamming_top2_device -- no exchange, no merge; the result
    stays sharded by query."""
    return frame_shard(n_query, rank, world)


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

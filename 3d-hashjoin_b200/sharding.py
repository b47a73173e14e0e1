"""Host-side plumbing of the multi-GPU join: one process per GPU, bucket-range ownership.

Every bucket -- hence every chain and every key group -- has exactly one owner
(`hj3d_owner_range`: owner = bucket // ceil(D / G)), so the ranks exchange (key, global row id) records once per
relation and then join locally; counters add, checksums add / xor.  The records come out of
`hj3d_partition_by_owner` grouped by owner; this module is the exchange and the merge, written against
`torch.distributed` only (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch

MASK64 = (1 << 64) - 1


def exchange_records(dist, part, counts, device):
    """all-to-all-v of records grouped by owner: part[(sum(counts)), 2] int32, counts[g] records for rank g.
    Returns (received records, received counts per source rank)."""
    sc = torch.tensor(counts, dtype=torch.int64, device=device)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc)
    rcounts = [int(x) for x in rc.tolist()]
    recv = torch.empty((sum(rcounts), 2), dtype=torch.int32, device=device)
    dist.all_to_all_single(recv, part, output_split_sizes=rcounts, input_split_sizes=[int(x) for x in counts])
    return recv, rcounts


def exchange_many(dist, parts, counts_list, device, recv_bufs=None):
    """The same exchange for several relations at once: ONE small collective carries the counts of all of them (one host
    synchronisation instead of one per relation), then one all-to-all-v per relation into `recv_bufs` (optional
    preallocated [capacity, 2] int32 buffers; a fresh tensor is used when the capacity does not suffice).
    Returns [(received records, received counts), ...]."""
    k = len(parts)
    sc = torch.tensor(counts_list, dtype=torch.int64, device=device).t().contiguous()      # [world, k]: row g goes to rank g
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc)
    rcounts = [[int(x) for x in col] for col in rc.t().tolist()]
    out = []
    for i in range(k):
        need = sum(rcounts[i])
        buf = recv_bufs[i] if recv_bufs is not None and recv_bufs[i] is not None and recv_bufs[i].shape[0] >= need else None
        recv = buf[:need] if buf is not None else torch.empty((need, 2), dtype=torch.int32, device=device)
        dist.all_to_all_single(recv, parts[i], output_split_sizes=rcounts[i], input_split_sizes=[int(x) for x in counts_list[i]])
        out.append((recv, rcounts[i]))
    return out


def merge_counters(dist, c, device):
    """Counters of the sharded join = counters of the unsharded one: sums (checksum_sum modulo 2^64), xor of checksum_xor."""
    world = dist.get_world_size()
    keys = ("matches", "num_cmps", "out_tuples", "checksum_sum", "checksum_xor")
    mine = torch.tensor([[c[k] & 0xFFFFFFFF, (c[k] >> 32) & 0xFFFFFFFF] for k in keys], dtype=torch.int64, device=device)
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    vals = [[int(v[i, 0]) | (int(v[i, 1]) << 32) for i in range(len(keys))] for v in allv]
    out = {}
    for i, k in enumerate(keys):
        if k == "checksum_xor":
            x = 0
            for v in vals:
                x ^= v[i]
            out[k] = x
        else:
            out[k] = sum(v[i] for v in vals) & MASK64
    return out

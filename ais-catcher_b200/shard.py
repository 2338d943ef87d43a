"""Multi-GPU plumbing for the stream batch: streams are independent, so the batch is cut into one
contiguous slice per rank, no sample ever crosses NVLink, and the only collective is a SUM of a few uint64 counters
(decoded messages, samples) plus a MAX of the device time.  One process per GPU over torch.distributed (NCCL on the
GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist

COUNTER_NAMES = ["frames_crc_ok", "messages", "samples_per_stream", "submits", "frames_dropped", "messages_a", "messages_b", "reserved"]


def stream_slice(n_streams, rank, world):
    """Streams [lo, hi) owned by `rank`: GPU g owns [g*B/G, (g+1)*B/G) (balanced to within one stream)."""
    if not (0 <= rank < world) or n_streams < 0:
        raise ValueError("bad rank/world/n_streams")
    lo = (n_streams * rank) // world
    hi = (n_streams * (rank + 1)) // world
    return lo, hi


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def gather_counts(counters, device="cpu"):
    """All-reduce (SUM) of this rank's aisgpu_counters() vector; every rank gets the job totals."""
    t = torch.tensor([int(c) for c in counters], dtype=torch.int64, device=device)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.tolist()]


def max_over_ranks(value, device="cpu"):
    """A multi-GPU time is the max over ranks of the device-measured time."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_messages(local_msgs, lo):
    """Concatenate per-rank message lists in stream order on every rank.  local_msgs: list of (local_stream, payload);
    `lo` is this rank's first global stream.  Used by the parity check of the sharded run (host side only)."""
    mine = [(s + lo, p) for s, p in local_msgs]
    if world() == 1:
        return sorted(mine, key=lambda q: q[0])
    out = [None] * world()
    dist.all_gather_object(out, mine)
    merged = [q for part in out for q in part]
    return sorted(merged, key=lambda q: q[0])  # stable: per-stream emission order is kept

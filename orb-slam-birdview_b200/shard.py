"""Multi-GPU plumbing: frame sharding and result gathering.

The per-frame path has no communication: frames (or whole sequences) are independent units, one process per
GPU owns a contiguous range, and results are gathered on the host.  `torch.distributed` is only used for the
barrier around timed regions, the max-over-ranks of a measured time and the gather of per-rank results
(NCCL on the GPU box, gloo in the CPU tests)."""
import hashlib


def shard_ranges(n_frames, world, overlap=0):
    """Contiguous frame range [start, end) per rank, sizes differing by at most one.  `overlap` extra
    predecessor frames are prepended to every range but the first, for frame-to-frame matching (frame t is
    matched against t-1, src/Tracking.cc:1227,1241): the rank then *owns* [start+overlap', end) and only
    reads the overlap.  Returns a list of (read_start, own_start, end)."""
    base, rem = divmod(n_frames, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < rem else 0)
        out.append((max(0, s - overlap), s, e))
        s = e
    return out


def sequences_to_ranks(n_sequences, world):
    """Whole sequences to ranks round-robin (C5: 256-frame sequences; sequences never split when there are
    at least as many sequences as GPUs)."""
    return [list(range(r, n_sequences, world)) for r in range(world)]


def digest(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        h.update(memoryview(a).cast("B") if hasattr(a, "dtype") else bytes(a))
    return h.hexdigest()


def gather_objects(obj, dist=None):
    """All ranks' `obj` as a list on every rank (host-side gather; no collective on the data path)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


def max_over_ranks(value, dist=None, device=None):
    """Max of a Python float over ranks (timings are reported as the slowest rank's)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])

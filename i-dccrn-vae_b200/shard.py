"""Utterance sharding across the GPUs of one box (one process per GPU, torch.distributed).

The enhancement path has no cross-utterance operation in eval mode (ComplexBatchNormal uses running
statistics, LSTM state is per sequence: SURVEY §8(e)), so inference shards the utterance batch by rank with a full
weight replica per GPU and NO data-path collective.  The only collectives are control-plane: a barrier around
timed regions, a MAX-reduce of the per-rank device time, and (optionally) an all-gather of the enhanced
waveforms for callers that want the whole batch on every rank.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, world, rank):
    """Contiguous, balanced [lo, hi) slice of ``n_items`` for ``rank`` (first n_items % world ranks get one more)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank %d/%d" % (world, rank))
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def max_over_ranks(value, device):
    """MAX over ranks of a python float (device time of a timed region)."""
    world, _ = world_info()
    if world == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def enhance_sharded(x, encoder, decoder, device, eps=None, gather=True, decoder_kwargs=None):
    """Enhance the utterances of ``x`` (B, L) that belong to this rank: x[lo:hi] -> device -> encoder -> decoder.
    Returns (waveforms, (lo, hi)); with ``gather`` every rank receives all B waveforms (all-gather of the results,
    ragged shards padded), otherwise only its own slice.  ``eps``: optional supplied eps list for the whole batch."""
    world, rank = world_info()
    B = x.shape[0]
    lo, hi = shard_bounds(B, world, rank)
    decoder_kwargs = decoder_kwargs or {}
    if hi > lo:
        xs = x[lo:hi].to(device, non_blocking=True)
        es = [e[lo:hi].to(device) for e in eps] if eps is not None else None
        with torch.no_grad():
            r = encoder(xs, train=False, eps=es)
            stft_x, z, skiper, C, F = r[-1], r[0], r[-4], r[-3], r[-2]
            sig, _ = decoder(stft_x, z, skiper, C, F, train=False, **decoder_kwargs)
    else:
        sig = torch.empty((0, 0), dtype=torch.float32, device=device)
    if not gather or world == 1:
        return sig, (lo, hi)
    S = getattr(decoder, "num_samples", 1)
    width = torch.tensor([sig.shape[1] if sig.numel() else 0], device=device)
    dist.all_reduce(width, op=dist.ReduceOp.MAX)
    per = (B + world - 1) // world * S
    buf = torch.zeros((per, int(width.item())), dtype=torch.float32, device=device)
    if sig.numel():
        buf[:sig.shape[0]] = sig
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = []
    for r_, p in enumerate(parts):
        l, h = shard_bounds(B, world, r_)
        out.append(p[:(h - l) * S])
    return torch.cat(out, 0), (lo, hi)

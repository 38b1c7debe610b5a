"""CPU tier: the N>1 host path (utterance sharding, result gather, max-over-ranks timing) with world_size 2 over
gloo, compute through the emulated C-ABI contract."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common as C
from idccrn_b200 import shard


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 5, 64, 65, 256):
        for w in (1, 2, 3, 8):
            spans = [shard.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_bounds(4, 2, 2)


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import abi_emulator
    from idccrn_b200 import lib
    lib.call = abi_emulator.call                       # test double of the C ABI (CPU tier only)
    lib.require_f32_cuda = lambda t, what: t.contiguous()
    torch.set_num_threads(2)
    B, L, seed = 3, 400, 0
    enc, dec = C.build_vae(1, 1, "skip_prepare", "real_imag", seed, "cpu")
    x, eps = C.vae_inputs(B, L, 1, 1, seed, "cpu")
    out, (lo, hi) = shard.enhance_sharded(x, enc, dec, "cpu", eps=eps, gather=True)
    t = shard.max_over_ranks(10.0 + rank, "cpu")
    if rank == 0:
        torch.save({"out": out, "t": t}, tmp)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_enhancement_matches_single_process(emulated_abi, tmp_path):
    B, L, seed = 3, 400, 0
    enc, dec = C.build_vae(1, 1, "skip_prepare", "real_imag", seed, "cpu")
    x, eps = C.vae_inputs(B, L, 1, 1, seed, "cpu")
    want = C.run_vae(enc, dec, x, eps, "skip_prepare")["recon_sig"]
    tmp = str(tmp_path / "rank0.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, tmp), nprocs=2, join=True)
    got = torch.load(tmp)
    assert got["out"].shape == want.shape
    assert C.rel_l2(got["out"], want) < 1e-6            # ragged shards (2 + 1 utterances), same arithmetic
    assert got["t"] == 11.0                             # MAX over ranks


def _train_worker(rank, world, port, tmp, overlap=True, steps=1):
    """Data-parallel training step: each rank back-propagates its own shard, FlatAdam all-reduces (mean) the flat
    gradient bucket and applies the same update on every rank."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import abi_emulator
    from idccrn_b200 import lib, ops, losses
    from idccrn_b200.train import FlatAdam
    import test_train_step as TS
    lib.call = abi_emulator.call
    lib.require_f32_cuda = lambda t, what: t.contiguous()
    ops.set_gemm_mode("tc")
    torch.set_num_threads(2)
    noisy, frozen = TS.build_step(1, 3, "cpu")
    opt = FlatAdam(noisy.parameters(), lr=1e-3, weight_decay=1e-3, process_group=dist.group.WORLD, world_size=world,
                   overlap=overlap)
    x = C.synth_waveform(2, 500, seed=50 + rank)                       # a different shard per rank
    eps = [torch.zeros(2, 1, 6, C.ZDIM)] * 2                           # the emulator has no Philox: supply eps
    with torch.no_grad():
        rc, rn = frozen[0](x, train=False, eps=eps), frozen[1](x, train=False, eps=eps)
    for _ in range(steps):
        r = noisy(x, train=True, eps=eps)
        loss, _, _ = losses.nsvae_kl_loss(r, rc, rn, C.ZDIM, 1, 1.0)
        opt.zero_grad()
        loss.backward()
        local = torch.cat([p.grad.reshape(-1) for p in opt.params if p.grad is not None]).clone()
        opt.step()
    torch.save({"local": local, "reduced": opt.gflat.clone(), "flat": opt.flat.clone(),
                "overlapped": opt.overlapped_elements}, tmp % rank)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_adam(emulated_abi, tmp_path):
    tmp = str(tmp_path / "rank%d.pt")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_train_worker, args=(2, port, tmp), nprocs=2, join=True)
    a, b = torch.load(tmp % 0), torch.load(tmp % 1)
    assert C.rel_l2(a["local"], b["local"]) > 1e-2                      # different shards, different gradients
    want = (a["local"] + b["local"]) / 2
    assert C.rel_l2(a["reduced"], want) < 1e-6 and torch.equal(a["reduced"], b["reduced"])
    assert torch.equal(a["flat"], b["flat"])                            # identical parameters after the step


def test_overlapped_allreduce_gives_the_same_update(emulated_abi, tmp_path):
    """From the second step on (the flat layout exists) the encoder's slice of the gradient bucket is all-reduced
    asynchronously as soon as its backward node has written the gradients; the parameters after two steps must equal
    those of the non-overlapped optimiser bit for bit."""
    res = {}
    for overlap in (True, False):
        tmp = str(tmp_path / ("ov%d_rank%%d.pt" % overlap))
        port = 33500 + os.getpid() % 2000 + (7 if overlap else 0)
        mp.spawn(_train_worker, args=(2, port, tmp, overlap, 2), nprocs=2, join=True)
        res[overlap] = [torch.load(tmp % r) for r in range(2)]
    assert res[True][0]["overlapped"] > 0 and res[False][0]["overlapped"] == 0
    assert torch.equal(res[True][0]["flat"], res[True][1]["flat"])
    assert torch.equal(res[True][0]["flat"], res[False][0]["flat"])
    assert torch.equal(res[True][0]["reduced"], res[False][0]["reduced"])

"""CPU tier: train.FlatAdam is a drop-in torch.optim.Adam (train_nsvae.py:L200, L318-330): same updates, param_groups
driven by torch's lr schedulers, state_dict interchangeable with torch.optim.Adam, parameters without a gradient are
skipped like torch does.  The update kernel is the emulated contract of idv_adam_step."""
import copy

import torch

from idccrn_b200.train import FlatAdam


def _params(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g)) for s in ((7, 3), (5,), (2, 4, 3), (6,))]


def _set_grads(ps, step, skip=()):
    g = torch.Generator().manual_seed(100 + step)
    for i, p in enumerate(ps):
        p.grad = None if i in skip else torch.randn(p.shape, generator=g)


def _close(a, b, tol=2e-6):
    return all(torch.allclose(x.detach(), y.detach(), rtol=tol, atol=tol) for x, y in zip(a, b))


def test_matches_torch_adam_with_scheduler_and_missing_grads(emulated_abi):
    pa, pb = _params(), _params()
    oa = FlatAdam(pa, lr=1e-2, weight_decay=1e-3)
    ob = torch.optim.Adam(pb, lr=1e-2, weight_decay=1e-3)
    assert isinstance(oa, torch.optim.Optimizer) and oa.param_groups[0]["lr"] == 1e-2
    sa = torch.optim.lr_scheduler.StepLR(oa, step_size=2, gamma=0.5)
    sb = torch.optim.lr_scheduler.StepLR(ob, step_size=2, gamma=0.5)
    # parameter 3 never gets a gradient in the first steps (the encoder's unused dense.*), parameter 1 loses its
    # gradient later (a layer frozen mid-training), parameter 3 joins late with its own step count
    plan = [(3,), (3,), (3,), (1, 3), (1,), (1,), ()]
    for step, skip in enumerate(plan):
        _set_grads(pa, step, skip)
        _set_grads(pb, step, skip)
        oa.step()
        ob.step()
        sa.step()
        sb.step()
        assert oa.param_groups[0]["lr"] == ob.param_groups[0]["lr"]
        assert _close(pa, pb), step
    oa.zero_grad()
    assert all(p.grad is None for p in pa)


def test_state_dict_round_trips_with_torch_adam(emulated_abi):
    pa, pb = _params(1), _params(1)
    oa = FlatAdam(pa, lr=3e-3, weight_decay=1e-3)
    ob = torch.optim.Adam(pb, lr=3e-3, weight_decay=1e-3)
    for step in range(3):
        _set_grads(pa, step, (3,))
        _set_grads(pb, step, (3,))
        oa.step()
        ob.step()
    sd_a, sd_b = oa.state_dict(), ob.state_dict()
    assert sorted(sd_a["state"]) == sorted(sd_b["state"]) == [0, 1, 2]
    for i in sd_b["state"]:
        assert float(sd_a["state"][i]["step"]) == float(sd_b["state"][i]["step"]) == 3.0
        assert torch.allclose(sd_a["state"][i]["exp_avg"], sd_b["state"][i]["exp_avg"], atol=1e-6)
        assert torch.allclose(sd_a["state"][i]["exp_avg_sq"], sd_b["state"][i]["exp_avg_sq"], atol=1e-7)
    # resume: torch's checkpoint into a fresh FlatAdam, FlatAdam's checkpoint into a fresh torch Adam
    pc, pd = [torch.nn.Parameter(p.detach().clone()) for p in pb], [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oc = FlatAdam(pc, lr=1.0)
    oc.load_state_dict(copy.deepcopy(sd_b))
    od = torch.optim.Adam(pd, lr=1.0)
    od.load_state_dict(copy.deepcopy(sd_a))
    assert oc.param_groups[0]["lr"] == 3e-3 and oc.param_groups[0]["weight_decay"] == 1e-3 and oc.step_count == 3
    for step in range(3, 6):
        for ps in (pa, pb, pc, pd):
            _set_grads(ps, step, (3,))
        for o in (oa, ob, oc, od):
            o.step()
        assert _close(pa, pb) and _close(pc, pb) and _close(pd, pb), step
    # loading into an optimiser that already has a layout overwrites its moments
    oa.load_state_dict(copy.deepcopy(ob.state_dict()))
    _set_grads(pa, 9, (3,))
    _set_grads(pb, 9, (3,))
    oa.step()
    ob.step()
    assert _close(pa, pb)


def test_parameters_stay_views_of_one_flat_buffer(emulated_abi):
    ps = _params(2)
    o = FlatAdam(ps, lr=1e-3)
    _set_grads(ps, 0)
    v0 = [p._version for p in ps]
    o.step()
    assert all(p._version > v for p, v in zip(ps, v0))            # weight-pack caches see the update
    base = o.flat.data_ptr()
    off = 0
    for p in ps:
        assert p.data_ptr() == base + 4 * off
        off += p.numel()
    assert o.gflat.numel() == off

"""GPU tier: every C-ABI entry point against the torch restatement of its contract (tests/abi_emulator.py)
on edge-case shapes: row counts that are not tile multiples, N tails, K chunks that are not multiples of
the K step, two-source taps, single-frame / single-utterance inputs."""
import pytest
import torch

import abi_emulator as E
import common as C
from idccrn_b200 import lib, pack

pytestmark = pytest.mark.gpu


def _both(name, args, outs):
    """args: list of CPU tensors / scalars / None; outs: indices of output tensors.  Returns max rel_l2."""
    cpu = [a.clone() if isinstance(a, torch.Tensor) else a for a in args]
    E.call(name, *cpu)
    gpu = [a.cuda() if isinstance(a, torch.Tensor) else a for a in args]
    lib.call(name, *gpu)
    torch.cuda.synchronize()
    def val(t):
        t = t.detach().cpu()
        if t.dtype == torch.bfloat16:                      # split activation: hi + lo
            f = t.view(-1).to(torch.float64)
            return f[:f.numel() // 2] + f[f.numel() // 2:]
        return t
    return max(C.rel_l2(val(gpu[i]), val(cpu[i])) for i in outs)


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


@pytest.mark.parametrize("R,Tp,N,kcs,dts,two_src,prelu,tv", [
    (300, 0, 128, [16, 32], [0, 1], False, False, 0),
    (130, 13, 64, [64, 24], [1, 0], True, True, 0),    # kc not a multiple of 16, N = 64 path, two sources
    (1, 0, 16, [8], [0], False, True, 0),               # single row, N tail (16 < 64)
    (517, 47, 384, [128, 128, 128], [0, 1, 0], True, True, 0),
    (470, 47, 128, [64, 64], [0, -1], False, True, 41),   # non-causal taps (x[t], x[t+1]), 41 of 46 frames valid
])
def test_tapgemm(R, Tp, N, kcs, dts, two_src, prelu, tv):
    F0, ld0, ld1 = 3, 136, 136
    a0 = _rand(F0, R, ld0, seed=1)
    a1 = _rand(F0, R, ld1, seed=2) if two_src else None
    taps, woff = [], 0
    for i, (kc, dt) in enumerate(zip(kcs, dts)):
        src = 1 if (two_src and i % 2) else 0
        taps.append([src, i % F0, dt, 8 if src == 0 else 4, kc, woff])
        woff += kc * N
    units = [[0, len(taps), 1, 4, N, 0], [0, 1, 0, 4, 0, 0]]
    w = _rand(woff, seed=3) * 0.1
    bias = _rand(2 * N, seed=4)
    out_ld = N + 8
    out = torch.zeros(2, R, out_ld)
    args = [a0, ld0, R * ld0, a1, ld1 if two_src else 0, R * ld1 if two_src else 0, R, Tp, w, bias, N,
            torch.tensor(units, dtype=torch.int32), torch.tensor(taps, dtype=torch.int32), 2, out, out_ld,
            R * out_ld, 1 if prelu else 0, 0.2, tv]
    assert _both("idv_tapgemm_f32", args, [14]) < 1e-5


@pytest.mark.parametrize("B,L", [(1, 300), (3, 6400), (2, 12799)])
def test_stft_istft(B, L):
    basis = pack.pack_stft_basis(512, 400, "cpu")
    T = L // 100 + 1
    x = _rand(B, L, seed=5)
    out = torch.zeros(B, 257, T, 2)
    assert _both("idv_stft_fwd", [x, B, L, basis, 512, 100, 400, out], [7]) < 1e-5
    ib, wsq = pack.pack_istft_basis(512, 400, "cpu")
    spec = _rand(B, 257, T, 2, seed=6)
    frames, y = torch.zeros(B * T, 400), torch.zeros(B, 100 * (T - 1))
    assert _both("idv_istft_fwd", [spec, B, T, ib, wsq, 512, 100, 400, frames, y], [8, 9]) < 1e-5


@pytest.mark.parametrize("causal", [1, 0])
@pytest.mark.parametrize("B,Fin,T,Cout", [(1, 257, 2, 32), (3, 33, 70, 64), (5, 17, 129, 32)])
def test_enc0(B, Fin, T, Cout, causal):
    Fout = (Fin - 1) // 2 + 1
    n = Fout * B * (T + 1) * 2 * Cout
    args = [_rand(B, Fin, T, 2, seed=7), B, Fin, T, _rand(10, 2, 2 * Cout, seed=8), _rand(2 * Cout, seed=9), Cout, 0.3,
            torch.zeros(n), 0, causal, T if causal else T - 1, None, 0]
    assert _both("idv_enc0_fwd", args, [8]) < 1e-5
    args[8], args[9] = torch.zeros(2 * n, dtype=torch.bfloat16), 1
    assert _both("idv_enc0_fwd", args, [8]) < 2e-5
    if causal:                                   # streaming: x[-1] supplied, pad rows of the output kept
        args[8], args[12], args[13] = _to_split(_rand(n, seed=31)).reshape(-1), _rand(B, Fin, 2, seed=30), 1
        assert _both("idv_enc0_fwd", args, [8]) < 2e-5


def _to_split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.stack((hi.reshape(-1), lo.reshape(-1))).contiguous()


@pytest.mark.parametrize("NB,Fin,T,p_cp,s_cp,mask,S", [(2, 9, 40, 64, 64, 1, 1), (3, 5, 7, 32, 0, 0, 2)])
def test_dec5_head(NB, Fin, T, p_cp, s_cp, mask, S):
    R, Fout = NB * (T + 1), 2 * Fin - 1
    p = _rand(Fin, R, p_cp, seed=10)
    p.view(Fin, NB, T + 1, p_cp)[:, :, 0] = 0
    skip = None
    if s_cp:
        skip = _rand(Fin, R, s_cp, seed=11)
        skip.view(Fin, NB, T + 1, s_cp)[:, :, 0] = 0
    args = [p, p_cp, skip, s_cp, 0, NB, Fin, T, _rand(10, p_cp + s_cp, 2, seed=12) * 0.1, _rand(2, seed=13), 0.25, mask,
            _rand(NB, Fout, T, 2, seed=14), torch.zeros(NB * S, Fout, T, 2), S, S - 1]
    assert _both("idv_dec5_head_fwd", args, [13]) < 1e-5
    args[0], args[2], args[4] = _to_split(p), (_to_split(skip) if skip is not None else None), 1
    assert _both("idv_dec5_head_fwd", args, [13]) < 1e-5


@pytest.mark.parametrize("NB,T,H,tv", [(1, 5, 8, 0), (3, 20, 384, 0), (64, 3, 128, 0), (5, 4, 768, 0), (3, 20, 128, 14)])
def test_lstm_recurrent_and_combine(NB, T, H, tv):
    R = NB * (T + 1)
    g = _rand(2, R, 8 * H, seed=15)
    whh = _rand(2, 4 * H, H, seed=16) / (H ** 0.5)
    hseq = torch.full((4, R, H), 7.0)
    sync = torch.zeros(2, dtype=torch.int32)
    hsplit = torch.zeros(2 * 4 * R * H, dtype=torch.bfloat16)
    hseq.view(4, NB, T + 1, H)[:, :, 1 + (tv or T):] = 0          # frames beyond the valid length are not written
    assert _both("idv_lstm_recurrent_fwd", [g, 4 * H, R * 8 * H, 8 * H, whh, NB, T, H, hseq, hsplit, sync, tv], [8, 9]) < 2e-5
    assert _both("idv_lstm_recurrent_fwd", [g, 4 * H, R * 8 * H, 8 * H, whh, NB, T, H, hseq, None, sync, tv], [8]) < 1e-5
    hs = _rand(4, R, H, seed=17)
    assert _both("idv_lstm_combine_fwd", [hs, NB, T, H, torch.zeros(NB, tv or T, H, 2), tv], [4]) < 1e-6


def test_reparam_supplied_eps():
    NB, T, z, S = 3, 11, 128, 2
    lat = _rand(NB, T, 6 * z, 2, seed=18)
    for variant in (0, 1):
        args = [lat * (1 + 9 * variant), NB, T, 6 * z, 3 * z, z, S, _rand(NB, S, T, z, seed=19), _rand(NB, S, T, z, seed=20),
                0, 0, None, variant, torch.zeros(NB * S, T, z, 2)]
        assert _both("idv_reparam_fwd", args, [13]) < 1e-5
    x = _rand(NB, 257, T, 2, seed=21)
    sc, sh = _rand(257, 2, seed=22), _rand(257, 2, seed=23)
    for ze in (0, 1):
        assert _both("idv_bin_affine", [x, NB, 257, T, sc, sh, ze, torch.zeros_like(x)], [7]) < 1e-6


@pytest.mark.parametrize("NB,C_,F,T,tv", [(2, 3, 5, 7, 0), (1, 32, 129, 33, 0), (3, 1, 4, 65, 0), (2, 5, 3, 40, 34)])
def test_layout_roundtrip(NB, C_, F, T, tv):
    """tv: valid frames (< T: the user tensor has tv frames inside a T-frame row layout)."""
    Cp = 2 * ((C_ + 7) // 8 * 8)
    Tv = tv or T
    x = _rand(NB, C_, F, Tv, 2, seed=21)
    planes = torch.full((F * NB * (T + 1) * Cp,), 3.0)
    assert _both("idv_user_to_planes", [x, NB, C_, F, T, planes, 0, tv], [5]) < 1e-7
    E.call("idv_user_to_planes", x, NB, C_, F, T, planes, 0, tv)
    assert _both("idv_planes_to_user", [planes, 0, NB, C_, F, T, torch.zeros_like(x), tv], [6]) < 1e-7
    sp = torch.full((2 * planes.numel(),), 3.0, dtype=torch.bfloat16)
    assert _both("idv_user_to_planes", [x, NB, C_, F, T, sp, 1, tv], [5]) < 1e-7
    E.call("idv_user_to_planes", x, NB, C_, F, T, sp, 1, tv)
    assert _both("idv_planes_to_user", [sp, 1, NB, C_, F, T, torch.zeros_like(x), tv], [6]) < 1e-7
    z = _rand(NB * 2, Tv, 16, 2, seed=22)
    assert _both("idv_z_to_planes", [z, NB, 2, 1, T, 16, torch.full((NB * (T + 1) * 32,), 5.0), 0, tv], [6]) < 1e-7
    assert _both("idv_z_to_planes", [z, NB, 2, 1, T, 16, torch.zeros(2 * NB * (T + 1) * 32, dtype=torch.bfloat16), 1,
                                     tv], [6]) < 1e-7
    zb = _rand(C_, 6, seed=23)
    assert _both("idv_cbn_eval_user", [x, NB, C_, F * Tv, zb, torch.zeros_like(x)], [5]) < 1e-6
    # train-mode CBN on planes: statistics over the valid frames only, in-place apply leaves the other rows alone
    for split, buf in ((0, planes), (1, sp)):
        acc = torch.zeros(C_ * 5, dtype=torch.float64)
        assert _both("idv_cbn_stats_planes", [buf, split, NB, C_, F, T, acc, tv], [6]) < 1e-6
        assert _both("idv_cbn_apply_planes", [buf.clone(), split, NB, C_, F, T, zb.reshape(-1), 1, 0.3, tv, None, 0], [0]) < 1e-6
        assert _both("idv_cbn_apply_planes", [buf.clone(), split, NB, C_, F, T, zb.reshape(-1), 1, 0.3, tv,
                                              torch.full_like(buf, 2.0), split], [10]) < 1e-6        # out of place
        other = torch.full_like(sp if not split else planes, 2.0)                     # out of place, other format
        assert _both("idv_cbn_apply_planes", [buf.clone(), split, NB, C_, F, T, zb.reshape(-1), 1, 0.3, tv, other,
                                              1 - split], [10]) < 1e-5


def test_streaming_state_kernels():
    """csrc/stream.cu against the contract: framing with the reflect start, history shift, LSTM cell step, carry,
    carried overlap-add (start-of-stream envelope and steady state)."""
    NB, k, hop, win, kpad = 3, 2, 100, 400, 448
    hist, xn = _rand(NB, win - hop, seed=40), _rand(NB, hop * k, seed=41)
    for base in (-200, -100, 0, 12300):
        fr = torch.zeros(2 * NB * k * kpad, dtype=torch.bfloat16)
        assert _both("idv_stream_frames_split", [hist, xn, NB, k, base, hop, win, kpad, fr], [8]) < 1e-7
    assert _both("idv_stream_hist_shift", [hist.clone(), xn, NB, k, hop, win], [0]) < 1e-7
    assert _both("idv_stream_hist_shift", [hist.clone(), _rand(NB, hop * 5, seed=42), NB, 5, hop, win], [0]) < 1e-7
    H, T = 128, 3
    R = NB * (T + 1)
    g_in, g_rec = _rand(2, R, 8 * H, seed=43), _rand(4, NB, 4 * H, seed=44)
    c, hs, hseq = _rand(4 * NB * H, seed=45), torch.zeros(2 * 4 * NB * H, dtype=torch.bfloat16), torch.zeros(4 * R * H)
    assert _both("idv_lstm_cell_step", [g_in, 4 * H, R * 8 * H, 8 * H, g_rec, NB, H, T, 1, c, hs, hseq], [9, 10, 11]) < 1e-5
    assert _both("idv_lstm_cell_step", [None, 0, 0, 0, g_rec, NB, H, T, 2, c, hs, None], [9, 10]) < 1e-5
    frames, wsq = _rand(NB * k, 512, seed=46), pack.pack_istft_basis(512, 400, "cpu")[1]
    for t0 in (0, 2, 4, 1000):
        acc = _rand(NB, win - hop, seed=47)
        assert _both("idv_stream_ola", [frames, 512, wsq, acc, NB, k, t0, hop, win, torch.zeros(NB, hop * k)], [3, 9]) < 1e-5
    st = _rand(NB, 257, k, 2, seed=48)
    assert _both("idv_stream_last_frame", [st, NB, 257, k, torch.zeros(NB, 257, 2)], [4]) < 1e-7


def test_carry_rows_and_counter():
    from idccrn_b200.streaming import _carry_table
    NB, Tp = 3, 4
    a = _rand(2 * 5, NB, Tp, 64, seed=50).to(torch.bfloat16).cuda()
    b = _rand(7, NB, Tp, 16, seed=51).cuda()
    wa, wb = a.clone(), b.clone()
    wa[:, :, 0], wb[:, :, 0] = wa[:, :, Tp - 1].clone(), wb[:, :, Tp - 1].clone()
    table = _carry_table([(a, 10, NB, Tp), (b, 7, NB, Tp)], "cuda")
    counter = torch.zeros(1, dtype=torch.int64).cuda()
    lib.call("idv_carry_rows", table, 2, counter)
    lib.call("idv_carry_rows", table, 2, counter)
    assert torch.equal(a, wa) and torch.equal(b, wb) and int(counter) == 2


def test_tapgemm_keeps_pad_rows_when_streaming():
    """Tp < 0: pad rows of the output are left untouched (they carry x[t-1] of the previous step)."""
    R, Tp, N = 260, 13, 128
    a0 = _to_split(_rand(1, R, 72, seed=1))
    w = _to_split(_rand(1, N, 64, seed=3) * 0.1)
    out0 = _to_split(_rand(1, R, N, seed=5)).reshape(-1)
    args = [a0, 72, 1, None, 0, 0, R, -Tp, w, 64, 1, _rand(N, seed=4), N, torch.tensor([[0, 1, 0, 0, 0, 1]], dtype=torch.int32),
            torch.tensor([[0, 0, 1, 8, 64, 0]], dtype=torch.int32), 1, out0, N, R * N, R * N, 1, 1, 0.2, 0]
    assert _both("idv_tapgemm_tc", args, [16]) < 1e-5
    a32, w32 = _rand(1, R, 72, seed=1), _rand(64 * N, seed=3) * 0.1
    args = [a32, 72, R * 72, None, 0, 0, R, -Tp, w32, _rand(N, seed=4), N, torch.tensor([[0, 1, 0, 0, 0, 0]], dtype=torch.int32),
            torch.tensor([[0, 0, 1, 8, 64, 0]], dtype=torch.int32), 1, _rand(R * N, seed=6), N, R * N, 1, 0.2, 0]
    assert _both("idv_tapgemm_f32", args, [14]) < 1e-5


@pytest.mark.parametrize("NB,C_,F,T", [(3, 32, 9, 8), (2, 5, 3, 33), (1, 256, 5, 64)])
def test_cbn_prelu_backward_kernels(NB, C_, F, T):
    """csrc/backward.cu: reduce / finalize / apply against the contract; the emulator's finalize differentiates cbn()'s
    per-channel algebra by autograd, the kernel uses the hand-derived chain."""
    Ch = (C_ + 7) // 8 * 8
    n = F * NB * (T + 1) * 2 * Ch
    y = _rand(n, seed=60) * 2 + 0.5
    g = _rand(n, seed=61)
    stats = torch.stack((_rand(C_, seed=62), _rand(C_, seed=63), 1 + _rand(C_, seed=64).abs(), 0.3 * _rand(C_, seed=65),
                         0.8 + _rand(C_, seed=66).abs()), 1).contiguous()
    zb = _rand(C_, 6, seed=67)
    acc = torch.zeros(C_ * 8, dtype=torch.float64)
    assert _both("idv_cbn_bwd_reduce", [y, 0, g, 0, NB, C_, F, T, stats, zb, 0.25, acc, 0], [11]) < 1e-6
    E.call("idv_cbn_bwd_reduce", y, 0, g, 0, NB, C_, F, T, stats, zb, 0.25, acc, 0)
    gam = [1 + 0.2 * _rand(C_, seed=68), _rand(C_, seed=69), 1 + 0.2 * _rand(C_, seed=70)]
    outs = [torch.zeros(C_ * 10)] + [_rand(C_, seed=71 + i) for i in range(5)] + [torch.ones(1, dtype=torch.float64)]
    args = [acc, float(NB * F * T), C_, stats] + gam + outs
    assert _both("idv_cbn_bwd_finalize", args, [7, 8, 9, 10, 11, 12, 13]) < 2e-5
    coef = torch.zeros(C_ * 10)
    E.call("idv_cbn_bwd_finalize", acc, float(NB * F * T), C_, stats, *gam, coef, *[torch.zeros(C_) for _ in range(5)],
           torch.zeros(1, dtype=torch.float64))
    for dy_split in (0, 1):
        dy = torch.zeros(2 * n, dtype=torch.bfloat16) if dy_split else torch.zeros(n)
        ys = _to_split(y).reshape(-1)
        assert _both("idv_cbn_bwd_apply", [ys, 1, g, 0, NB, C_, F, T, stats, zb, coef, 0.25, dy, dy_split, 0], [12]) < 2e-5


def test_lstm_backward_and_misc_kernels():
    NB, T, H = 3, 6, 64
    R = NB * (T + 1)
    P_ = _rand(4, R, 4 * H, seed=80)
    cst = torch.zeros(4 * R * H)
    assert _both("idv_lstm_scan_c", [P_, NB, T, H, cst, 0], [4]) < 1e-5
    E.call("idv_lstm_scan_c", P_, NB, T, H, cst, 0)
    dH = _rand(4, R, H, seed=81)
    dP, dstep = torch.zeros(4 * R * 4 * H), torch.zeros(2 * 4 * NB * 4 * H, dtype=torch.bfloat16)
    dc, dhr = _rand(4 * NB * H, seed=82), _rand(3 * 4 * NB * H, seed=83)
    assert _both("idv_lstm_cell_bwd_step", [P_, cst, dH, None, dc.clone(), NB, T, H, T - 1, 1, 0, dP, dstep], [4, 11, 12]) < 1e-5
    assert _both("idv_lstm_cell_bwd_step", [P_, cst, dH, dhr, dc.clone(), NB, T, H, 2, 0, 1, dP, dstep], [4, 11, 12]) < 1e-5
    assert _both("idv_lstm_cell_bwd_step", [P_, cst, dH, dhr, dc.clone(), NB, T, H, 0, 0, 3, dP, dstep], [4, 11, 12]) < 1e-5   # split-K partial planes
    dl = _rand(NB, T, H, 2, seed=84)
    assert _both("idv_lstm_combine_bwd", [dl, NB, T, H, torch.full((4 * R * H,), 3.0), 0], [4]) < 1e-7
    x = _rand(1000, 200, seed=85)
    assert _both("idv_colsum_add", [x, 1000, 130, 200, _rand(130, seed=86)], [4]) < 1e-5
    B, Fin, Tt, Cout = 2, 33, 19, 32
    Fout = (Fin - 1) // 2 + 1
    st = _rand(B, Fin, Tt, 2, seed=87)
    dY = _rand(Fout, B, Tt + 1, 2 * Cout, seed=88)
    dY[:, :, 0] = 0
    assert _both("idv_enc0_wgrad", [st, dY, B, Fin, Tt, Cout, 1, torch.zeros(20 * 2 * Cout)], [7]) < 1e-5
    n = 4096 + 12
    pr, gr, m, v = _rand(n, seed=89), _rand(n, seed=90), _rand(n, seed=91) * 0.1, _rand(n, seed=92).abs() * 0.01
    assert _both("idv_adam_step", [pr, gr, m, v, n, 1e-3, 0.9, 0.999, 1e-8, 1e-3, 7], [0, 2, 3]) < 1e-5
    pl = _rand(3, 50, 24, seed=93)
    for shift in (0, 1):
        out = torch.zeros(2 * 3 * 24 * 64, dtype=torch.bfloat16)
        assert _both("idv_planes_transpose_split", [pl, 0, 3, 50, 24, 64, shift, out], [7]) < 1e-7
        assert _both("idv_planes_transpose_split", [_to_split(pl).reshape(-1), 1, 3, 50, 24, 64, shift, out], [7]) < 1e-7
    assert _both("idv_f32_to_split", [pl, pl.numel(), torch.zeros(2 * pl.numel(), dtype=torch.bfloat16)], [2]) < 1e-7


def test_kl_loss_kernel():
    """Fused closed-form KL value + gradient against autograd of the oracle's cal_kl (incl. the |delta| <= sigma
    protection branch on both distributions)."""
    n_bt, z = 37, 128
    lat1, lat2 = _rand(n_bt, 6 * z, 2, seed=95) * 0.7, _rand(n_bt, 3 * z, 2, seed=96) * 0.7
    for ch1 in (0, 3 * z):
        acc = torch.zeros(2, dtype=torch.float64)
        dl = _rand(n_bt, 6 * z, 2, seed=97)
        assert _both("idv_kl_fwd_bwd", [lat1, 6 * z, ch1, lat2, 3 * z, 0, n_bt, z, -0.3, 1.0 / n_bt, dl, acc], [10, 11]) < 2e-5


def test_bad_arguments_return_error_codes():
    with pytest.raises(RuntimeError, match="idv_stft_fwd"):
        lib.call("idv_stft_fwd", torch.zeros(1, 100).cuda(), 1, 100, torch.zeros(4).cuda(), 512, 100, 400,
                 torch.zeros(4).cuda())          # L <= n_fft/2: reflect padding impossible
    with pytest.raises(RuntimeError, match="H"):
        lib.call("idv_lstm_recurrent_fwd", torch.zeros(4).cuda(), 0, 0, 4, torch.zeros(4).cuda(), 1, 1, 6,
                 torch.zeros(4).cuda(), None, torch.zeros(2, dtype=torch.int32).cuda(), 0)


# ---------------------------------------------------------------------------------------------------
# tensor-core tap-GEMM (tcgen05 / TMEM / TMA) against the contract incl. its split arithmetic
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("R,Tp,N,kcs,dts,two_src,prelu,out_split,tv", [
    (128, 0, 64, [64], [0], False, False, 0, 0),                # one tile, one K step, fp32 out
    (128, 0, 64, [64], [0], False, False, 1, 0),                # same, split out
    (300, 0, 128, [64, 128], [0, 1], False, True, 1, 0),        # row tail, shifted tap, 3 K steps
    (517, 47, 256, [128, 64, 128], [0, 1, 0], True, True, 1, 0),   # two sources, pad rows, BN = 256
    (1000, 0, 512, [192], [1], False, False, 0, 0),             # two N tiles, many row tiles (persistent loop)
    (40000, 641, 32, [64, 64], [1, 0], False, True, 1, 0),      # > 148 tiles: several tiles per CTA, both TMEM stages
    (470, 47, 128, [64, 64], [0, -1], False, True, 1, 41),      # non-causal taps (x[t], x[t+1]), 41 of 46 frames valid
    (1284, 642, 256, [128, 128], [-1, 0], True, True, 0, 640),  # row+1 tap at the end of the tensor (zero fill)
    (256, 2, 512, [128, 64], [1, 0], False, True, 1, 0),        # streaming-sized: few tiles -> narrow N tiles (BN 64)
    (128, 0, 1536, [192, 64], [0, 0], False, False, 0, 0),        # one LSTM step: N = 4H in 24 narrow tiles
    (10300, 103, 256, [128, 64], [1, 0], True, True, 1, 0),     # >= 148 wide tiles: BN = 256 (CTA pairs), odd row-tile count
    (6000, 0, 512, [64, 128], [0, 1], False, False, 0, 0),      # BN = 256, two N tiles, fp32 out
])
@pytest.mark.parametrize("dynamic_tiles", [0, 1, 2])
def test_tapgemm_tc(R, Tp, N, kcs, dts, two_src, prelu, out_split, tv, dynamic_tiles):
    """dynamic_tiles: 0 = static tiles (wide tiles run as CTA pairs, cta_group::2), 1 = tiles claimed from a global
    counter, 2 = static tiles with the CTA pairs switched off."""
    lib.set_option("gemm_dynamic_tiles", 1 if dynamic_tiles == 1 else 0)
    lib.set_option("gemm_cta_pairs", 0 if dynamic_tiles == 2 else 1)
    try:
        _tapgemm_tc_case(R, Tp, N, kcs, dts, two_src, prelu, out_split, tv)
    finally:
        lib.set_option("gemm_dynamic_tiles", 0)
        lib.set_option("gemm_cta_pairs", 1)


def _tapgemm_tc_case(R, Tp, N, kcs, dts, two_src, prelu, out_split, tv=0):
    F0, cp0, cp1 = 3, 264, 136
    a0 = _to_split(_rand(F0, R, cp0, seed=1))
    a1 = _to_split(_rand(2, R, cp1, seed=2)) if two_src else None
    kc_max = max(kcs)
    taps, nk = [], 0
    for i, (kc, dt) in enumerate(zip(kcs, dts)):
        src = 1 if (two_src and i % 2) else 0
        taps.append([src, i % (2 if src else F0), dt, 8, kc, i])
        nk += kc // 64
    units = [[0, len(taps), 1, 8, N, nk], [0, 1, 0, 8, 0, kcs[0] // 64]]
    w = _rand(len(taps), N, kc_max, seed=3) * 0.1
    for i, kc in enumerate(kcs):
        w[i, :, kc:] = 0
    wt = _to_split(w)
    bias = _rand(2 * N, seed=4)
    out_ld = N + 16
    n_out = 2 * R * out_ld
    out = torch.zeros(2 * n_out, dtype=torch.bfloat16) if out_split else torch.zeros(n_out)
    args = [a0, cp0, F0, a1, cp1 if two_src else 0, 2 if two_src else 0, R, Tp, wt, kc_max, len(taps), bias, N,
            torch.tensor(units, dtype=torch.int32), torch.tensor(taps, dtype=torch.int32), 2, out, out_ld,
            R * out_ld, n_out, out_split, 1 if prelu else 0, 0.2, tv]
    assert _both("idv_tapgemm_tc", args, [16]) < 1e-5


@pytest.mark.parametrize("R,Tp,N,kcs,dts,two_src,prelu,out_split,splitk", [
    (128, 0, 512, [512, 512, 256, 192], [0, 0, 0, 0], False, True, 1, 1),      # one row tile, 23 K steps over up to 23 CTAs
    (128, 0, 512, [512, 512, 256, 192], [0, 0, 0, 0], False, True, 1, 0),      # same with the switch off
    (256, -2, 256, [256, 128], [1, 0], True, True, 1, 1),                       # streaming layout, pad rows kept, two sources
    (200, 0, 1536, [384, 384], [0, 0], False, False, 0, 1),                     # fp32 out (LSTM step projection), row tail
    (128, 0, 64, [64, 64, 64], [0, 0, 0], False, False, 1, 1),                  # 3 K steps, narrow tile
])
def test_tapgemm_tc_splitk(R, Tp, N, kcs, dts, two_src, prelu, out_split, splitk):
    """idv_tapgemm_tc_splitk: the K steps of every tile divided over several CTAs (partial sums in the library's
    workspace) - same contract; run twice: the workspace must be left clean."""
    lib.set_option("gemm_splitk", splitk)
    try:
        F0, cp0, cp1 = 3, 520, 264
        a0 = _to_split(_rand(F0, R, cp0, seed=1))
        a1 = _to_split(_rand(2, R, cp1, seed=2)) if two_src else None
        kc_max = max(kcs)
        taps, nk = [], 0
        for i, (kc, dt) in enumerate(zip(kcs, dts)):
            src = 1 if (two_src and i % 2) else 0
            taps.append([src, i % (2 if src else F0), dt, 8, kc, i])
            nk += kc // 64
        units = [[0, len(taps), 1, 8, N, nk], [0, 1, 0, 8, 0, kcs[0] // 64]]
        w = _rand(len(taps), N, kc_max, seed=3) * 0.1
        for i, kc in enumerate(kcs):
            w[i, :, kc:] = 0
        out_ld = N + 16
        n_out = 2 * R * out_ld
        for rep in range(2):
            out = _to_split(_rand(n_out, seed=7)).reshape(-1) if out_split else _rand(n_out, seed=7)
            args = [a0, cp0, F0, a1, cp1 if two_src else 0, 2 if two_src else 0, R, Tp, _to_split(w), kc_max, len(taps),
                    _rand(2 * N, seed=4), None, N, torch.tensor(units, dtype=torch.int32),
                    torch.tensor(taps, dtype=torch.int32), 2, out, out_ld, R * out_ld, n_out, out_split, 1 if prelu else 0,
                    0.2, 0, min(u[5] for u in units)]
            assert _both("idv_tapgemm_tc_splitk", args, [17]) < 1e-5
    finally:
        lib.set_option("gemm_splitk", 1)


def test_tapgemm_stream_tables_equal_the_row_layout():
    """TapGemmPack.tc_stream: the same tap-GEMM on the live rows of a k-frame streaming plane set ((k + 1) * cp channels
    per stream, time taps as channel offsets) writes what the row layout with kept pad rows writes."""
    from idccrn_b200 import ops
    from idccrn_b200.ops import Planes
    NB, k, C0, C1, F, N = 5, 3, 32, 64, 4, 128
    taps = [[0, 1, 1, 0, 64, 0], [0, 2, 0, 0, 64, 64 * N], [1, 0, 1, 0, 128, 128 * N], [1, 3, 0, 0, 128, 256 * N]]
    units = [[0, 4, 0, 0, 0, 0], [1, 3, 1, 0, N, 0]]
    pk = pack.TapGemmPack(_rand(384 * N, seed=1) * 0.1, _rand(2 * N, seed=2), units, taps, N, 2, N, True, 0.1, "cuda")
    R = NB * (k + 1)
    a0 = Planes(_to_split(_rand(F, R, 64, seed=3)).reshape(-1).cuda(), NB, C0, F, k, split=True)
    a1 = Planes(_to_split(_rand(F, R, 128, seed=4)).reshape(-1).cuda(), NB, C1, F, k, split=True)
    outs = []
    for live in (True, False):
        ops.STREAM_LIVE_ROWS[0] = live
        try:
            out = _to_split(_rand(2 * R * N, seed=5)).reshape(-1).cuda()
            ops.tapgemm(pk, a0, a1, NB, k, out=out)
        finally:
            ops.STREAM_LIVE_ROWS[0] = True
        outs.append(out.view(2, -1).float().sum(0).cpu())
    assert C.rel_l2(outs[0], outs[1]) < 1e-6, C.rel_l2(outs[0], outs[1])


@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("R,Tp,N,n_units", [(2600, 13, 256, 7), (5000, 0, 128, 5), (1100, 11, 512, 3)])
def test_tapgemm_tc_tile_orders(order, R, Tp, N, n_units):
    """Both tile orders (unit-major / row-major with the units as the inner loop) on problems with many units, several
    row tiles per CTA (pairs and tails) and, for N = 512, two N tiles: every unit reads its own planes / writes its own."""
    F0, cp0, kc = n_units + 1, 136, 128
    a0 = _to_split(_rand(F0, R, cp0, seed=1))
    taps, units = [], []
    for u in range(n_units):
        taps += [[0, u, 1, 0, kc, 0], [0, u + 1, 0, 8, kc, 1 + u % 2]]
        units.append([2 * u, 2, n_units - 1 - u, 0, (u % 2) * N, 4])
    wt = _to_split(_rand(3, N, kc, seed=3) * 0.1)
    bias = _rand(2 * N, seed=4)
    n_out = n_units * R * N
    out = torch.zeros(2 * n_out, dtype=torch.bfloat16)
    args = [a0, cp0, F0, None, 0, 0, R, Tp, wt, kc, 3, bias, N, torch.tensor(units, dtype=torch.int32),
            torch.tensor(taps, dtype=torch.int32), n_units, out, N, R * N, n_out, 1, 1, 0.2, 0]
    lib.set_option("gemm_tile_order", order)
    try:
        assert _both("idv_tapgemm_tc", args, [16]) < 1e-5
    finally:
        lib.set_option("gemm_tile_order", 1)


@pytest.mark.parametrize("tma", [1, 0])
@pytest.mark.parametrize("R,Tp,N,tv,pairs", [(2600, 13, 128, 0, 1), (1111, 11, 64, 7, 1), (130, 13, 256, 0, 0), (4160, 65, 192, 60, 1)])
def test_tapgemm_tc_tma_store_epilogue(tma, R, Tp, N, tv, pairs):
    """Split-bf16 outputs leave through shared memory + tensor stores (gemm_tma_store = 1, the default) or through
    per-thread stores (0): same values.  Units at column offsets 0 / N of rows twice as wide (the dense layer's layout),
    row tails that are not multiples of the 32-row store boxes, valid-frame masks, and a first-frame bias (b2 entry)."""
    F0, cp0, kc = 3, 136, 128
    a0 = _to_split(_rand(F0, R, cp0, seed=1))
    taps = [[0, 0, 1, 0, kc, 0], [0, 1, 0, 8, kc, 1], [0, 2, 0, 0, kc, 2], [0, 1, 1, 8, kc, 0]]
    units = [[0, 2, 0, 0, 0, 4], [2, 2, 0, N, N, 4], [1, 2, 1, N, 0, 4], [0, 1, 1, 0, N, 2]]
    wt = _to_split(_rand(3, N, kc, seed=3) * 0.1)
    bias, bias1 = _rand(2 * N, seed=4), _rand(2 * N, seed=5)
    ld = 2 * N
    n_out = 2 * R * ld
    out = torch.full((2 * n_out,), 3.0, dtype=torch.bfloat16)
    base = [a0, cp0, F0, None, 0, 0, R, Tp, wt, kc, 3, bias]
    tail = [N, torch.tensor(units, dtype=torch.int32), torch.tensor(taps, dtype=torch.int32), 4, out, ld, R * ld, n_out, 1, 1,
            0.2, tv]
    lib.set_option("gemm_tma_store", tma)
    lib.set_option("gemm_cta_pairs", pairs)
    try:
        assert _both("idv_tapgemm_tc", base + tail, [16]) < 1e-5
        assert _both("idv_tapgemm_tc_b2", base + [bias1] + tail, [17]) < 1e-5
    finally:
        lib.set_option("gemm_tma_store", 1)
        lib.set_option("gemm_cta_pairs", 1)


@pytest.mark.parametrize("NB,T,zdim,latent_num,S,tv,split,keep_pad", [
    (3, 7, 128, 1, 1, 0, 1, 0), (2, 5, 128, 2, 3, 0, 1, 0), (5, 9, 16, 2, 1, 6, 0, 0), (1, 3, 128, 1, 10, 0, 1, 0),
    (4, 2, 128, 1, 1, 0, 1, 1)])
def test_latent_fused(NB, T, zdim, latent_num, S, tv, split, keep_pad):
    """idv_latent_fwd = lstm_combine + reparam (per latent) + z_to_planes (per sample) of the contract emulator;
    keep_pad: the pad rows of the z planes (3.0 here) survive."""
    H = 3 * zdim * latent_num
    R = NB * (T + 1)
    Tv = tv if 0 < tv < T else T
    hseq = _rand(4, R, H, seed=31) * 0.5
    eps = [_rand(NB, S, Tv, zdim, seed=40 + i) for i in range(4)]
    latent = torch.zeros(NB, Tv, H, 2)
    z0, z1 = torch.zeros(NB * S, Tv, zdim, 2), (torch.zeros(NB * S, Tv, zdim, 2) if latent_num == 2 else None)
    n_pl = S * R * 2 * ((zdim + 7) // 8 * 8)
    zpl = torch.full((2 * n_pl,), 3.0, dtype=torch.bfloat16) if split else torch.full((n_pl,), 3.0)
    args = [hseq, NB, T, H, tv, zdim, latent_num, S, eps[0], eps[1], eps[2] if latent_num == 2 else None,
            eps[3] if latent_num == 2 else None, 0, 0, None, latent, z0, z1, zpl, split, keep_pad]
    outs = [15, 16, 18] + ([17] if latent_num == 2 else [])
    # the split planes of several samples are compared sample by sample (hi | lo halves per sample)
    cpu = [a.clone() if isinstance(a, torch.Tensor) else a for a in args]
    E.call("idv_latent_fwd", *cpu)
    gpu = [a.cuda() if isinstance(a, torch.Tensor) else a for a in args]
    lib.call("idv_latent_fwd", *gpu)
    torch.cuda.synchronize()
    for i in outs:
        a, b = gpu[i].cpu(), cpu[i]
        if a.dtype == torch.bfloat16:
            a, b = a.view(S, 2, -1).double().sum(1), b.view(S, 2, -1).double().sum(1)
        assert C.rel_l2(a, b) < 2e-6, (i, C.rel_l2(a, b))


def test_latent_fused_philox_draws_are_independent():
    """No eps supplied: N(0,1) draws; consecutive draw counters must not share a uniform (the radius uniform of one
    draw used to be the angle uniform of the next: |eps| correlated across draws), ranks / latents differ."""
    NB, T, zdim = 4, 50, 128
    H, R = 3 * zdim, 4 * 51
    hseq = torch.zeros(4, R, H).cuda()                     # mu = 0, log sigma = 0, delta = 0 -> z = eps / sqrt(2)-ish scale
    def draw(offset, seed=1234):
        latent = torch.zeros(NB, T, H, 2).cuda()
        z0 = torch.zeros(NB, T, zdim, 2).cuda()
        zpl = torch.zeros(R * 2 * zdim).cuda()
        lib.call("idv_latent_fwd", hseq, NB, T, H, 0, zdim, 1, 1, None, None, None, None, seed, offset, None, latent, z0,
                 None, zpl, 0, 0)
        torch.cuda.synchronize()
        return z0.cpu().double()
    a, b, c = draw(1), draw(2), draw(1)
    assert torch.equal(a, c)                               # same (seed, counter) -> same draw
    assert abs(float(a.mean())) < 0.02 and 0.4 < float(a[..., 0].var()) < 0.6      # sigma = 1, delta = 0: var = 1/2
    ra, rb = a.pow(2).sum(-1).flatten(), b.pow(2).sum(-1).flatten()
    corr = float(torch.corrcoef(torch.stack((ra, rb)))[0, 1])
    assert abs(corr) < 0.02, corr                          # |eps|^2 of consecutive draws uncorrelated
    assert abs(float(torch.corrcoef(torch.stack((a[..., 0].flatten(), b[..., 1].flatten())))[0, 1])) < 0.02
    d = draw(1, seed=99)
    assert C.rel_l2(d, a) > 0.5
    # the stand-alone kernel (streaming, reparameterization()) follows the same rule
    lat = torch.zeros(NB, T, H, 2).cuda()
    def draw2(offset):
        z = torch.zeros(NB, T, zdim, 2).cuda()
        lib.call("idv_reparam_fwd", lat, NB, T, H, 0, zdim, 1, None, None, 1234, offset, None, 0, z)
        torch.cuda.synchronize()
        return z.cpu().double()
    x, y = draw2(5), draw2(6)
    corr = float(torch.corrcoef(torch.stack((x.pow(2).sum(-1).flatten(), y.pow(2).sum(-1).flatten())))[0, 1])
    assert abs(corr) < 0.02, corr


@pytest.mark.parametrize("NB,T,H,tv,split", [(3, 7, 128, 0, 1), (2, 9, 20, 5, 0)])
def test_lstm_combine_planes(NB, T, H, tv, split):
    R = NB * (T + 1)
    Tv = tv if 0 < tv < T else T
    hseq = _rand(4, R, H, seed=33)
    latent = torch.zeros(NB, Tv, H, 2)
    n = R * 2 * ((H + 7) // 8 * 8)
    pl = torch.full((2 * n,), 3.0, dtype=torch.bfloat16) if split else torch.full((n,), 3.0)
    assert _both("idv_lstm_combine_planes", [hseq, NB, T, H, tv, latent, pl, split], [5, 6]) < 2e-6


@pytest.mark.parametrize("R,Tp,out_split", [(300, 13, 1), (1300, 0, 0)])
def test_tapgemm_tc_columns_wrap_into_planes(R, Tp, out_split):
    """N > out_ld: column n of a unit goes to plane out_f + n / out_ld (two output planes of a narrow transposed conv
    as one tile); the plane past the end of the tensor (odd plane count) is not written."""
    F0, cp0, N, ld, kc, n_planes = 3, 136, 128, 64, 64, 5
    a0 = _to_split(_rand(F0, R, cp0, seed=1))
    taps = [[0, 0, 0, 8, kc, 0], [0, 1, 1, 8, kc, 1], [0, 2, 1, 0, kc, 2], [0, 1, 0, 0, kc, 1]]
    units = [[0, 2, 0, 0, 0, 2], [2, 2, 2, 0, 0, 2], [1, 2, 4, 0, 0, 2]]      # planes (0,1), (2,3), (4, -)
    wt = _to_split(_rand(3, N, kc, seed=3) * 0.1)
    bias = _rand(N, seed=4)
    n_out = n_planes * R * ld
    out = torch.full((2 * n_out,), 3.0, dtype=torch.bfloat16) if out_split else torch.full((n_out,), 3.0)
    args = [a0, cp0, F0, None, 0, 0, R, Tp, wt, kc, 3, bias, N, torch.tensor(units, dtype=torch.int32),
            torch.tensor(taps, dtype=torch.int32), 3, out, ld, R * ld, n_out, out_split, 1, 0.2, 0]
    assert _both("idv_tapgemm_tc", args, [16]) < 1e-5


@pytest.mark.parametrize("NB,T,H,tv", [(3, 6, 128, 0), (64, 40, 384, 0), (5, 9, 768, 0), (70, 5, 128, 0), (3, 12, 128, 7),
                                       (130, 4, 768, 0), (200, 3, 384, 2)])      # row groups that are not co-resident: consecutive launches
def test_lstm_recurrent_tc(NB, T, H, tv):
    from idccrn_b200 import pack as PK
    n_cols, n_ctas = lib.lstm_tc_config(H)
    assert (n_cols, n_ctas) == E._lstm_tc_config(H)
    R = NB * (T + 1)
    g = _rand(2, R, 8 * H, seed=15)
    whh = _rand(2, 4 * H, H, seed=16) / (H ** 0.5)
    wp = PK.pack_lstm_whh_tc({"weight_hh_l0": whh[0]}, {"weight_hh_l0": whh[1]}, 0, n_cols, n_ctas, "cpu")
    n_rg = (NB + 63) // 64
    hseq = torch.zeros(4, R, H)
    hsplit = torch.zeros(2 * 4 * R * H, dtype=torch.bfloat16)
    hx = torch.zeros(n_rg * 2 * 2 * 2 * 128 * H, dtype=torch.bfloat16)
    sync = torch.zeros(n_rg * 2, dtype=torch.int32)
    args = [g, 4 * H, R * 8 * H, 8 * H, wp, NB, T, H, hseq, hsplit, hx, sync, tv]
    assert _both("idv_lstm_recurrent_tc", args, [8, 9]) < 2e-5


@pytest.mark.parametrize("pairs", [1, 0])
@pytest.mark.parametrize("NB,T,H,tv", [(3, 6, 128, 0), (64, 30, 768, 0), (5, 9, 768, 4), (64, 12, 384, 0), (130, 4, 768, 0),
                                       (70, 5, 512, 0)])
def test_lstm_layer_pair_tc(NB, T, H, tv, pairs):
    """One nn.LSTM layer per launch on the CTA-pair kernel (48 gate columns per CTA at H = 768), both as pairs and as
    its one-CTA-per-tile cooperative fallback."""
    from idccrn_b200 import pack as PK
    n_cols, n_ctas, work_bytes = lib.lstm_layer_pair_config(H)
    assert n_cols == (48 if H == 768 else 64) and n_ctas * (n_cols // 4) == H
    R = NB * (T + 1)
    g = _rand(2, R, 8 * H, seed=15)
    whh = _rand(2, 4 * H, H, seed=16) / (H ** 0.5)
    wp = PK.pack_lstm_whh_tc({"weight_hh_l0": whh[0]}, {"weight_hh_l0": whh[1]}, 0, n_cols, n_ctas, "cpu")
    hseq = torch.zeros(4, R, H)
    hsplit = torch.zeros(2 * 4 * R * H, dtype=torch.bfloat16)
    work = torch.zeros(work_bytes, dtype=torch.uint8)
    sync = torch.zeros(384, dtype=torch.int32)
    args = [g, 4 * H, R * 8 * H, 8 * H, wp, NB, T, H, hseq, hsplit, work, sync, tv]
    lib.set_option("lstm_wave_cta_pairs", pairs)
    try:
        assert _both("idv_lstm_layer_pair_tc", args, [8, 9]) < 2e-5
    finally:
        lib.set_option("lstm_wave_cta_pairs", 1)


@pytest.mark.parametrize("chunk_sync,publish,sync_mode", [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 2, 0), (1, 2, 2), (1, -1, 1)])
@pytest.mark.parametrize("NB,T,H", [(64, 20, 768), (100, 7, 768), (9, 12, 384)])
def test_lstm_layer_pair_tc_switches(NB, T, H, chunk_sync, publish, sync_mode):
    """Every combination of the publish / poll switches of the one-layer kernel (per-K-chunk counters, how h(t) is
    written, release forms) computes the same thing; the two-layer wavefront kernel under the same switches as well."""
    lib.set_option("lstm_chunk_sync", chunk_sync)
    lib.set_option("lstm_tma_publish", publish)
    lib.set_option("lstm_sync_mode", sync_mode)
    try:
        test_lstm_layer_pair_tc(NB, T, H, 0, 1)
        if H == 384:
            test_lstm2_wave_tc(NB if NB <= 64 else 64, T, H, 0)
    finally:
        lib.set_option("lstm_chunk_sync", 1)
        lib.set_option("lstm_tma_publish", -1)
        lib.set_option("lstm_sync_mode", 0)


@pytest.mark.parametrize("NB,Fin,T,two_src,mask,S", [(2, 9, 40, True, 2, 1), (3, 5, 130, False, 1, 2)])
def test_tapgemm_tc_head(NB, Fin, T, two_src, mask, S):
    """Last decoder layer + head fused in the tensor-core epilogue, packed by pack.pack_dec5_tc."""
    from idccrn_b200 import pack as PK
    R, Fout = NB * (T + 1), 2 * Fin - 1
    kcs = [64, 64] if two_src else [64]

    def planes(seed, cp):
        x = _rand(Fin, NB, T + 1, cp, seed=seed)
        x[:, :, 0] = 0
        return _to_split(x)
    p = planes(10, 64)
    skip = planes(11, 64) if two_src else None
    w10 = _rand(10, sum(kcs), 2, seed=12) * 0.1
    hp = PK.pack_dec5_tc(w10, _rand(2, seed=13), 0.25, Fin, kcs, "cpu")
    predict = torch.zeros(NB * S, Fout, T, 2)
    args = [p, 64, Fin, skip, 64 if two_src else 0, Fin if two_src else 0, R, T + 1, hp["wt"], hp["kc_max"],
            hp["n_slots"], hp["bias"], 32, hp["units"], hp["taps"], hp["n_units"], None, 0, 0, 0, 0, 1, hp["slope"],
            mask, Fout, S, S - 1, _rand(NB, Fout, T, 2, seed=14), predict, 0]
    assert _both("idv_tapgemm_tc_head", args, [28]) < 1e-5
    # and the packed form agrees with the SIMT head kernel's contract on the same folded weights
    ref = torch.zeros(NB * S, Fout, T, 2)
    E.call("idv_dec5_head_fwd", p, 64, skip, 64 if two_src else 0, 1, NB, Fin, T, w10, hp["bias"][:2].clone(), 0.25,
           1 if mask == 2 else 0, args[27], ref, S, S - 1)
    cpu = torch.zeros(NB * S, Fout, T, 2)
    a2 = list(args)
    a2[28] = cpu
    E.call("idv_tapgemm_tc_head", *a2)
    assert C.rel_l2(cpu, ref) < 2e-5


@pytest.mark.parametrize("NB,T,H,tv", [(3, 9, 128, 0), (64, 30, 384, 0), (17, 2, 384, 0), (4, 16, 128, 10),
                                       (70, 7, 128, 0), (130, 5, 384, 3)])       # > 64 utterances: chunks of 64 inside the entry point
def test_lstm2_wave_tc(NB, T, H, tv):
    from idccrn_b200 import pack as PK
    n_cols, n_ctas, work_bytes = lib.lstm2_wave_config(H)
    R = NB * (T + 1)
    g = _rand(2, R, 8 * H, seed=15)
    mods = []
    for mi in range(2):
        mods.append({"weight_hh_l0": _rand(4 * H, H, seed=20 + mi) / (H ** 0.5),
                     "weight_ih_l1": _rand(4 * H, H, seed=22 + mi) / (H ** 0.5),
                     "weight_hh_l1": _rand(4 * H, H, seed=24 + mi) / (H ** 0.5),
                     "bias_ih_l1": _rand(4 * H, seed=26 + mi) * 0.1, "bias_hh_l1": _rand(4 * H, seed=28 + mi) * 0.1})
    w0 = PK.pack_lstm_whh_tc(mods[0], mods[1], 0, n_cols, n_ctas, "cpu")
    wi = PK.pack_lstm_whh_tc(mods[0], mods[1], 1, n_cols, n_ctas, "cpu", "ih")
    w1 = PK.pack_lstm_whh_tc(mods[0], mods[1], 1, n_cols, n_ctas, "cpu")
    b1 = PK.pack_lstm_bias_tc(mods[0], mods[1], 1, n_cols, n_ctas, "cpu")
    hseq = torch.zeros(4, R, H)
    work = torch.zeros(work_bytes, dtype=torch.uint8)
    sync = torch.zeros(384, dtype=torch.int32)
    args = [g, 4 * H, R * 8 * H, 8 * H, w0, wi, w1, b1, NB, T, H, hseq, work, sync, tv]
    assert _both("idv_lstm2_wave_tc", args, [11]) < 2e-5


@pytest.mark.parametrize("NB,T,H,tv", [(4, 40, 384, 0), (8, 9, 384, 0), (1, 3, 384, 0), (1, 1, 384, 0), (2, 2, 384, 0),
                                       (5, 33, 128, 20), (8, 12, 256, 0), (3, 700, 384, 0),
                                       (9, 11, 384, 0), (16, 40, 128, 0), (17, 9, 384, 0), (32, 21, 384, 0), (32, 30, 128, 7),
                                       (24, 5, 256, 0), (100, 6, 128, 0), (70, 4, 384, 3)])
def test_lstm2_cluster_tc(NB, T, H, tv):
    """Small-batch cluster recurrence (DSMEM exchange) against the contract it shares with the wavefront kernel."""
    from idccrn_b200 import pack as PK
    upc, cs, work_bytes = lib.lstm2_cluster_config(H, NB, T)
    R = NB * (T + 1)
    g = _rand(2, R, 8 * H, seed=15)
    mods = []
    for mi in range(2):
        mods.append({"weight_hh_l0": _rand(4 * H, H, seed=20 + mi) / (H ** 0.5),
                     "weight_ih_l1": _rand(4 * H, H, seed=22 + mi) / (H ** 0.5),
                     "weight_hh_l1": _rand(4 * H, H, seed=24 + mi) / (H ** 0.5),
                     "bias_ih_l1": _rand(4 * H, seed=26 + mi) * 0.1, "bias_hh_l1": _rand(4 * H, seed=28 + mi) * 0.1})
    w0 = PK.pack_lstm_cluster_tc(mods[0], mods[1], 0, upc, cs, "cpu")
    wi = PK.pack_lstm_cluster_tc(mods[0], mods[1], 1, upc, cs, "cpu", "ih")
    w1 = PK.pack_lstm_cluster_tc(mods[0], mods[1], 1, upc, cs, "cpu")
    b1 = PK.pack_lstm_cluster_bias(mods[0], mods[1], 1, upc, cs, "cpu")
    hseq = torch.zeros(4, R, H)
    work = torch.zeros(work_bytes, dtype=torch.uint8)
    sync = torch.zeros(128 * -(-NB // 8), dtype=torch.int32)
    args = [g, 4 * H, R * 8 * H, 8 * H, w0, wi, w1, b1, NB, T, H, hseq, work, sync, tv]
    assert _both("idv_lstm2_cluster_tc", args, [11]) < 2e-5


@pytest.mark.parametrize("interleave", [1, 0])
@pytest.mark.parametrize("NB,T,H", [(128, 9, 384), (70, 12, 128), (200, 4, 384)])
def test_lstm2_wave_tc_two_interleaved_chunks(NB, T, H, interleave):
    """More than 64 utterances: two chunks of 64 run as two interleaved recurrences inside one launch
    (lstm_interleave = 1, default) or as consecutive launches (0): same numbers either way, and the same as the contract."""
    lib.set_option("lstm_interleave", interleave)
    try:
        test_lstm2_wave_tc(NB, T, H, 0)
        if H == 384:
            test_lstm_layer_pair_tc(NB, T, 768, 0, 1)
    finally:
        lib.set_option("lstm_interleave", 1)


@pytest.mark.parametrize("B,L", [(1, 300), (3, 6400), (2, 12799)])
def test_stft_istft_tensor_core_pieces(B, L):
    """Operand preparation kernels, the STFT epilogue mode and the overlap-add against their contracts."""
    from idccrn_b200 import pack as PK
    T = L // 100 + 1
    R = B * T
    hp = PK.pack_stft_tc(512, 400, "cpu")
    x = _rand(B, L, seed=5)
    frames = torch.zeros(2 * R * hp["kpad"], dtype=torch.bfloat16)
    assert _both("idv_stft_frames_split", [x, B, L, 512, 100, 400, hp["kpad"], None, frames], [8]) < 1e-7
    lens = torch.tensor([max(257, L - 173 * b) for b in range(B)], dtype=torch.int32)          # ragged batch
    assert _both("idv_stft_frames_split", [x, B, L, 512, 100, 400, hp["kpad"], lens, frames], [8]) < 1e-7
    E.call("idv_stft_frames_split", x, B, L, 512, 100, 400, hp["kpad"], None, frames)
    out = torch.zeros(B, 257, T, 2)
    args = [frames, hp["kpad"], 1, None, 0, 0, R, T, hp["wt"], hp["kc_max"], 1, hp["bias"], hp["N"], hp["units"],
            hp["taps"], 1, None, 0, 0, 0, 0, 0, 0.0, 3, 257, 1, 0, None, out, 0]
    assert _both("idv_tapgemm_tc_head", args, [28]) < 1e-5
    # the same launch also writing the split-bf16 activation rows of the first encoder layer (causal row layout)
    ld, col0 = PK.ENC0_ROWS_LD, PK.ENC0_COL0
    erows = torch.zeros(2 * B * (T + 1) * ld, dtype=torch.bfloat16)
    args2 = list(args)
    args2[16], args2[17], args2[19], args2[26] = erows, ld, B * (T + 1) * ld, col0
    assert _both("idv_tapgemm_tc_head", args2, [16, 28]) < 1e-5
    # ... and the first encoder layer as a tap-GEMM on those rows against the SIMT kernel's contract on the spectrum
    E.call("idv_tapgemm_tc_head", *args2)
    Cout = 32
    w20 = _rand(10, 2, 2 * Cout, seed=8)
    want = torch.zeros(129 * B * (T + 1) * 2 * Cout)
    wre = w20[:, 0, :Cout].t().reshape(Cout, 1, 5, 2).contiguous()       # undo pack_enc0's layout: re-in -> re-out = conv_re
    wim = w20[:, 0, Cout:].t().reshape(Cout, 1, 5, 2).contiguous()       #                          re-in -> im-out = conv_im
    pk = PK.pack_enc0_tc(wre, torch.zeros(Cout), wim, torch.zeros(Cout), None, 0.3, 257, "cpu")
    tc = pk.tc()
    Rp = B * (T + 1)
    got = torch.zeros(2 * 129 * Rp * 64, dtype=torch.bfloat16)
    a = [erows, ld, 1, None, 0, 0, Rp, T + 1, tc["wt"], tc["kc_max"], tc["n_slots"], tc["bias"], tc["N"], tc["units"], tc["taps"],
         tc["n_units"], got, 64, Rp * 64, 129 * Rp * 64, 1, 1, 0.3, T]
    assert _both("idv_tapgemm_tc", a, [16]) < 1e-5
    # same numbers as the SIMT contract evaluated with the complex weights (w_re, w_im) the pack was built from
    w_ref = torch.zeros(10, 2, 2 * Cout)
    w_ref[:, 0, :Cout], w_ref[:, 0, Cout:] = w20[:, 0, :Cout], w20[:, 0, Cout:]
    w_ref[:, 1, :Cout], w_ref[:, 1, Cout:] = -w20[:, 0, Cout:], w20[:, 0, :Cout]
    E.call("idv_enc0_fwd", out, B, 257, T, w_ref, torch.zeros(2 * Cout), Cout, 0.3, want, 0, 1, T, None, 0)
    E.call("idv_tapgemm_tc", *a)
    gf = got.view(2, -1).double().sum(0)
    assert C.rel_l2(gf, want) < 2e-5
    ip = PK.pack_istft_tc(512, 400, "cpu")
    spec = _rand(B, 257, T, 2, seed=6)
    rows = torch.zeros(2 * R * ip["kpad"], dtype=torch.bfloat16)
    assert _both("idv_spec_rows_split", [spec, B, 257, T, ip["kpad"], rows], [5]) < 1e-7
    fr = _rand(R, 512, seed=7)
    assert _both("idv_ola_fwd", [fr, 512, ip["wsq"], B, T, 512, 100, 400, None, torch.zeros(B, 100 * (T - 1))], [9]) < 1e-5
    assert _both("idv_ola_fwd", [fr, 512, ip["wsq"], B, T, 512, 100, 400, lens, torch.full((B, 100 * (T - 1)), 3.0)], [9]) < 1e-5


@pytest.mark.parametrize("n,w", [(1, (1.0, 1.0)), (257 * 33 * 2, (0.3, 0.7)), (100003, (1.0, 0.0)), (4099, (0.0, 0.0))])
def test_spec_loss(n, w):
    pred, ori = _rand(n, 2, seed=40), _rand(n, 2, seed=41)
    pred[: n // 7] = 0.0                                       # |pred| = 0: the 1e-6 under the root keeps d|pred| finite
    d_pred = _rand(n, 2, seed=42) * 0.01                       # accumulates (+=)
    acc = torch.full((2,), 0.25, dtype=torch.float64)
    assert _both("idv_spec_loss_fwd_bwd", [pred, ori, n, w[0], w[1], 1.0 / 66, d_pred, acc], [6, 7]) < 2e-5
    acc2 = torch.zeros(2, dtype=torch.float64)
    assert _both("idv_spec_loss_fwd_bwd", [pred, ori, n, w[0], w[1], 1.0 / 66, None, acc2], [7]) < 2e-5


@pytest.mark.parametrize("B,L", [(1, 300), (3, 6400), (5, 1999)])
def test_sisnr_and_ola_backward(B, L):
    src, est = _rand(B, L, seed=30), _rand(B, L, seed=31) * 0.5 + 0.3 * _rand(B, L, seed=30)
    d_est = _rand(B, L, seed=32) * 0.01                       # accumulates (+=)
    sums, loss = torch.zeros(B * 3, dtype=torch.float64), torch.full((1,), 0.5, dtype=torch.float64)
    assert _both("idv_sisnr_fwd_bwd", [src, est, B, L, 0.7, d_est, sums, loss], [5, 6, 7]) < 2e-5
    from idccrn_b200 import pack as PK
    hop = 100
    Ls = L // hop * hop
    T = Ls // hop + 1
    wsq = PK.pack_istft_basis(512, 400, "cpu")[1]
    dsig = _rand(B, Ls, seed=33)
    for ld in (400, 448):
        dfr = torch.full((B * T, ld), 7.0)
        assert _both("idv_ola_bwd", [dsig, wsq, B, T, 512, hop, 400, ld, dfr], [8]) < 1e-6
    # adjoint identity: <ola(f), s> == <f, ola_bwd(s)>
    fr = _rand(B * T, 448, seed=34).cuda()
    out = torch.zeros(B, Ls).cuda()
    lib.call("idv_ola_fwd", fr, 448, wsq.cuda(), B, T, 512, hop, 400, None, out)
    dfr = torch.zeros(B * T, 448).cuda()
    lib.call("idv_ola_bwd", dsig.cuda(), wsq.cuda(), B, T, 512, hop, 400, 448, dfr)
    lhs, rhs = float((out.double() * dsig.cuda().double()).sum()), float((fr[:, :400].double() * dfr[:, :400].double()).sum())
    assert abs(lhs - rhs) < 1e-6 * max(1.0, abs(lhs))


@pytest.mark.parametrize("NB,F,T,mask,with_dpred", [(2, 17, 9, 0, False), (3, 257, 13, 1, True), (1, 5, 1, 1, False)])
def test_head_backward(NB, F, T, mask, with_dpred):
    raw = _rand(NB, F, T, 2, seed=40)
    zb = torch.tensor([0.9, 0.3, -0.2, 1.1, 0.05, -0.1])
    stft_x = _rand(NB, F, T, 2, seed=41)
    ld = 2 * F + 6
    drows = _rand(NB * T, ld, seed=42)
    dpred = _rand(NB, F, T, 2, seed=43) if with_dpred else None
    R = NB * (T + 1)
    yp, gp = torch.full((F * R * 16,), 3.0), torch.full((F * R * 16,), 3.0)
    args = [raw, zb, 0.25, mask, stft_x if mask else None, drows, ld, dpred, NB, F, T, yp, gp]
    assert _both("idv_head_bwd", args, [11, 12]) < 2e-5


@pytest.mark.parametrize("NB,Fin,T,cp,cs", [(2, 5, 7, 64, 64), (1, 9, 1, 64, 0), (3, 129, 6, 64, 64), (2, 4, 5, 16, 32)])
def test_dec5_backward(NB, Fin, T, cp, cs):
    Fout, R, ktot = 2 * Fin - 1, NB * (T + 1), cp + cs
    dy = _rand(Fout, NB, T + 1, 16, seed=50)
    dy[:, :, 0] = 0                                            # pad rows of a gradient are zero
    w10 = _rand(10, ktot, 2, seed=51)
    for (k_off, c) in ((0, cp), (cp, cs)):
        if c == 0:
            continue
        dx = torch.full((Fin * R * c,), 5.0)
        assert _both("idv_dec5_dgrad", [dy, w10, ktot, k_off, c, Fin, NB, T, dx], [8]) < 1e-5
        x = _rand(Fin, NB, T + 1, c, seed=52 + k_off)
        x[:, :, 0] = 0
        xs = torch.zeros(2 * x.numel(), dtype=torch.bfloat16)
        E._wr(xs, 1, x)
        dW = _rand(10, ktot, 2, seed=53)                       # accumulates (+=)
        assert _both("idv_dec5_wgrad", [xs, 1, dy, ktot, k_off, c, Fin, NB, T, dW], [9]) < 1e-5
        dW = torch.zeros(10, ktot, 2)
        assert _both("idv_dec5_wgrad", [x, 0, dy, ktot, k_off, c, Fin, NB, T, dW], [9]) < 1e-5


def test_reparam_backward():
    NB, T, zdim, Htot, ch0 = 3, 11, 128, 768, 384
    lat = _rand(NB, T, Htot, 2, seed=60) * 0.7
    er, ei = _rand(NB, 1, T, zdim, seed=61), _rand(NB, 1, T, zdim, seed=62)
    dz = _rand(NB, T, zdim, 2, seed=63)
    dlat = _rand(NB, T, Htot, 2, seed=64) * 0.1               # accumulates (+=)
    assert _both("idv_reparam_bwd", [lat, NB, T, Htot, ch0, zdim, er, ei, dz, dlat], [9]) < 2e-5

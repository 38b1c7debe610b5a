"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A functional, state_dict-driven restatement (plain PyTorch, CPU, fp32 or fp64) of the enhancement
forward path of iris1997jiatong/I-DCCRN-VAE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this file; the product
package never does (its ops raise if the CUDA library is missing).

Pinning: the reference ships no tests or golden vectors, so this port is pinned against
the *reference itself*: ``oracle/make_golden.py`` imports /root/reference in the build container,
runs the reference modules on seeded inputs/weights/eps and (a) asserts this port matches them,
(b) commits the reference's outputs to ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py``
re-checks the port against those files everywhere.

Every function cites the reference lines it restates (paths relative to /root/reference).
Third-party arithmetic (torch.stft / conv2d / conv_transpose2d / LSTM / linear) is called exactly
like the reference calls it; ``stft_dense`` / ``istft_dense`` restate the two FFT calls as explicit
DFT matrices so the kernels' formulation is itself checked on the CPU.
"""
import math

import torch
import torch.nn.functional as F

EPS_CBN = 1e-5      # model/complex_progress.py:L101
EPS_REPARAM = 1e-6  # model/pvae_module.py:L2175


# ------------------------------------------------------------------------------------------------
# STFT / iSTFT
# ------------------------------------------------------------------------------------------------
def stft(signal, n_fft=512, hop=100, win=400):
    """model/pvae_module.py:L21-27 — torch.stft, periodic Hann, center/reflect, onesided."""
    w = torch.hann_window(win, dtype=signal.dtype, device=signal.device)
    spec = torch.stft(signal, n_fft=n_fft, hop_length=hop, win_length=win, window=w,
                      return_complex=True)
    return torch.view_as_real(spec)                         # (B, n_fft/2+1, T, 2)


def istft(spec_c, n_fft=512, hop=100, win=400):
    """model/pvae_module.py:L38-42 — torch.istft on a complex (B, F, T) spectrum."""
    w = torch.hann_window(win, dtype=spec_c.real.dtype, device=spec_c.device)
    return torch.istft(spec_c, n_fft=n_fft, hop_length=hop, win_length=win, window=w,
                       return_complex=False)


def padded_window(n_fft=512, win=400, dtype=torch.float64):
    wp = torch.zeros(n_fft, dtype=dtype)
    off = (n_fft - win) // 2
    wp[off:off + win] = torch.hann_window(win, dtype=dtype)
    return wp


def stft_dense(signal, n_fft=512, hop=100, win=400):
    """reflect-pad, frame, windowed dense DFT (what the CUDA kernel computes)."""
    dt = signal.dtype
    xp = F.pad(signal[:, None, :], (n_fft // 2, n_fft // 2), mode="reflect")[:, 0]
    frames = xp.unfold(1, n_fft, hop)                       # (B, T, n_fft)
    n = torch.arange(n_fft, dtype=torch.float64)
    k = torch.arange(n_fft // 2 + 1, dtype=torch.float64)
    ang = 2 * math.pi * torch.outer(n, k) / n_fft
    wp = padded_window(n_fft, win)
    cr = (torch.cos(ang) * wp[:, None]).to(dt)
    ci = (-torch.sin(ang) * wp[:, None]).to(dt)
    return torch.stack((frames @ cr, frames @ ci), dim=-1).permute(0, 2, 1, 3).contiguous()


def istft_dense(spec_ri, n_fft=512, hop=100, win=400):
    """per-frame inverse DFT * window, overlap-add, / sum(w^2), trim n_fft/2."""
    dt = spec_ri.dtype
    B, Fb, T, _ = spec_ri.shape
    n = torch.arange(n_fft, dtype=torch.float64)
    k = torch.arange(Fb, dtype=torch.float64)
    ck = torch.full((Fb,), 2.0, dtype=torch.float64)
    ck[0] = 1.0
    ck[-1] = 1.0
    ang = 2 * math.pi * torch.outer(k, n) / n_fft
    wp = padded_window(n_fft, win)
    br = (ck[:, None] / n_fft * torch.cos(ang) * wp[None, :]).to(dt)
    bi = (-ck[:, None] / n_fft * torch.sin(ang) * wp[None, :]).to(dt)
    xr = spec_ri[..., 0].permute(0, 2, 1)                   # (B, T, F)
    xi = spec_ri[..., 1].permute(0, 2, 1)
    frames = xr @ br + xi @ bi                              # (B, T, n_fft)
    total = n_fft + hop * (T - 1)
    y = torch.zeros(B, total, dtype=dt)
    env = torch.zeros(total, dtype=torch.float64)
    for t in range(T):
        y[:, t * hop:t * hop + n_fft] += frames[:, t]
        env[t * hop:t * hop + n_fft] += wp * wp
    half = n_fft // 2
    return y[:, half:total - half] / env[half:total - half].to(dt)


# ------------------------------------------------------------------------------------------------
# complex primitives
# ------------------------------------------------------------------------------------------------
def complex_conv2d(x, sd, pre, stride=(2, 1), padding=(2, 1), causal=True):
    """model/complex_progress.py:L16-22 (causal) / L32-36: four real convs, +- combine,
    causal variant drops the last time column."""
    wr, br = sd[pre + "conv_re.weight"], sd[pre + "conv_re.bias"]
    wi, bi = sd[pre + "conv_im.weight"], sd[pre + "conv_im.bias"]
    xr, xi = x[..., 0], x[..., 1]
    re = F.conv2d(xr, wr, br, stride, padding) - F.conv2d(xi, wi, bi, stride, padding)
    im = F.conv2d(xi, wr, br, stride, padding) + F.conv2d(xr, wi, bi, stride, padding)
    if causal:
        re, im = re[..., :-1], im[..., :-1]
    return torch.stack((re, im), dim=-1)


def complex_conv_transpose2d(x, sd, pre, stride=(2, 1), padding=(2, 0), causal=True):
    """model/complex_progress.py:L244-250 (causal) / L275-279."""
    wr, br = sd[pre + "tconv_re.weight"], sd[pre + "tconv_re.bias"]
    wi, bi = sd[pre + "tconv_im.weight"], sd[pre + "tconv_im.bias"]
    xr, xi = x[..., 0], x[..., 1]
    re = F.conv_transpose2d(xr, wr, br, stride, padding) - F.conv_transpose2d(xi, wi, bi, stride, padding)
    im = F.conv_transpose2d(xi, wr, br, stride, padding) + F.conv_transpose2d(xr, wi, bi, stride, padding)
    if causal:
        re, im = re[..., :-1], im[..., :-1]
    return torch.stack((re, im), dim=-1)


def cbn_whiten_affine(sd, pre):
    """Per-channel 2x2 matrix Z and the quantities of ``cbn`` — model/complex_progress.py:L168-205."""
    vrr, vri, vii = sd[pre + "Vrr"], sd[pre + "Vri"], sd[pre + "Vii"]      # (1,C,1,1)
    C = vrr.shape[1]
    tau = vrr + vii
    delta = torch.clamp(vrr * vii - vri ** 2 + EPS_CBN, min=1e-8)
    s = torch.sqrt(delta)
    t = torch.sqrt(tau + 2 * s + EPS_CBN)
    ist = 1.0 / (s * t + EPS_CBN)
    wrr, wii, wri = (vii + s) * ist, (vrr + s) * ist, -vri * ist
    g_rr = sd[pre + "gamma_rr"].view(1, C, 1, 1)
    g_ri = sd[pre + "gamma_ri"].view(1, C, 1, 1)
    g_ii = sd[pre + "gamma_ii"].view(1, C, 1, 1)
    zrr = g_rr * wrr + g_ri * wri
    zri = g_rr * wri + g_ri * wii
    zir = g_ri * wrr + g_ii * wri
    zii = g_ri * wri + g_ii * wii
    return zrr, zri, zir, zii


def cbn_eval(x, sd, pre):
    """ComplexBatchNormal.forward(train=False) — model/complex_progress.py:L161-166 + cbn."""
    C = x.shape[1]
    rc = x[..., 0] - sd[pre + "running_mean_real"]
    ic = x[..., 1] - sd[pre + "running_mean_imag"]
    zrr, zri, zir, zii = cbn_whiten_affine(sd, pre)
    out_r = zrr * rc + zri * ic + sd[pre + "beta_r"].view(1, C, 1, 1)
    out_i = zir * rc + zii * ic + sd[pre + "beta_i"].view(1, C, 1, 1)
    return torch.stack((out_r, out_i), dim=-1)


def cbn_train(x, sd, pre):
    """ComplexBatchNormal.forward(train=True) — model/complex_progress.py:L131-160 + cbn: batch mean / biased
    (co)variances over (B, F, T), eps added to Vrr and Vii, differentiable through the statistics exactly like the
    reference (running-buffer updates are a side effect the port does not reproduce)."""
    C = x.shape[1]
    re, im = x[..., 0], x[..., 1]
    rc = re - re.mean((0, 2, 3), keepdim=True)
    ic = im - im.mean((0, 2, 3), keepdim=True)
    st = dict(sd)
    st[pre + "Vrr"] = (rc * rc).mean((0, 2, 3), keepdim=True) + EPS_CBN
    st[pre + "Vii"] = (ic * ic).mean((0, 2, 3), keepdim=True) + EPS_CBN
    st[pre + "Vri"] = (rc * ic).mean((0, 2, 3), keepdim=True)
    zrr, zri, zir, zii = cbn_whiten_affine(st, pre)
    out_r = zrr * rc + zri * ic + sd[pre + "beta_r"].view(1, C, 1, 1)
    out_i = zir * rc + zii * ic + sd[pre + "beta_i"].view(1, C, 1, 1)
    return torch.stack((out_r, out_i), dim=-1)


def prelu(x, sd, pre):
    """nn.PReLU() with one shared slope on the 5-D tensor — model/pvae_module.py:L58,L67."""
    return F.prelu(x, sd[pre + "weight"])


def encoder_block(x, sd, pre, causal=True, train=False):
    """Encoder.forward(x, train) — model/pvae_module.py:L64-68."""
    pad = (2, 1) if causal else (2, 0)
    y = complex_conv2d(x, sd, pre + "conv.", (2, 1), pad, causal)
    return prelu((cbn_train if train else cbn_eval)(y, sd, pre + "bn."), sd, pre + "prelu.")


def decoder_block(x, sd, pre, causal=True, train=False):
    """Decoder.forward(x, train) — model/pvae_module.py:L88-93 (if_bn always True)."""
    y = complex_conv_transpose2d(x, sd, pre + "transconv.", (2, 1), (2, 0), causal)
    return prelu((cbn_train if train else cbn_eval)(y, sd, pre + "bn."), sd, pre + "prelu.")


def _lstm_module(sd, pre, input_size, hidden, layers, dtype):
    sub = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    dev = next(iter(sub.values())).device
    m = torch.nn.LSTM(input_size=input_size, hidden_size=hidden, num_layers=layers).to(device=dev, dtype=dtype)
    m.load_state_dict(sub)
    return m.eval()


def _lstm_functional(x, sd, pre, hidden, layers):
    """nn.LSTM (unidirectional, zero initial state) written out with the tensors of ``sd`` so that autograd reaches
    them (training-step oracle).  Gate order i, f, g, o; x: (T, B, D)."""
    T, B, _ = x.shape
    inp = x
    for l in range(layers):
        wih, whh = sd[pre + "weight_ih_l%d" % l], sd[pre + "weight_hh_l%d" % l]
        b = sd[pre + "bias_ih_l%d" % l] + sd[pre + "bias_hh_l%d" % l]
        gin = inp @ wih.t() + b
        h = x.new_zeros(B, hidden)
        c = x.new_zeros(B, hidden)
        outs = []
        for t in range(T):
            a = gin[t] + h @ whh.t()
            i, f, g, o = a[:, :hidden], a[:, hidden:2 * hidden], a[:, 2 * hidden:3 * hidden], a[:, 3 * hidden:]
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        inp = torch.stack(outs)
    return inp


def complex_lstm(x, sd, pre, hidden, layers=2, grad=False):
    """ComplexLSTM.forward — model/complex_progress.py:L58-74.  x: (T, B, D, 2).  grad=True evaluates the same
    recurrences functionally on the tensors of ``sd`` (differentiable w.r.t. them)."""
    d_in = x.shape[2]
    if grad:
        run_re = lambda v: _lstm_functional(v, sd, pre + "lstm_re.", hidden, layers)
        run_im = lambda v: _lstm_functional(v, sd, pre + "lstm_im.", hidden, layers)
        rr, ri, ii, ir = run_re(x[..., 0]), run_im(x[..., 0]), run_im(x[..., 1]), run_re(x[..., 1])
        return torch.stack((rr - ii, ir + ri), dim=-1)
    lre = _lstm_module(sd, pre + "lstm_re.", d_in, hidden, layers, x.dtype)
    lim = _lstm_module(sd, pre + "lstm_im.", d_in, hidden, layers, x.dtype)
    with torch.no_grad():
        rr, _ = lre(x[..., 0])
        ri, _ = lim(x[..., 0])
        ii, _ = lim(x[..., 1])
        ir, _ = lre(x[..., 1])
    return torch.stack((rr - ii, ir + ri), dim=-1)


def complex_dense(x, sd, pre):
    """ComplexDense.forward — model/complex_progress.py:L83-89 (no cross terms)."""
    re = F.linear(x[..., 0], sd[pre + "linear_read.weight"], sd[pre + "linear_read.bias"])
    im = F.linear(x[..., 1], sd[pre + "linear_imag.weight"], sd[pre + "linear_imag.bias"])
    return torch.stack((re, im), dim=-1)


def reparameterize(miu, log_sigma, delta, eps_r, eps_i, num_samples):
    """reparameterization — model/pvae_module.py:L2177-2231 (== L1832-1886).
    miu/log_sigma/delta: (B,T,H,2); eps_r/eps_i: (B,S,T,H) supplied instead of randn_like."""
    e = EPS_REPARAM
    mr, mi = miu[..., 0], miu[..., 1]
    sig = torch.exp(log_sigma[..., 0])
    dr, di = delta[..., 0], delta[..., 1]
    ad = torch.sqrt(dr ** 2 + di ** 2 + e)
    tmp = sig * 0.99 / (ad + e)
    clamp = ad >= (sig - 1e-3)
    dr = torch.where(clamp, dr * tmp, dr)
    di = torch.where(clamp, di * tmp, di)
    ad = torch.sqrt(dr ** 2 + di ** 2 + e)
    den = torch.sqrt(2 * (sig + dr) + e)
    num_r = sig + dr
    sx = di / (den + e)
    sy = torch.sqrt(sig ** 2 - ad ** 2 + e) / (den + e)
    den, num_r, sx, sy = (v.unsqueeze(1) for v in (den, num_r, sx, sy))
    mr = mr.unsqueeze(1).repeat(1, num_samples, 1, 1)
    mi = mi.unsqueeze(1).repeat(1, num_samples, 1, 1)
    zr = mr + (num_r / (den + e)) * eps_r
    zi = mi + sx * eps_r + sy * eps_i
    B, S, T, H = zr.shape
    return torch.stack((zr.reshape(B * S, T, H), zi.reshape(B * S, T, H)), dim=-1)


def mask_head(mask, stft_x, num_samples=1):
    """recon_type == 'mask' — model/pvae_module.py:L2594-2609 (== DCCRN_ L224-234).
    mask: (B*S,1,F,T,2); stft_x: (B,F,T,2) -> complex (B*S,F,T)."""
    m_r, m_i = mask[..., 0], mask[..., 1]
    mag = torch.tanh(torch.sqrt(m_r ** 2 + m_i ** 2))
    ph = torch.atan2(m_i / (mag + 1e-8), m_r / (mag + 1e-8))
    b, fr, t, d = stft_x.shape
    sx = stft_x.unsqueeze(1).repeat(1, num_samples, 1, 1, 1).view(b * num_samples, fr, t, d).unsqueeze(1)
    in_mag = torch.sqrt(sx[..., 0] ** 2 + sx[..., 1] ** 2)
    in_ph = torch.arctan2(sx[..., 1], sx[..., 0])
    pred = in_mag * mag * torch.exp(1j * (in_ph + ph))
    return pred.squeeze(1)


# ------------------------------------------------------------------------------------------------
# model graphs
# ------------------------------------------------------------------------------------------------
def encoder_stack(x5, sd, n_layers=6, causal=True, train=False):
    skiper = []
    for i in range(n_layers):
        x5 = encoder_block(x5, sd, "encoders.%d." % i, causal, train)
        skiper.append(x5)
    return x5, skiper


def cal_kl(miu1, miu2, log_sigma1, log_sigma2, delta1, delta2, zdim, eps=1e-10):
    """standard_nsvae_loss_true_kl.cal_kl — model/nsvae_loss.py:L275-328: closed-form KL between two complex
    Gaussians per (B, T) (distribution 1 = the noisy encoder's posterior, 2 = the frozen target's)."""
    m1r, m1i, m2r, m2i = miu1[..., 0], miu1[..., 1], miu2[..., 0], miu2[..., 1]
    s1, s2 = torch.exp(log_sigma1[..., 0]), torch.exp(log_sigma2[..., 0])

    def protect(d, s):
        dr, di = d[..., 0], d[..., 1]
        ad = torch.sqrt(dr.pow(2) + di.pow(2) + eps)
        tmp = s * 0.99 / (ad + eps)
        cl = ad >= (s - 1e-3)
        dr, di = torch.where(cl, dr * tmp, dr), torch.where(cl, di * tmp, di)
        return dr, di, dr.pow(2) + di.pow(2)
    d1r, d1i, a1 = protect(delta1, s1)
    d2r, d2i, a2 = protect(delta2, s2)
    log_det_c1 = torch.log(0.25 * (s1.pow(2) - a1) + eps)
    log_det_c2 = torch.log(0.25 * (s2.pow(2) - a2) + eps)
    coeff = 2 / (s2.pow(2) - a2 + eps)
    trace_term = s1 * s2 - d2r * d1r - d2i * d1i
    dr, di = m2r - m1r, m2i - m1i
    quadra = dr.pow(2) * (s2 - d2r) - 2 * d2i * dr * di + di.pow(2) * (s2 + d2r)
    return 0.5 * torch.sum(coeff * (trace_term + quadra) + log_det_c2 - log_det_c1, dim=2) - zdim


def nsvae_kl_loss(noisy, clean, noise, zdim=128, latent_num=1, alpha=1.0):
    """standard_nsvae_loss_true_kl.kl_loss — model/nsvae_loss.py:L330-347 (the phase-1 training loss of
    train_nsvae.py:L539-544 with w_kl = 1, w_dismiu = 0).  noisy / clean / noise: encoder state dicts of
    vae_encoder_forward.  Returns (loss, kl_clean, kl_noise)."""
    kc = cal_kl(noisy["miu_speech"], clean["miu_speech"], noisy["log_sigma_speech"], clean["log_sigma_speech"],
                noisy["delta_speech"], clean["delta_speech"], zdim)
    if latent_num == 1:
        kn = cal_kl(noisy["miu_speech"], noise["miu_speech"], noisy["log_sigma_speech"], noise["log_sigma_speech"],
                    noisy["delta_speech"], noise["delta_speech"], zdim)
        return kc.mean() - alpha * kn.mean(), kc.mean(), kn.mean()
    kn = cal_kl(noisy["miu_noise"], noise["miu_speech"], noisy["log_sigma_noise"], noise["log_sigma_speech"],
                noisy["delta_noise"], noise["delta_speech"], zdim)
    return kc.mean() + alpha * kn.mean(), kc.mean(), kn.mean()


def si_snr(source, estimate, eps=1e-8):
    """two_phase_loss.si_snr — model/nsvae_loss.py:L877-889.  The reference builds the B x B matrix
    ``estimate @ source.T`` and keeps its diagonal; the per-utterance dot product is the same number."""
    dot = (estimate * source).sum(1, keepdim=True)
    s_target = dot * source / ((source ** 2).sum(1, keepdim=True) + eps)
    e_noise = estimate - s_target
    snr = 10 * torch.log10((s_target ** 2).sum(1) / ((e_noise ** 2).sum(1) + eps) + eps)
    return -snr.mean()


def multi_recon_loss(predict, ori_stft, source, estimate, weights):
    """two_phase_loss.multi_recon_loss — model/nsvae_loss.py:L891-913 (``ori_mag`` uses real^2 + real^2 like L899).
    predict: complex (B,F,T); ori_stft: (B,F,T,2).  Returns (final, loss_cpx, loss_mag, loss_sisnr)."""
    pr, pi = predict.real, predict.imag
    p_mag = torch.sqrt(pr ** 2 + pi ** 2 + 1e-6)
    o_r, o_i = ori_stft[..., 0], ori_stft[..., 1]
    o_mag = torch.sqrt(o_r ** 2 + o_r ** 2 + 1e-6)
    l_cpx = (((pr - o_r) ** 2).sum(1) + ((pi - o_i) ** 2).sum(1)).mean()
    l_mag = ((p_mag - o_mag) ** 2).sum(1).mean()
    l_si = si_snr(source, estimate)
    return weights[0] * l_cpx + weights[1] * l_mag + weights[2] * l_si, l_cpx, l_mag, l_si


def vae_encoder_forward(sd, signal, zdim=128, latent_num=1, num_samples=1, eps=None, causal=True,
                        stft_params=(512, 100, 400), train=False, grad=False):
    """nsvae_pvae_dccrn_encoder_twophase.forward(x, train=False) — model/pvae_module.py:L2233-2268;
    with latent_num == 1 it is also pvae_dccrn_encoder_skip_prepare.forward (L1888-1914).
    eps: list of (B,S,T,zdim) tensors in draw order [speech_r, speech_i, (noise_r, noise_i)].
    Returns a dict of every stage."""
    st = {}
    stft_x = stft(signal, *stft_params)
    st["stft_x"] = stft_x
    x, skiper = encoder_stack(stft_x.unsqueeze(1), sd, 6, causal, train)
    st["skiper"] = skiper
    B, C, Fq, T, D = x.shape
    lstm_in = x.reshape(B, -1, T, D).permute(2, 0, 1, 3)
    hidden = 3 * zdim * latent_num
    lat = complex_lstm(lstm_in, sd, "lstms.0.", hidden, 2, grad).permute(1, 0, 2, 3)   # (B,T,hidden,2)
    st["latent"] = lat
    z = zdim
    st["miu_speech"], st["log_sigma_speech"], st["delta_speech"] = lat[:, :, 0:z], lat[:, :, z:2 * z], lat[:, :, 2 * z:3 * z]
    st["z_speech"] = reparameterize(st["miu_speech"], st["log_sigma_speech"], st["delta_speech"],
                                    eps[0], eps[1], num_samples)
    if latent_num == 2:
        st["miu_noise"], st["log_sigma_noise"], st["delta_noise"] = lat[:, :, 3 * z:4 * z], lat[:, :, 4 * z:5 * z], lat[:, :, 5 * z:6 * z]
        st["z_noise"] = reparameterize(st["miu_noise"], st["log_sigma_noise"], st["delta_noise"],
                                       eps[2], eps[3], num_samples)
    st["C"], st["F"] = C, Fq
    return st


def vae_decoder_forward(sd, stft_x, z, skiper, C, Fq, num_samples=1, recon_type="real_imag",
                        skip_mode="zero", skip_to_use=(0, 1, 2, 3, 4, 5), causal=True,
                        stft_params=(512, 100, 400), train=False):
    """pvae_dccrn_decoder_skip_prepare.forward (skip_mode='zero', model/pvae_module.py:L2082-2122)
    and nsvae_pvae_dccrn_decoder_twophase.forward(pad='zero'|'sig', use_sc=True, L2548-2619)."""
    BS, T, zdim, D = z.shape
    dense_out = complex_dense(z.reshape(BS * T, -1, D), sd, "dense.")
    p = dense_out.reshape(BS, T, C, Fq, D).permute(0, 2, 3, 1, 4)
    outs = [p]
    for i in range(6):
        if i in skip_to_use:
            sk = skiper[len(skiper) - i - 1]
            if skip_mode == "zero":
                sk = torch.zeros((BS,) + tuple(sk.shape[1:]), dtype=p.dtype, device=p.device)
            else:
                sk = sk.unsqueeze(1).repeat(1, num_samples, 1, 1, 1, 1).view((BS,) + tuple(sk.shape[1:]))
            p = torch.cat([p, sk], dim=1)
        p = decoder_block(p, sd, "decoders.%d." % i, causal, train)
        outs.append(p)
    if recon_type == "real_imag":
        predict = torch.complex(p[..., 0], p[..., 1]).squeeze(1)
    else:
        predict = mask_head(p, stft_x, num_samples)
    sig = istft(predict, *stft_params)
    return {"recon_sig": sig, "predict": predict, "dense_out": outs[0], "decoder_outputs": outs[1:]}


def dccrn_forward(sd, signal, skip_to_use=(0, 1, 2, 3, 4, 5), recon_type="mask", causal=True,
                  hidden=128, stft_params=(512, 100, 400), data_norm=None):
    """DCCRN_.forward(signal, train=False) + standard_DCCRN.forward — model/pvae_module.py:L215-255,
    L174-198 (data_norm off, resynthesis off).  Keys are prefixed ``std_DCCRN.``."""
    pre = "std_DCCRN."
    sub = {k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}
    stft_x = stft(signal, *stft_params)
    if data_norm is not None:                                   # model/pvae_module.py:L217-221
        stft_x = ((stft_x - data_norm[0]) / (data_norm[1] + 1e-6)).clone()
        stft_x[:, 0, :, 1] = 0
        stft_x[:, -1, :, 1] = 0
    x, skiper = encoder_stack(stft_x.unsqueeze(1), sub, 6, causal)
    B, C, Fq, T, D = x.shape
    lstm_in = x.reshape(B, -1, T, D).permute(2, 0, 1, 3)
    lat = complex_lstm(lstm_in, sub, "lstms.0.", hidden, 2).permute(1, 0, 2, 3)
    dense_out = complex_dense(lat.reshape(B * T, -1, D), sub, "dense.")
    p = dense_out.reshape(B, T, C, Fq, D).permute(0, 2, 3, 1, 4)
    for i in range(6):
        if i in skip_to_use:
            p = torch.cat([p, skiper[len(skiper) - i - 1]], dim=1)
        p = decoder_block(p, sub, "decoders.%d." % i, causal)
    if recon_type == "mask":
        predict = mask_head(p, stft_x, 1)
    else:
        predict = torch.complex(p[..., 0], p[..., 1]).squeeze(1)
    if data_norm is not None:                                   # L236-239, L248-249
        pr = data_norm[1] * torch.view_as_real(predict) + data_norm[0]
        predict = torch.complex(pr[..., 0], pr[..., 1])
    return {"clean": istft(predict, *stft_params), "predict": predict, "latent": lat,
            "stft_x": stft_x, "skiper": skiper}


# ------------------------------------------------------------------------------------------------
# parity metrics
# ------------------------------------------------------------------------------------------------
def rel_l2(a, b):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    if a.is_complex():
        a = torch.view_as_real(a)
    if b.is_complex():
        b = torch.view_as_real(b)
    a, b = a.double(), b.double()
    return float(torch.linalg.norm((a - b).flatten()) / (torch.linalg.norm(b.flatten()) + 1e-30))


def si_sdr_db(est, ref):
    """utils/eval_metrics.py:L49-64 formula (scale-invariant SDR in dB), per utterance."""
    est = torch.as_tensor(est).double()
    ref = torch.as_tensor(ref).double()
    alpha = (est * ref).sum(-1, keepdim=True) / ((ref * ref).sum(-1, keepdim=True) + 1e-30)
    tgt = alpha * ref
    noise = est - tgt
    return 10 * torch.log10((tgt ** 2).sum(-1) / ((noise ** 2).sum(-1) + 1e-30))

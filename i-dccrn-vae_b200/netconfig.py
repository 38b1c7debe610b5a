"""Network hyper-parameter dicts with the reference's ``get_net_params()`` dict API.

Mirrors the *values* of /root/reference/model/causal_netconfig.py:L5-103 (causal) and
model/net_config.py:L5-103 (non-causal; the two differ only in the encoder time padding,
L33-40).  Only the channel counts, kernel/stride/padding tuples, ``lstm_dim``, ``dense`` and
``lstm_layer_num`` are consumed by the hot path; of the ``*_chw`` triples only C is used
(model/pvae_module.py:L57 ``ComplexBatchNormal(chw[0], ...)``).
"""

_BASE = 32


def _freq_sizes(n_layers, f0=257):
    # Fout = (Fin + 2*2 - 5)//2 + 1 for kernel 5 / stride 2 / pad 2
    sizes, f = [], f0
    for _ in range(n_layers):
        f = (f + 4 - 5) // 2 + 1
        sizes.append(f)
    return sizes


def get_net_params(causal=True):
    enc_ch = [1, _BASE, _BASE * 2, _BASE * 4, _BASE * 4, _BASE * 8, _BASE * 8]
    dec_ch = [_BASE * 8, _BASE * 8, _BASE * 4, _BASE * 4, _BASE * 2, _BASE, 1]
    n = len(enc_ch) - 1
    fs = _freq_sizes(n)                         # 129, 65, 33, 17, 9, 5
    t_pad = 1 if causal else 0
    p = {
        "encoder_channels": enc_ch,
        "encoder_kernel_sizes": [(5, 2)] * n,
        "encoder_strides": [(2, 1)] * n,
        "encoder_paddings": [(2, t_pad)] * n,
        "lstm_dim": [enc_ch[-1] * fs[-1], 128],   # 1280 -> 128 (128 only used by DCCRN_)
        "dense": [128, enc_ch[-1] * fs[-1]],
        "lstm_layer_num": 2,
        "decoder_channels": dec_ch,
        "decoder_kernel_sizes": [(5, 2)] * n,
        "decoder_strides": [(2, 1)] * n,
        "decoder_paddings": [(2, 0)] * n,
    }
    # (C, F, T) triples: T is a nominal 10 s figure in the reference and is never read.
    p["encoder_chw"] = [(enc_ch[i + 1], fs[i], 1600 - i) for i in range(n)]
    dec_f = fs[::-1][1:] + [257]
    p["decoder_chw"] = [(dec_ch[i + 1], dec_f[i], 1596 + i) for i in range(n)]
    return p


def get_causal_net_params():
    return get_net_params(causal=True)

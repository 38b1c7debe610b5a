"""idccrn_b200 — B200-native (sm_100a) implementation of the I-DCCRN-VAE enhancement forward path behind
the reference's torch.nn.Module API.  See DESIGN.md / INTEGRATION.md."""
from . import build, compat, config, lib, metrics, netconfig, ops, pack, pipeline, ragged, shard, synth, wavio, modules  # noqa: F401
from .modules import *  # noqa: F401,F403
from .modules import (STFT, ISTFT, ConvSTFT, ConviSTFT, ComplexConv2d, causal_complex_conv2d,  # noqa: F401
                      ComplexConvTranspose2d, causal_ComplexConvTranspose2d, ComplexBatchNormal,
                      ComplexBatchNorm, ComplexLSTM, NavieComplexLSTM, ComplexDense, Encoder, Decoder,
                      SkipList, nsvae_pvae_dccrn_encoder_twophase, pvae_dccrn_encoder_skip_prepare,
                      pvae_dccrn_decoder_skip_prepare, nsvae_pvae_dccrn_decoder_twophase, standard_DCCRN, DCCRN_)
from .netconfig import get_net_params  # noqa: F401
# imported eagerly (after modules) so that the ``idccrn_b200`` alias covers them: one module object per file
from . import losses, streaming, train  # noqa: F401,E402

__version__ = "0.1.0"

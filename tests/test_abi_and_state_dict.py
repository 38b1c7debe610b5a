"""CPU tier: the shared library loads and exports every symbol include/idv.h declares; the module surface
keeps the reference's state_dict keys and shapes."""
import ctypes
import json
import os
import re

import pytest
import torch

import common as C
import idccrn_b200 as M
from idccrn_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "idv.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(idv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 15
    cdll = ctypes.CDLL(lib.lib_path()) if os.path.exists(lib.lib_path()) else lib.load()
    missing = [s for s in syms if not hasattr(cdll, s)]
    assert not missing, missing
    assert set(syms) == set(lib.EXPORTS), set(syms) ^ set(lib.EXPORTS)
    cdll.idv_abi_version.restype = ctypes.c_int
    from idccrn_b200 import lib as _lib
    assert cdll.idv_abi_version() == _lib.ABI_VERSION == 8


def test_product_has_no_cpu_fallback():
    """CPU tensors must be rejected loudly (no eager / oracle fallback)."""
    enc, dec = C.build_vae(1, 1, "skip_prepare", "real_imag", 0, "cpu")
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(torch.zeros(1, 800), train=False)
    for mod in ("ops", "modules", "pack", "lib"):
        src = open(os.path.join(ROOT, "i-dccrn-vae_b200", mod + ".py")).read()
        assert "oracle" not in src.replace("no eager / oracle", "")


def test_state_dict_layout_matches_reference():
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    net = M.get_net_params()
    built = {
        "nsvae_pvae_dccrn_encoder_twophase_l1": M.nsvae_pvae_dccrn_encoder_twophase(net, True, "cpu", 128, 512, 100, 400, 1, 1),
        "nsvae_pvae_dccrn_encoder_twophase_l2": M.nsvae_pvae_dccrn_encoder_twophase(net, True, "cpu", 128, 512, 100, 400, 1, 2),
        "pvae_dccrn_encoder_skip_prepare": M.pvae_dccrn_encoder_skip_prepare(net, True, "cpu", 128, 512, 100, 400, 1),
        "pvae_dccrn_decoder_skip_prepare": M.pvae_dccrn_decoder_skip_prepare(net, True, "cpu", 1, 128, 512, 100, 400, "real_imag", C.SKIPS),
        "nsvae_pvae_dccrn_decoder_twophase": M.nsvae_pvae_dccrn_decoder_twophase(net, True, "cpu", 1, 128, 512, 100, 400, "mask", True, C.SKIPS, False),
        "DCCRN_": M.DCCRN_(512, 100, net, True, "cpu", 400, C.SKIPS, "mask", False, None, None),
    }
    for name, mod in built.items():
        got = [[k, list(v.shape)] for k, v in mod.state_dict().items()]
        assert got == want[name], name


def test_netconfig_matches_reference_values():
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))["net_params_causal"]
    got = json.loads(json.dumps(M.get_net_params()))
    assert got == want

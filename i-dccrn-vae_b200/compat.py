"""Drop-in shim for the reference's drivers.

The reference's scripts build their models with ``from model.pvae_module import *`` /
``import model.causal_netconfig as ...`` (e.g. i_dccrn_vae/nsvae_dccrn/test_nsvae_se.py:L17-19, train_nsvae.py:L25-28).
``install()`` registers modules under those names whose classes are the B200-native ones of this package, so an
unmodified script run as

    python -c "import idccrn_b200.compat as c; c.install(); import runpy; runpy.run_path('test_nsvae_se.py', run_name='__main__')"

constructs, loads (``load_state_dict`` of the reference's ``.pt`` checkpoints: same keys and shapes) and calls our
modules.  The loss code (model/nsvae_loss.py, model/pretrain_pvaes_loss.py) is plain torch on the modules' outputs and
is left alone.  Nothing here computes anything."""
import sys
import types

_NAMES = ("model", "model.pvae_module", "model.complex_progress", "model.causal_netconfig", "model.net_config")


def install(force=False):
    """Register ``model.pvae_module`` / ``model.complex_progress`` / ``model.causal_netconfig`` / ``model.net_config``
    backed by this package.  Refuses to shadow an already imported reference ``model`` package unless force=True."""
    from . import modules, netconfig
    if not force:
        for n in _NAMES:
            if n in sys.modules and not getattr(sys.modules[n], "__idccrn_b200_shim__", False):
                raise RuntimeError("%s is already imported (the reference's own package?); call install() first or pass "
                                   "force=True" % n)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__idccrn_b200_shim__ = True
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    public = {k: v for k, v in vars(modules).items() if not k.startswith("_")}
    pkg = mod("model")
    pkg.__path__ = []                                       # a package: ``import model.x`` resolves through sys.modules
    pkg.pvae_module = mod("model.pvae_module", **public)
    prim = ("causal_complex_conv2d", "ComplexConv2d", "ComplexLSTM", "ComplexDense", "ComplexBatchNormal",
            "causal_ComplexConvTranspose2d", "ComplexConvTranspose2d")
    pkg.complex_progress = mod("model.complex_progress", **{k: public[k] for k in prim})
    pkg.causal_netconfig = mod("model.causal_netconfig", get_net_params=lambda: netconfig.get_net_params(True))
    pkg.net_config = mod("model.net_config", get_net_params=lambda: netconfig.get_net_params(False))
    return pkg


def uninstall():
    for n in _NAMES:
        if getattr(sys.modules.get(n), "__idccrn_b200_shim__", False):
            del sys.modules[n]

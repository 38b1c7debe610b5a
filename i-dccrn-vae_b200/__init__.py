from . import netconfig, synth  # noqa: F401

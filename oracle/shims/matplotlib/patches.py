"""ORACLE / TEST INFRASTRUCTURE ONLY -- inert matplotlib submodule stub."""
from . import _Inert


def __getattr__(name):
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    return _Inert()

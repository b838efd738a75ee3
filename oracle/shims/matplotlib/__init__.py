"""ORACLE / TEST INFRASTRUCTURE ONLY.

Inert stand-in for matplotlib (absent from this image, no network).  The
reference imports it at module top (scp.py:3,7; position_generator.py:12-14;
plot_runtime_boxplot.py) but never touches it on the compute path.  Every
attribute resolves to a callable that returns another inert object, so the
plotting methods run to completion without drawing anything.
"""


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Inert()

    def __iter__(self):
        return iter((_Inert(), _Inert()))

    def __getitem__(self, i):
        return _Inert()

    def __len__(self):
        return 0


def __getattr__(name):
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    return _Inert()

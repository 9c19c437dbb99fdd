"""TEST INFRASTRUCTURE — not product code.

Makes the *unmodified* reference (``/root/reference/latentfusion``) importable in this
authoring container so that ``oracle/make_golden.py`` can run it and dump golden vectors.
The GPU box has no ``/root/reference``; nothing here may be used by ``-m gpu`` tests,
``smoke()`` or ``bench.py`` (they consume the committed ``tests/golden/*.npz`` instead).

What is stubbed (packages the reference imports at module-import time but never touches on
the reconstruct->render path; list from SURVEY.md §8c / Appendix B):

* ``structlog``  (latentfusion/__init__.py:7,56-71)
* ``IPython``    (latentfusion/__init__.py:14-25)
* ``imageio``    (latentfusion/observation.py:6)
* ``skimage.morphology`` (latentfusion/imutils.py:4, pose/initialization.py:2)
* ``plyfile``    (latentfusion/pointcloud.py:4)

and one shim: ``ReduceLROnPlateau(verbose=...)`` (pose/estimation.py:589-594) — the kwarg was
removed from torch.
"""
import sys
import types

import os as _os

# the reference as it lies in the authoring container, else the verbatim copy staged for the GPU box (oracle/stage_ref.py)
_STAGED = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), '_ref')
REFERENCE_ROOT = '/root/reference' if _os.path.isdir('/root/reference/latentfusion') else _STAGED


class _Noop:
    """Attribute sink: any non-dunder attribute is a callable returning another sink."""

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _Noop()

    def __call__(self, *args, **kwargs):
        return _Noop()

    def bind(self, **kwargs):
        return self


def _stub_module(name, **attrs):
    mod = types.ModuleType(name)

    def _getattr(attr):
        if attr.startswith('__') and attr.endswith('__'):
            raise AttributeError(attr)
        return _Noop()

    mod.__getattr__ = _getattr
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def install():
    """Install the stubs + shim and put the reference on sys.path. Idempotent."""
    import importlib.util
    if getattr(install, '_done', False):
        return
    install._done = True
    if 'structlog' not in sys.modules and importlib.util.find_spec('structlog') is None:
        sl = _stub_module('structlog', get_logger=lambda *a, **k: _Noop(),
                          configure=lambda **k: None)
        for sub in ('stdlib', 'processors', 'dev'):
            setattr(sl, sub, _stub_module(f'structlog.{sub}'))
    if 'IPython' not in sys.modules and importlib.util.find_spec('IPython') is None:
        _stub_module('IPython', get_ipython=lambda: None)
    if 'toml' not in sys.modules and importlib.util.find_spec('toml') is None:
        # latentfusion/pose/estimation.py:10 imports toml only to read config FILES; the harness passes dicts
        import tomllib

        def _toml_load(path):
            with open(path, 'rb') as fh:
                return tomllib.load(fh)
        _stub_module('toml', load=_toml_load)
    for name in ('imageio', 'plyfile'):
        if name not in sys.modules and importlib.util.find_spec(name) is None:
            _stub_module(name)
    if 'skimage' not in sys.modules and importlib.util.find_spec('skimage') is None:
        sk = _stub_module('skimage')
        sk.morphology = _stub_module('skimage.morphology')

    import torch
    base = torch.optim.lr_scheduler.ReduceLROnPlateau
    if not getattr(base, '_lf_shimmed', False):
        class ReduceLROnPlateau(base):
            _lf_shimmed = True

            def __init__(self, *args, verbose=None, **kwargs):
                super().__init__(*args, **kwargs)

        torch.optim.lr_scheduler.ReduceLROnPlateau = ReduceLROnPlateau

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def available():
    import os
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'latentfusion'))

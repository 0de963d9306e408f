"""Load the reference's OWN files verbatim, by path, under the shims.  TEST INFRASTRUCTURE.

Container only: /root/reference is not present on the GPU box, so nothing that runs there
(-m gpu tests, smoke(), bench.py) may call this.  Used by oracle/make_golden.py and by the
CPU tests that pin oracle/restated.py to the reference.

Nothing is copied: the three files are executed from where they lie
(/root/reference/x_transformers_rl/{distributed,evolution,x_transformers_rl}.py).
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('XTRL_REFERENCE_ROOT', '/root/reference')
_SHIMS = ('einx', 'x_transformers', 'assoc_scan', 'hl_gauss_pytorch', 'ema_pytorch',
          'adam_atan2_pytorch', 'accelerate')
_PKG = 'x_transformers_rl'


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, _PKG, 'x_transformers_rl.py'))


def _install_shims():
    for name in _SHIMS:
        if name in sys.modules:
            continue
        sys.modules[name] = importlib.import_module(f'oracle.shims.{name}')


def load():
    """Returns the reference's `x_transformers_rl.x_transformers_rl` module (with `.evolution`
    and `.distributed` reachable as `sys.modules['x_transformers_rl.evolution']` etc.)."""
    if not available():
        raise FileNotFoundError(f'reference tree not found under {REFERENCE_ROOT}')
    full = f'{_PKG}.x_transformers_rl'
    if full in sys.modules and getattr(sys.modules[full], '__xtrl_ref__', False):
        return sys.modules[full]
    _install_shims()
    pkg_dir = os.path.join(REFERENCE_ROOT, _PKG)
    pkg = types.ModuleType(_PKG)
    pkg.__path__ = [pkg_dir]          # submodules resolve to the reference's own files
    pkg.__xtrl_ref__ = True
    sys.modules[_PKG] = pkg
    mods = {}
    for sub in ('distributed', 'evolution', 'x_transformers_rl'):
        name = f'{_PKG}.{sub}'
        spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, f'{sub}.py'))
        mod = importlib.util.module_from_spec(spec)
        mod.__xtrl_ref__ = True
        sys.modules[name] = mod
        sys.dont_write_bytecode, prev = True, sys.dont_write_bytecode   # /root/reference is read-only
        try:
            spec.loader.exec_module(mod)
        finally:
            sys.dont_write_bytecode = prev
        setattr(pkg, sub, mod)
        mods[sub] = mod
    pkg.Learner, pkg.Agent = mods['x_transformers_rl'].Learner, mods['x_transformers_rl'].Agent
    return mods['x_transformers_rl']


def evolution():
    load()
    return sys.modules[f'{_PKG}.evolution']

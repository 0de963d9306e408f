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


_CACHE = {}


def load():
    """Returns the reference's `x_transformers_rl.x_transformers_rl` module.  The reference files import
    each other as `x_transformers_rl.*`, so that name is bound to a synthetic parent package while
    they execute and then RESTORED: the repo's own drop-in package of the same import name
    (x_transformers_rl/ at the repo root) stays importable in the same process."""
    if not available():
        raise FileNotFoundError(f'reference tree not found under {REFERENCE_ROOT}')
    if 'x_transformers_rl' in _CACHE:
        return _CACHE['x_transformers_rl']
    _install_shims()
    saved = {k: v for k, v in sys.modules.items() if k == _PKG or k.startswith(_PKG + '.')}
    for k in saved:
        del sys.modules[k]
    pkg_dir = os.path.join(REFERENCE_ROOT, _PKG)
    pkg = types.ModuleType(_PKG)
    pkg.__path__ = [pkg_dir]          # submodules resolve to the reference's own files
    pkg.__xtrl_ref__ = True
    sys.modules[_PKG] = pkg
    try:
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
            _CACHE[sub] = mod
    finally:
        for k in [k for k in sys.modules if k == _PKG or k.startswith(_PKG + '.')]:
            del sys.modules[k]
        sys.modules.update(saved)
    pkg.Learner, pkg.Agent = _CACHE['x_transformers_rl'].Learner, _CACHE['x_transformers_rl'].Agent
    _CACHE['package'] = pkg
    return _CACHE['x_transformers_rl']


def evolution():
    load()
    return _CACHE['evolution']


def distributed():
    load()
    return _CACHE['distributed']

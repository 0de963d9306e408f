"""Drop-in import name of the reference (`from x_transformers_rl import Learner, Agent`,
reference x_transformers_rl/__init__.py:1-4), backed by the B200 package `x-transformers-rl_b200`."""
import importlib as _importlib
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
b200 = _importlib.import_module('x-transformers-rl_b200')
ops = b200.ops
XtrlError = b200.XtrlError


def __getattr__(name):
    return getattr(b200, name)

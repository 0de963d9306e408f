"""xtrl -- B200-native rollout-and-learn hot path with the API of tinycrops/x-transformers-rl.

    from x_transformers_rl import Learner, Agent        # drop-in alias package at the repo root

The directory name carries a hyphen (the build contract names it `x-transformers-rl_b200`), so it is
imported through importlib; `x_transformers_rl/` at the repo root re-exports it under the
reference's import name (x_transformers_rl/__init__.py:1-4 of the reference).

All arithmetic on the hot path runs in libxtrl.so (hand-written sm_100a CUDA, include/xtrl.h);
there is no CPU fallback: a missing library or a non-CUDA tensor raises.
"""
from ._lib import XtrlError, load as load_library, lib_path  # noqa: F401
from . import ops  # noqa: F401

__all__ = ['XtrlError', 'load_library', 'lib_path', 'ops', 'Learner', 'Agent']


def __getattr__(name):  # Learner / Agent import torch.nn machinery; keep `import` of the ops light
    if name in ('Learner', 'Agent'):
        from . import learner as _learner
        return getattr(_learner, name)
    raise AttributeError(name)

"""Build libxtrl.so (the C-ABI CUDA library, include/xtrl.h) in-tree with nvcc for sm_100a.

    python "x-transformers-rl_b200/build.py"        # or __graft_entry__.build()

The .so is written next to this file (lib/libxtrl.so): it is git-ignored but travels to the GPU
box with the gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_DIR = os.path.join(HERE, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libxtrl.so')
STAMP = LIB_PATH + '.srchash'
SOURCES = ('api.cu', 'gae.cu', 'ppo_loss.cu', 'evolve.cu', 'optim.cu', 'wm_step.cu', 'learn_ff.cu', 'learn_attn.cu', 'learn_attn_core.cu', 'learn_ln.cu', 'learn_mlp.cu', 'learn_prep.cu')
NVCC_FLAGS = ('-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '-cudart', 'shared')


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: libxtrl.so cannot be built (there is no CPU fallback)')


def source_hash() -> str:
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cu', '.cuh'))]
    files.append(os.path.join(os.path.dirname(HERE), 'include', 'xtrl.h'))
    for f in files:                                   # names, not absolute paths: the snapshot on a GPU box lives elsewhere
        h.update(os.path.basename(f).encode()); h.update(open(f, 'rb').read())
    h.update((' '.join(NVCC_FLAGS) + os.environ.get('XTRL_NVCC_EXTRA', '')).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    return os.path.exists(LIB_PATH) and os.path.exists(STAMP) and open(STAMP).read().strip() == source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and is_fresh():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    import fcntl
    with open(os.path.join(LIB_DIR, '.build.lock'), 'w') as lock:      # one builder at a time (torchrun starts N ranks at once)
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and is_fresh():                                   # another rank built it while we waited
            return LIB_PATH
        tmp = LIB_PATH + f'.tmp{os.getpid()}'
        cmd = [_nvcc(), *NVCC_FLAGS, *os.environ.get('XTRL_NVCC_EXTRA', '').split(), '-Xptxas', '-v', '-o', tmp, *[os.path.join(CSRC, s) for s in SOURCES]]
        res = subprocess.run(cmd, capture_output=True, text=True)
        with open(os.path.join(LIB_DIR, 'build.log'), 'w') as f:
            f.write(' '.join(cmd) + '\n' + res.stdout + res.stderr)
        if res.returncode != 0:
            if os.path.exists(tmp):
                os.remove(tmp)
            raise RuntimeError('nvcc failed:\n' + res.stderr[-4000:])
        os.replace(tmp, LIB_PATH)                                      # atomic: readers never see a half-written library
        with open(STAMP, 'w') as f:
            f.write(source_hash())
    if verbose:
        print(res.stderr, file=sys.stderr)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

"""oracle/ -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).

CPU restatement of the rollout-and-learn hot path of tinycrops/x-transformers-rl
(reference tree: /root/reference, read-only; citations are relative to it).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import anything from here -- and only as the checker or as
the timed CPU baseline, never as the thing shipped.  The product package
(``x-transformers-rl_b200/``) must not import ``oracle`` and fails loudly when its CUDA
library is missing.

Pieces
  shims/          clean-room stand-ins for the seven third-party packages the reference
                  imports but this image lacks (einx, x_transformers, assoc_scan,
                  hl_gauss_pytorch, ema_pytorch, adam_atan2_pytorch, accelerate).  Only the
                  call surface the reference uses.
  ref_loader.py   loads /root/reference/x_transformers_rl/{distributed,evolution,
                  x_transformers_rl}.py VERBATIM by path under those shims.  Container only:
                  /root/reference does not exist on the GPU box.
  restated.py     own-words torch-CPU port of the same arithmetic; this is what travels to
                  the GPU box and what the CUDA kernels are checked against.
  make_golden.py  runs the reference-owned code (ref_loader) on seeded inputs and writes
                  tests/golden/*.npz; tests/ pins restated.py to those vectors.

PARITY STATUS
  reference-owned arithmetic (calc_gae, RSNorm, Discrete, normalize, the four loss
  functions, WorldModelActorCritic.forward head wiring, LatentGenePool.evolve_):
      PINNED -- restated.py equals the reference's own code run here, see tests/golden.
  third-party arithmetic reached through the shims (x_transformers Decoder internals,
  HLGaussLoss, AssocScan, AdoptAtan2, EMA):
      PARITY UNPINNED -- the packages are absent from /root/reference and from this
      image (pyproject.toml:26-37 gives lower bounds only, no lockfile); the shims restate
      their published behaviour from memory and *define* the oracle there.
"""

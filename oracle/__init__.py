"""CPU oracle for the VQ-VAE inference hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker.  The
product (``vqvae_b200/``, ``models/``) never imports it.

Two restatements live here:

* ``oracle.cref``  -- plain C (``oracle/csrc/oracle.c``, built by ``oracle.build()``
  with gcc into ``oracle/_build/``) driven through ctypes.  Defines the canonical
  fp32 arithmetic the CUDA VQ kernel must match bit for bit, and double-accumulated
  convolutions for tolerance checks.
* ``oracle.torch_port`` -- the same path written with the torch CPU ops the reference
  itself calls (``F.conv2d``, ``F.conv_transpose2d``, ``matmul``, ``argmin``), used
  as the CPU baseline ("port": identical oneDNN/MKL back ends to the reference) and
  as a second checker.

Parity pin: the reference has no tests or golden vectors, so both
restatements are pinned against outputs of the UNMODIFIED reference imported from
/root/reference in the authoring container; ``oracle/make_golden.py`` is the
generating script and ``tests/golden/*.npz`` the committed vectors.
"""
from .build import build, build_ref, lib_path, ref_path  # noqa: F401

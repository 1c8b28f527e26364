"""oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU (plain PyTorch fp32 / numpy) restatement of the reference's reverse-diffusion
denoising step (OvO1111/JointImageGeneration, "GuideGen"), used exclusively as the
*checker* for the sm_100a CUDA path:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import anything from here;
  * nothing under ``jointimagegeneration_b200/`` imports it -- the product path fails
    loudly when the CUDA library is missing, it never falls back to this code.

Parity pinning: the reference ships no tests / golden vectors, so the
restatement is pinned against the *reference itself*, imported read-only from
``/root/reference`` in the build container through ``oracle.refshim`` (four import
shims, no code copied).  ``oracle/make_golden.py`` runs the reference modules on seeded
inputs and commits the resulting vectors under ``tests/golden/``; ``tests/test_oracle_*``
check the restatement against those vectors everywhere, and against the live reference
whenever ``/root/reference`` is present.

Every function cites the reference file:line it restates (paths relative to
``/root/reference``).
"""

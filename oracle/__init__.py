"""CPU oracle for the DreamWaltz-G SDS hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the algorithms of the reference's per-step SDS path
(SURVEY.md section 8a, rows R0-R17).  It exists so that the CUDA product path under
``dreamwaltz-g_b200/`` can be checked against an independent implementation.

Rules (enforced by tests/test_no_oracle_in_product.py):
  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
    ``--impl reference`` legs may import anything from here;
  * nothing under ``dreamwaltz-g_b200/`` imports it; the product path raises if its CUDA
    library is missing rather than falling back to this code.

Parity pinning status (see DESIGN.md section "Oracle"):
  * in-tree reference code (RigidTransform, GeneralLinearBlendSkinning, lbs_transform,
    mesh-bound Gaussians, eval_sh, DeformNetwork, MLP): PINNED -- the reference's own
    Python was executed in the build container against these restatements and the
    resulting vectors are committed under tests/golden/ (tests/golden/make_golden.py).
  * un-vendored third-party maths (smplx.lbs, pytorch3d.transforms,
    diff_gaussian_rasterization, diffusers, and the CUDA-only gridencoder):
    PARITY UNPINNED -- restated from the published algorithms,
    anchored on the reference's call sites and on hand-computed known-answer tests.
"""

"""CPU oracle for the SuperPoint/MagicPoint inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as the thing shipped.

What it is: a CPU restatement (torch-CPU fp32 / numpy / plain C) of the
reference algorithm for the path named in SURVEY.md section 8, each function citing the
reference file:line it follows (paths relative to
``/root/reference/superpoint/superpoint/``).

Modules: ``spn_oracle`` (model forward, box_nms, sampler, homography adaptation, pre-processing, detector labels),
``eval_oracle`` (the rows SURVEY.md section 8f adds: repeatability, keep_shared_points, cross-checked matching, the
NeRF re-projection step), ``kornia_shim`` (kornia 0.7.0 restated), ``nms_ref.c`` (greedy NMS in plain C).

Pinning status
--------------
* The reference ships no tests, golden vectors or fixtures for this path
  (SURVEY.md section 4), so the oracle is pinned against *outputs of the reference
  itself*, produced in the build container by ``tests/golden/make_golden.py``
  (which imports the unmodified reference from ``/root/reference`` with
  ``oracle/kornia_shim.py`` injected for the missing ``kornia`` package) and
  committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks
  every oracle function against those vectors.
* kornia 0.7.0 (``warp_perspective``, ``morphology.erosion``) is a pinned
  third-party dependency of the reference that is absent from ``/root/reference``
  and from this image.  ``oracle/kornia_shim.py`` restates its published
  algorithm; that part of the parity is therefore "restated, cross-checked
  against cv2.warpPerspective", not pinned against kornia's own binaries.
"""

"""CPU oracle for the cggp hot path.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This package is a NumPy restatement of the reference's algorithm for the conjugate-gradient
hot path (reference files under /root/reference/cggp, cited per function).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
it, and only as the checker or the timed CPU baseline.  Nothing under ``cggp_b200/`` imports it.

Pinning status
--------------
* ``oracle.cg`` (the CG loop, preconditioners, layout adapter) and ``oracle.models`` (CGGP /
  ClusterGP ``prior_kl`` / ``predict_f`` / ``elbo``) are pinned against the reference's OWN, UNMODIFIED
  source files ``cggp/conjugate_gradient.py`` and ``cggp/models.py`` executed in the build container over a
  NumPy shim of the TensorFlow / GPflow symbols they call (``tests/golden/_shim``, generator
  ``tests/golden/make_golden.py``); the resulting vectors are committed under ``tests/golden/*.npz``.
* ``oracle.gpflow_restated`` (SE / Matern kernels, ``square_distance``, ``Kuu``/``Kuf``, Gaussian
  ``variational_expectations``, ``SGPR``) restates GPflow 2.x, a third-party dependency that is NOT vendored in
  the reference and only lower-bounded (``requirements.txt:1-3``: gpflow>=2.5.2, tensorflow>=2.7.0,
  tensorflow_probability>=0.15.0; no lock file).  TensorFlow and GPflow are not installable here (no network,
  not in the wheelhouse) and the reference's tests hold no golden vectors or seeds for kernel values
  (``cggp/cg_test.py:20-26`` draws un-seeded inputs).  **Parity is therefore UNPINNED at the GPflow boundary**:
  kernel values are certified only against closed forms and SciPy (``tests/test_oracle_gpflow.py``).
"""

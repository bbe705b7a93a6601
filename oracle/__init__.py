"""CPU oracle for the bo_pr probabilistic-reparameterization (PR) hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``bo_pr_b200/`` imports this package; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may.

Parity status
-------------
* L1 (the reference's own code: PR input transforms, MC/analytic PR wrappers, score-function gradient,
  ``probabilistic_reparameterization.py`` / ``input.py``): **pinned** — ``oracle/gen_golden.py`` imports the
  reference files unmodified from ``/root/reference`` (through the BoTorch-surface stand-in in
  ``oracle/shim``), runs them on seeded inputs and commits the outputs under ``tests/golden/``;
  ``oracle/pr_oracle.py`` (the restatement) is checked against those vectors bit-for-bit on the discrete
  samples and to 1e-13 on values/gradients.
* L0 (GP posterior, Matérn / categorical kernels, EI / UCB / posterior mean): the arithmetic lives in
  BoTorch / GPyTorch, which are third-party, unpinned (``setup.py:9-20``: ``botorch>=0.6``, ``gpytorch``)
  and absent from ``/root/reference`` and from this image.  ``oracle/gp_oracle.py`` restates their
  published formulas.  The reference holds no golden vectors or tests for this
  boundary, so L0 is **parity unpinned** with respect to BoTorch/GPyTorch outputs: it is anchored on the
  reference's call sites (``experiment_utils.py:267-351, 391-393, 476-480``; ``kernels.py:18-101``), on
  internal known-answer checks (posterior at training inputs, finite differences, Σ probs = 1) and on an
  INDEPENDENT implementation of the same published formulas that is installed here: scikit-learn's
  ``GaussianProcessRegressor`` with ``Matern(nu=2.5)`` kernels (kernel matrix 1e-10, posterior mean 1e-8,
  variance 1e-6) and ``scipy.stats.norm`` for EI (``tests/test_oracle_l0_independent.py``).  Library-specific
  constants (EI variance floor, ``min_variance``, ``min_fixed_noise``) stay exposed as parameters.
"""

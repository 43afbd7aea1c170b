"""CPU oracle for the batched closed-loop rollout path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy / pure-Python fp64, plus a plain-C fp64
twin under ``oracle/c``) of the reference's algorithm for the one hot path this
repo accelerates:  tracking law -> bicycle-model step -> lap-metric accumulation
(SURVEY.md section 8).  It exists to CHECK the CUDA path, never to serve it.

Who may import / call / link / execute anything under ``oracle/``:
  * ``tests/``                            (the checker in parity tests)
  * ``__graft_entry__.smoke()``           (one tiny parity check on cuda:0)
  * ``bench.py``'s ``cpu_baseline`` leg and ``bench.py --impl reference``
Nothing in ``bgr_pathplanning_control_b200/`` imports it; the product path fails
loudly when the CUDA extension is missing and has no CPU fallback.

Parity status
-------------
The reference (/root/reference, 100 % Python) ships NO tests and NO golden
vectors.  The oracle is therefore pinned against OUTPUTS
OF THE REFERENCE ITSELF:  ``oracle/gen_golden.py`` imports the reference's
unmodified ``core.controllers``, ``core.data.car_state``, ``vehicle_config`` and
``preformance_analysis.metrics_calculator`` in the build container (stubbing only
the absent third-party imports, see ``oracle/ref_shim.py``) and freezes their
outputs on seeded inputs into ``tests/golden/*.json``;  ``tests/test_oracle_golden.py``
checks every oracle function against those vectors bit-for-bit (fp64).
``oracle/gen_golden_shadowed.py`` does the same for the dead-code Stanley variant (executed from the reference's
own source lines via ``ast``) and ``oracle/gen_golden_replan.py`` for the re-planning regime (the reference's
``Planner`` and ``Controller`` NODES run unmodified around a stubbed planner library and simulator call).
``tests/test_oracle_vs_reference_live.py`` additionally calls the reference live, on freshly drawn inputs, wherever
the tree is mounted.

Pieces whose arithmetic lives in third-party packages that are absent from
/root/reference and from this image are "parity unpinned" and say so where they
are defined:
  * simple_pid.PID        (PyPI ``simple-pid``, version un-pinned, environment.yaml:14)
  * control.dlqr          (``python-control``, un-pinned, environment.yaml:13)
  * cvxpy + OSQP optimum  (un-pinned, environment.yaml:15) -- only the cost /
                          dynamics definition (core/controllers.py:511-548) is restated.
ORACLE-DEFINED extensions (no reference code exists): the kinematic
bicycle step, per-step error logging for non-Stanley laws, cone hits, the
lap-completion rule for closed tracks, observation noise, and the window planner that stands in for
fsd-path-planning in the re-planning regime.  Each is documented in ``oracle/loop.py``.
"""

"""CPU oracle: a NumPy FP64 restatement of the reference's DDP/iLQR hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`trajectory_optimization_matrix_lie_groups_b200/`) may import this package; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs do.

The reference (`/root/reference`, chenghuailin/trajectory_optimization_matrix_lie_groups) is pure
Python but cannot be imported in this image (jax, manifpy, casadi are absent), and its Lie-group
arithmetic lives in the un-vendored, un-pinned third-party library manifpy (pybind11 wrapper of
artivis/manif).  This package therefore restates

  * manif's published closed forms for SO(3)/SE(3) (`oracle/lie.py`),
  * traoptlibrary's dynamics / cost / constraint callbacks (`oracle/models.py`),
  * traoptlibrary's controllers, loop by loop (`oracle/solvers.py`),

each function citing the reference file:line it follows.  Parity is PINNED: `tests/test_oracle_golden.py`
replays the result pickles the reference ships (`visualization/results_benchmark_*_draft/*.pkl`,
re-packed by `tests/golden/make_golden.py` into `tests/golden/*.npz`) and requires identical
iteration counts and <=1e-12 relative error on every cost-history entry.
"""

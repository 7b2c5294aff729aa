"""Drop-in mirror of the reference's `traoptlibrary` package for the DDP/iLQR tracking path.

Same module names, class names, constructor / `fit` signatures, callback signatures and return
tuples as chenghuailin/trajectory_optimization_matrix_lie_groups (traoptlibrary/traopt_*.py), so

    from traoptlibrary.traopt_controller import iLQR_Tracking_SE3_MS

becomes

    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary.traopt_controller import iLQR_Tracking_SE3_MS

The classes are parameter carriers: every number (dynamics, Jacobians, costs, Riccati sweeps,
rollouts) is produced by the native CUDA library through `BatchSolver`.  Controllers additionally
offer `fit_batch`, the batched replacement of the reference's joblib sweeps.
"""

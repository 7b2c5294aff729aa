"""Write the reference trajectories the BASELINE.json workloads track into the package's own data directory.

    python scripts/make_workload_data.py

Output: trajectory_optimization_matrix_lie_groups_b200/data/<name>.npy in the reference's own on-disk format
(consecutive np.save of q_ref, xi_ref, dt; see io.py), under the file names the reference's scripts load
(main_SE3ddp_tracking_exact_ms.py:105-110, benchmark_SO3_tracking.py:49-55, benchmark_drone_racing_tracking.py:50-54).
Source: /root/reference/visualization/optimized_trajectories/ when present (build container), else the problem
definitions embedded in the committed result fixtures (tests/golden/*.npz) — the arrays are identical, which this
script asserts when both exist.  `workloads.py` (and with it bench.py's timed workload) reads only the package data.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "trajectory_optimization_matrix_lie_groups_b200")
REF = "/root/reference/visualization/optimized_trajectories"
GOLD = os.path.join(ROOT, "tests", "golden")

FILES = {   # data file -> (golden fixture that embeds the same arrays, dt written into the file)
    "path_dense_random_columns_4obj": ("se3_n955_r1e-5", 0.004),
    "path_3dpendulum_8shape_tryout": ("so3_n249", 0.04),
}


def main():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_io", os.path.join(PKG, "io.py"))
    io = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(io)
    os.makedirs(os.path.join(PKG, "data"), exist_ok=True)
    for name, (gold, dt) in FILES.items():
        with np.load(os.path.join(GOLD, gold + ".npz")) as z:
            q, xi = z["prob_q_ref"], z["prob_xi_ref"]
        src = os.path.join(REF, name + ".npy")
        if os.path.exists(src):
            q2, xi2, dt2 = io.load_reference_trajectory(src)
            assert np.array_equal(q, q2) and np.array_equal(xi, xi2), name
            if dt2 is not None:
                assert dt2 == dt, (name, dt2)
        out = os.path.join(PKG, "data", name + ".npy")
        io.save_reference_trajectory(out, q, xi, dt)
        print(out, q.shape, xi.shape, dt)
    with np.load(os.path.join(GOLD, "drone_n150.npz")) as z, np.load(os.path.join(GOLD, "se3_n955_r1e-5.npz")) as s:
        assert np.array_equal(z["prob_q_ref"], s["prob_q_ref"][:151]) and np.array_equal(z["prob_xi_ref"], s["prob_xi_ref"][:151])


if __name__ == "__main__":
    main()

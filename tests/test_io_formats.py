"""Reference-trajectory ingestion against what the reference's own notebook produced:
rows of its planner CSV -> rows of the .npy it shipped, its twist-integrated and finite-differenced references."""
import csv
import os

import numpy as np
import pytest

from trajectory_optimization_matrix_lie_groups_b200 import io

FIX = np.load(os.path.join(os.path.dirname(__file__), "golden", "io_fixture.npz"))
REF_DIR = "/root/reference/visualization/optimized_trajectories/"


def test_csv_rows_convert_to_the_shipped_reference(tmp_path):
    p = tmp_path / "sample.csv"
    with open(p, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(FIX["csv_head"].tolist())
        for r in FIX["csv_rows"]:
            wr.writerow([repr(float(v)) for v in r])
    q, xi, _ = io.reference_from_csv(str(p), dt=0.004)
    assert np.max(np.abs(q - FIX["q_ref"])) < 1e-14
    assert np.max(np.abs(xi - FIX["xi_ref"])) < 1e-13
    # and the file round trip in the reference's three-np.save layout
    out = tmp_path / "ref.npy"
    io.save_reference_trajectory(str(out), q, xi, 0.004)
    q2, xi2, dt2 = io.load_reference_trajectory(str(out))
    assert np.array_equal(q, q2) and np.array_equal(xi, xi2) and dt2 == 0.004


def test_generated_references_match_the_shipped_files():
    xi, dt = FIX["sine_xi"], float(FIX["sine_dt"])
    q = io.twist_integrated_reference(FIX["sine_q"][0], xi[1:], dt)
    assert np.max(np.abs(q[::20] - FIX["sine_q"])) < 1e-12
    R, w = io.so3_reference_from_rotations(io.eight_shape_rotations(), float(FIX["pend_dt"]))
    assert np.max(np.abs(R[::25] - FIX["pend_q"])) < 1e-14
    assert np.max(np.abs(w[::25] - FIX["pend_w"])) < 1e-12


@pytest.mark.skipif(not os.path.isdir(REF_DIR), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("name", ["path_dense_random_columns_4obj", "path_dense_random_columns"])
def test_whole_csv_against_reference_npy(name):
    q, xi, _ = io.reference_from_csv(REF_DIR + name + ".csv")
    q0, xi0, _ = io.load_reference_trajectory(REF_DIR + name + ".npy")
    assert np.max(np.abs(q - q0)) < 1e-14 and np.max(np.abs(xi - xi0)) < 1e-13

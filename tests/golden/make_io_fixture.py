"""Fixture for the reference-trajectory ingestion (io.reference_from_csv etc.): a few rows of the reference's planner
CSV next to the rows of the .npy its notebook produced from them, and samples of its generated references.
Run in the build container (reads /root/reference):  python tests/golden/make_io_fixture.py"""
import csv

import numpy as np

D = "/root/reference/visualization/optimized_trajectories/"
ROWS = [0, 1, 2, 3, 200, 201, 480, 700, 954, 955]


def load(path):
    with open(path, "rb") as f:
        q, xi = np.load(f), np.load(f)
        try:
            dt = float(np.load(f))
        except Exception:
            dt = np.nan
    return q, xi, dt


with open(D + "path_dense_random_columns_4obj.csv", newline="") as f:
    rows = list(csv.reader(f))
head = rows[0]
data = np.array([[float(v) for v in r] for r in rows[1:]])
q, xi, dt = load(D + "path_dense_random_columns_4obj.npy")
sq, sxi, sdt = load(D + "path_se3_generate_sine_3.npy")
pq, pw, pdt = load(D + "path_3dpendulum_8shape_tryout.npy")
np.savez_compressed("tests/golden/io_fixture.npz", csv_head=np.array(head), csv_rows=data[ROWS], q_ref=q[ROWS], xi_ref=xi[ROWS],
                    dt=dt, sine_q=sq[::20], sine_xi=sxi, sine_dt=sdt, pend_q=pq[::25], pend_w=pw[::25], pend_dt=pdt)

"""Host-side logic that needs no GPU: layout conversions, workload builders, batch sharding and the
one collective of the path (all-gather of per-problem summaries) over gloo with world_size 2."""
import os
import socket

import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from trajectory_optimization_matrix_lie_groups_b200 import distributed as D
from trajectory_optimization_matrix_lie_groups_b200 import layout, workloads


def test_rot_quat_roundtrip_matches_scipy():
    rng = np.random.default_rng(0)
    R = Rotation.from_rotvec(rng.standard_normal((200, 3)) * np.linspace(1e-9, 3.1, 200)[:, None]).as_matrix()
    q = layout.rot_to_quat(R)
    qs = Rotation.from_matrix(R).as_quat()       # [x, y, z, w], what the reference feeds manif (traopt_utilis.py:331-342)
    sgn = np.sign(np.sum(q * qs, axis=1, keepdims=True))
    assert np.max(np.abs(q - sgn * qs)) < 1e-15
    assert np.max(np.abs(layout.quat_to_rot(q) - R)) < 1e-14
    T = np.tile(np.eye(4), (200, 1, 1))
    T[:, :3, :3] = R
    T[:, :3, 3] = rng.standard_normal((200, 3))
    rows = layout.se3_to_rows(T)
    assert rows.shape == (200, 7) and np.max(np.abs(layout.rows_to_se3(rows) - T)) < 1e-14


def test_pose_rows_accepts_the_api_types():
    assert layout.pose_rows(True, np.eye(3)).tolist() == [0, 0, 0, 1]
    assert layout.pose_rows(False, np.eye(4)).tolist() == [0, 0, 0, 1, 0, 0, 0]
    assert np.allclose(layout.pose_rows(True, np.array([0, 0, 0, 2.0])), [0, 0, 0, 1])
    with pytest.raises(ValueError):
        layout.pose_rows(False, np.eye(3))


@pytest.mark.parametrize("cfg,B,kind,method,N", [(1, 4, "se3", "ss", 955), (2, 16, "so3", "ms", 249), (3, 24, "se3", "ms", 955),
                                                (4, 24, "se3", "al_ms", 1400), (5, 24, "drone", "ms", 150)])
def test_workload_builders(cfg, B, kind, method, N):
    wl = workloads.CONFIGS[cfg](B=B)
    assert (wl.kind, wl.method, wl.N, wl.B) == (kind, method, N, B)
    assert wl.x0_rows.shape == (B, 7 if kind == "so3" else 13)
    assert np.max(np.abs(np.linalg.norm(wl.x0_rows[:, :4], axis=1) - 1)) < 1e-15
    again = workloads.CONFIGS[cfg](B=B)
    assert np.array_equal(again.x0_rows, wl.x0_rows)                  # seeded
    if kind != "so3":
        # one parameter at a time: problem b differs from problem 0 only in parameter b mod 12
        d = wl.x0_rows - wl.x0_rows[0]
        for b in range(1, B):
            j = b % 12
            quat, pos, w, v = d[b, :4], d[b, 4:7], d[b, 7:10], d[b, 10:13]
            changed = [np.any(quat != 0), np.any(w != 0), np.any(pos != 0), np.any(v != 0)]
            assert changed == [j < 3, 3 <= j < 6, 6 <= j < 9, j >= 9], (b, changed)
    # a bigger batch extends a smaller one?  No: the draws are per batch; rank shards therefore slice ONE global batch
    big = workloads.CONFIGS[cfg](B=2 * B)
    assert big.x0_rows.shape[0] == 2 * B


def test_helix_reference_is_a_constant_twist_curve():
    q, xi = workloads.helix_reference(50, 0.01)
    step = np.linalg.inv(q[0]) @ q[1]
    for i in range(50):
        assert np.max(np.abs(np.linalg.inv(q[i]) @ q[i + 1] - step)) < 1e-13
    assert np.all(xi == xi[0])


def test_shard_bounds_partition_the_batch():
    for B in (1, 7, 16, 16384, 1 << 20):
        for ws in (1, 2, 3, 8):
            cuts = [D.shard_bounds(B, r, ws) for r in range(ws)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        D.shard_bounds(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, ws, port, B, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        lo, hi = D.shard_bounds(B, rank, ws)
        idx = torch.arange(lo, hi, dtype=torch.float64)
        local = {"J": idx * 1.5, "grad": idx * 1e-13, "defect": idx * 1e-7, "iters": (idx % 50).to(torch.int32),
                 "status": (idx % 3).to(torch.int32)}
        table = D.all_gather_summaries(D.pack_summary(local), B)
        out = D.unpack_summary(table)
        q.put((rank, {k: v.numpy() for k, v in out.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [10, 11])
def test_all_gather_summaries_gloo_world2(B):
    """Even and ragged shards; every rank ends with the whole batch's summaries in problem order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(B, dtype=np.float64)
    for rank in (0, 1):
        o = got[rank]
        assert np.array_equal(o["J"], idx * 1.5) and np.array_equal(o["grad"], idx * 1e-13)
        assert np.array_equal(o["iters"], (idx % 50).astype(np.int32)) and o["iters"].dtype == np.int32
        assert np.array_equal(o["status"], (idx % 3).astype(np.int32))


def test_reference_arm_and_cpu_sample_smoke():
    """bench.py's CPU leg on a 2-core sample of the SO3 config (seconds): the JSON contract keys are present."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TRAJOPT_BENCH_CORES="2")
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-500:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "solves/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 2
    assert line["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_file_formats_roundtrip(tmp_path):
    """Reference trajectories (.npy triplets) and result pickles in the reference's own layouts."""
    import pickle
    from trajectory_optimization_matrix_lie_groups_b200 import io as tio
    q, xi = workloads.helix_reference(20, 0.01)
    p = tmp_path / "ref.npy"
    tio.save_reference_trajectory(p, q, xi, 0.01)
    with open(p, "rb") as f:                      # exactly how the reference scripts read it
        q2 = np.load(f); xi2 = np.load(f); dt2 = np.load(f)
    assert np.array_equal(q2, q) and np.array_equal(xi2, xi) and float(dt2) == 0.01
    q3, xi3, dt3 = tio.load_reference_trajectory(p)
    assert np.array_equal(q3, q) and dt3 == 0.01
    tio.save_reference_trajectory(p, q, xi)       # the variant without dt
    assert tio.load_reference_trajectory(p)[2] is None
    with pytest.raises(ValueError):
        tio.save_reference_trajectory(p, q[:, :3, :2], xi)
        tio.load_reference_trajectory(p)
    prob = dict(J=np.eye(6), dt=0.01, q_ref=q, xi_ref=xi, x0=[q[0], xi[0]], Q=np.eye(12), P=np.eye(12), R=np.eye(6))
    run = dict(xs=[[q[i], xi[i]] for i in range(21)], us=np.zeros((20, 6)), J_hist=[3.0, 2.0], grad_hist=[1.0, 0.5],
               defect_hist=[1.0, 1e-14, 1e-14])
    f = tmp_path / "res.pkl"
    tio.save_results_pickle(f, prob, ms_se3=run, ss_se3={k: v for k, v in run.items() if k != "defect_hist"})
    d = pickle.load(open(f, "rb"))
    assert set(d) == {"prob", "ms_se3", "ss_se3"} and set(d["prob"]) == {"J", "dt", "q_ref", "xi_ref", "x0", "Q", "P", "R"}
    assert len(d["ms_se3"]["xs"]) == 21 and np.array_equal(d["ms_se3"]["xs"][3][0], q[3])
    assert "defect_hist" not in d["ss_se3"] and d["ms_se3"]["J_hist"] == [3.0, 2.0]


def test_utilis_converters_and_lie_helpers():
    """traopt_utilis mirror: hat/vee/ad maps and the manif converters (traopt_utilis.py:13-92, 331-399)."""
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary import traopt_utilis as tu
    rng = np.random.default_rng(1)
    xi = rng.standard_normal(6)
    assert np.array_equal(tu.unskew(tu.skew(xi[:3])), xi[:3])
    assert np.array_equal(tu.se3_vee(tu.se3_hat(xi)), xi)
    ad = tu.adjoint(xi)
    assert np.array_equal(ad[:3, :3], tu.skew(xi[:3])) and np.array_equal(ad[3:, :3], tu.skew(xi[3:])) and np.all(ad[:3, 3:] == 0)
    assert np.array_equal(tu.coadjoint(xi), ad.T)
    T = np.eye(4)
    T[:3, :3] = Rotation.from_rotvec([0.3, -0.2, 0.9]).as_matrix()
    T[:3, 3] = [1.0, 2.0, 3.0]
    m = tu.SE32manifSE3(T)
    assert np.allclose(m.translation(), [1, 2, 3]) and abs(np.linalg.norm(m.quat()) - 1) < 1e-15
    assert np.max(np.abs(tu.manifSE32SE3(m) - T)) < 1e-15
    t = tu.se32manifse3(xi)
    assert np.array_equal(t.coeffs(), np.concatenate((xi[3:], xi[:3]))) and np.array_equal(tu.manifse32se3(t), xi)
    J = rng.standard_normal((6, 6))
    P = np.block([[np.zeros((3, 3)), np.eye(3)], [np.eye(3), np.zeros((3, 3))]])
    assert np.array_equal(tu.Jmnf2J(J), P @ J @ P)
    assert tu.is_pos_def(np.eye(3)) and not tu.is_pos_def(-np.eye(3)) and not tu.is_pos_def(np.array([[1.0, 2.0], [0.0, 1.0]]))
    assert abs(tu.SE32absangle(T) - np.rad2deg(np.linalg.norm([0.3, -0.2, 0.9]))) < 1e-10


def test_second_order_terms_fail_like_the_reference():
    """hessians=True: the reference's exact-dynamics classes never define `_f_xx` (traopt_dynamics.py:684-686), so its
    `fit` dies with AttributeError in `_linearization`; without the flag f_xx raises NotImplementedError (:852-866)."""
    import pytest
    from trajectory_optimization_matrix_lie_groups_b200.traoptlibrary import traopt_controller as tc, traopt_cost, traopt_dynamics
    J = np.diag([0.5, 0.7, 0.9, 1.0, 1.0, 1.0])
    q_ref, xi_ref = np.tile(np.eye(4), (6, 1, 1)), np.zeros((6, 6))
    plain = traopt_dynamics.SE3Dynamics(J, 0.01)
    with pytest.raises(NotImplementedError):
        plain.f_xx(None, None, 0)
    dyn = traopt_dynamics.SE3Dynamics(J, 0.01, hessians=True)
    assert dyn.has_hessians
    with pytest.raises(AttributeError):
        dyn.f_xx(None, None, 0)
    cost = traopt_cost.SE3TrackingQuadraticGaussNewtonCost(np.eye(12), np.eye(6), np.eye(12), q_ref, xi_ref)
    ctrl = tc.iLQR_Tracking_SE3_MS(dyn, cost, 5, q_ref, xi_ref, hessians=True)
    with pytest.raises(AttributeError):
        ctrl.fit([np.eye(4), np.zeros(6)], np.zeros((5, 6)), n_iterations=1)
    with pytest.warns(UserWarning):
        tc.iLQR_Tracking_SE3_MS(plain, cost, 5, q_ref, xi_ref, hessians=True)       # :2382-2383

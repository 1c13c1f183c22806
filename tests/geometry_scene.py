"""Deterministic synthetic sliding-window scene for the triangulation / outlier-rejection parity tests: WINDOW_SIZE + 1 = 11
body poses on a smooth trajectory, EuRoC-like camera-to-IMU extrinsics (11 cm stereo baseline), landmarks 2-30 m in front,
observations = normalised coordinates with a little noise, a few gross outliers, some landmarks without a right
observation, some already triangulated.  Layout = the flat landmark table of include/vins_fast_b200.h."""
import numpy as np


def rot(axis, a):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


def make_scene(seed=0, n=600, window=11):
    rng = np.random.default_rng(seed)
    Rs = np.stack([rot([0.2, 1.0, 0.1], 0.02 * k) @ rot([1, 0, 0], 0.01 * np.sin(k)) for k in range(window)])
    Ps = np.stack([[0.35 * k, 0.05 * np.sin(0.7 * k), 0.02 * k] for k in range(window)])
    ric = np.stack([rot([0, 0, 1], 0.01) @ rot([0, 1, 0], 0.02), rot([0, 0, 1], 0.012) @ rot([0, 1, 0], 0.018)])
    tic = np.array([[0.02, -0.06, 0.01], [0.02, 0.05, 0.01]])
    start, nobs, offs, rows, depth = [], [], [], [], []
    for i in range(n):
        s = int(rng.integers(0, window - 1))
        m = int(rng.integers(1, window - s + 1))
        # a world point in front of the first observing camera
        Rc0, tc0 = Rs[s] @ ric[0], Ps[s] + Rs[s] @ tic[0]
        pc = np.array([rng.uniform(-0.6, 0.6), rng.uniform(-0.4, 0.4), 1.0]) * rng.uniform(2.0, 30.0)
        pw = Rc0 @ pc + tc0
        offs.append(len(rows)); start.append(s); nobs.append(m)
        gross = rng.uniform() < 0.08
        for k in range(m):
            f = s + k
            row = np.zeros(7)
            for cam in (0, 1):
                Rc, tc = Rs[f] @ ric[cam], Ps[f] + Rs[f] @ tic[cam]
                q = Rc.T @ (pw - tc)
                uv = q[:2] / q[2] + rng.normal(0, 2e-4, 2) + (rng.normal(0, 0.02, 2) if gross and k > 0 else 0)
                if cam == 0:
                    row[0:3] = (uv[0], uv[1], 1.0)
                elif rng.uniform() < 0.8:
                    row[3] = 1.0
                    row[4:7] = (uv[0], uv[1], 1.0)
            rows.append(row)
        depth.append(float(rng.uniform(1, 20)) if rng.uniform() < 0.15 else -1.0)
    return dict(Ps=Ps, Rs=Rs.reshape(window, 9), tic=tic, ric=ric.reshape(2, 9), start_frame=np.array(start, np.int32),
                n_obs=np.array(nobs, np.int32), obs_offset=np.array(offs, np.int32), obs=np.array(rows), depth=np.array(depth))


def stereo_pairs(sc):
    """(pose0, pose1, point0, point1) of every landmark whose first frame has a right observation: TriangulateOnePoint inputs."""
    P0, P1, a, b = [], [], [], []
    Rs, ric = sc["Rs"].reshape(-1, 3, 3), sc["ric"].reshape(2, 3, 3)
    for i in range(len(sc["start_frame"])):
        row = sc["obs"][sc["obs_offset"][i]]
        if row[3] == 0:
            continue
        f = sc["start_frame"][i]
        poses = []
        for cam in (0, 1):
            R = Rs[f] @ ric[cam]
            t = sc["Ps"][f] + Rs[f] @ sc["tic"][cam]
            poses.append(np.hstack([R.T, (-R.T @ t)[:, None]]).reshape(12))
        P0.append(poses[0]); P1.append(poses[1]); a.append(row[0:2]); b.append(row[4:6])
    return np.array(P0), np.array(P1), np.array(a), np.array(b)

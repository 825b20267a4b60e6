"""TUM RGB-D on-disk formats around the path: the association file the reference reads
(LoadImages, Utils/utils.cpp:16-38) and the trajectory file it writes (Tracking::SaveTrajectory, System/tracking.cpp:544-580:
`timestamp tx ty tz qx qy qz qw`, fixed notation, 6 / 9 decimals).  Host-side I/O only — no arithmetic of the hot path lives here."""
import numpy as np


def load_associations(path):
    """LoadImages: every non-empty line is `t_rgb rgb_file t_depth depth_file`; the first timestamp is kept.
    Returns (timestamps float64 [n], rgb file names, depth file names)."""
    ts, rgb, dep = [], [], []
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            tok = line.split()
            ts.append(float(tok[0])); rgb.append(tok[1]); dep.append(tok[3])
    return np.array(ts, np.float64), rgb, dep


def quaternion_from_rotation(R):
    """Converter::toQuaternion (Utils/converter.cpp:149-161) = Eigen::Quaterniond(Matrix3d): (x, y, z, w) as float32."""
    m = np.asarray(R, np.float32).astype(np.float64)
    q = np.zeros(4)                                    # x y z w
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0:
        t = np.sqrt(t + 1.0); q[3] = 0.5 * t; t = 0.5 / t
        q[0] = (m[2, 1] - m[1, 2]) * t; q[1] = (m[0, 2] - m[2, 0]) * t; q[2] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]: i = 1
        if m[2, 2] > m[i, i]: i = 2
        j = (i + 1) % 3; k = (j + 1) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0); q[i] = 0.5 * t; t = 0.5 / t
        q[3] = (m[k, j] - m[j, k]) * t; q[j] = (m[j, i] + m[i, j]) * t; q[k] = (m[k, i] + m[i, k]) * t
    return q.astype(np.float32)


def camera_centre(Tcw):
    """twc = -Rwc * tcw as cv::Mat evaluates it (float products summed left to right), Rwc = Rcw^T."""
    T = np.asarray(Tcw, np.float32)
    Rwc = T[:3, :3].T
    twc = np.zeros(3, np.float32)
    for r in range(3):
        acc = np.float32(Rwc[r, 0] * T[0, 3]); acc = np.float32(acc + np.float32(Rwc[r, 1] * T[1, 3])); acc = np.float32(acc + np.float32(Rwc[r, 2] * T[2, 3]))
        twc[r] = -acc
    return Rwc, twc


def trajectory_lines(timestamps, poses_Tcw):
    out = []
    for t, T in zip(timestamps, poses_Tcw):
        Rwc, twc = camera_centre(T)
        q = quaternion_from_rotation(Rwc)
        out.append(f"{float(t):.6f} {float(twc[0]):.9f} {float(twc[1]):.9f} {float(twc[2]):.9f} {float(q[0]):.9f} {float(q[1]):.9f} {float(q[2]):.9f} {float(q[3]):.9f}")
    return out


def save_trajectory(path, timestamps, poses_Tcw):
    """One line per frame in the TUM trajectory format (what evaluate_ate.py / evaluate_rpe.py read)."""
    with open(path, "w") as f:
        for line in trajectory_lines(timestamps, poses_Tcw):
            f.write(line + "\n")


def load_trajectory(path):
    rows = [[float(v) for v in l.split()] for l in open(path) if l.strip() and not l.startswith("#")]
    a = np.array(rows, np.float64).reshape(-1, 8)
    return a[:, 0], a[:, 1:4], a[:, 4:8]

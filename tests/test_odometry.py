"""Odometry front-end against a literal restatement of the reference callback
(centralized_six_robots_implementation.py:20-37): quaternion z -> yaw, local frame -> global frame."""
import numpy as np


def _reference_callback(xr, yr, qz, init):
    """The callback body, line by line (3x3 rotation, homogeneous z = 0)."""
    x_i, y_i, th_i = init
    th = 2 * np.arcsin(qz)
    P_local = np.array([[xr], [yr], [0.0]])
    phi = th + th_i
    R = np.array([[np.cos(th_i), -np.sin(th_i), 0.0], [np.sin(th_i), np.cos(th_i), 0.0], [0.0, 0.0, 1.0]])
    Pg = np.matmul(R, P_local) + np.array([[x_i], [y_i], [0.0]])
    return np.array([Pg[0, 0], Pg[1, 0], phi])


def test_odom_to_state_matches_reference_callback(pkg):
    od = pkg.odometry
    rng = np.random.default_rng(3)
    Nr, B = 6, 5
    init = rng.uniform(-2, 2, (Nr, 3))
    x, y = rng.uniform(-1, 1, (B, Nr)), rng.uniform(-1, 1, (B, Nr))
    th = rng.uniform(-3.0, 3.0, (B, Nr))
    qz = np.sin(th / 2)
    st = od.odom_to_state(x, y, qz, init)
    assert st.shape == (B, 3 * Nr)
    for b in range(B):
        for i in range(Nr):
            np.testing.assert_allclose(st[b, 3 * i:3 * i + 3], _reference_callback(x[b, i], y[b, i], qz[b, i], init[i]), rtol=0, atol=1e-14)


def test_yaw_limitation_is_the_references(pkg):
    """2 arcsin(qz) cannot represent |yaw| > pi: a heading of 1.2 pi (qw < 0) comes back as 0.8 pi, as in the reference."""
    od = pkg.odometry
    th = 1.2 * np.pi
    assert abs(od.yaw_from_quaternion_z(np.sin(th / 2)) - 0.8 * np.pi) < 1e-12
    assert od.yaw_from_quaternion_z(1.0 + 1e-12) == np.pi      # clipped, no NaN from sensor noise

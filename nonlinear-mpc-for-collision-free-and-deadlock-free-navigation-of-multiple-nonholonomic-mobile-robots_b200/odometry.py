"""Odometry front-end of the reference's ROS scripts, as pure host code with no ROS dependency.

Each robot reports its pose in its own odometry frame (origin and heading = where it was switched on).  The scripts turn
that into the global pose the NMPC uses (centralized_six_robots_implementation.py:20-37, one callback per robot):

    th     = 2 arcsin(qz)                         yaw from the z component of the orientation quaternion
    P      = R(th_init) [x, y, 0]' + [x_init, y_init, 0]'
    pose   = [P_x, P_y, th + th_init]

The reference's formula is kept as it is, including its limitation: 2 arcsin(qz) ignores the sign of qw, so headings
beyond +-pi (qw < 0) alias.  Everything is vectorised over robots and over a leading batch dimension.
"""
import numpy as np


def yaw_from_quaternion_z(qz):
    """th = 2 arcsin(qz)   (centralized_six_robots_implementation.py:29).  qz is clipped to [-1, 1] against sensor noise."""
    return 2.0 * np.arcsin(np.clip(np.asarray(qz, dtype=np.float64), -1.0, 1.0))


def local_to_global(pose_local, pose_init):
    """pose_local [..., 3] = (x, y, th) in the robot's odometry frame, pose_init [..., 3] = that frame in the global one.
    Returns the global pose [..., 3]   (centralized_six_robots_implementation.py:30-36)."""
    pl = np.asarray(pose_local, dtype=np.float64)
    pi = np.asarray(pose_init, dtype=np.float64)
    c, s = np.cos(pi[..., 2]), np.sin(pi[..., 2])
    out = np.empty(np.broadcast_shapes(pl.shape, pi.shape), dtype=np.float64)
    out[..., 0] = c * pl[..., 0] - s * pl[..., 1] + pi[..., 0]
    out[..., 1] = s * pl[..., 0] + c * pl[..., 1] + pi[..., 1]
    out[..., 2] = pl[..., 2] + pi[..., 2]
    return out


def odom_to_state(x, y, qz, pose_init):
    """The callbacks' whole computation for Nr robots: odometry readings x, y, qz (each [..., Nr]) and the initial poses
    [..., Nr, 3] -> the stacked NMPC state [..., 3 Nr] = (x_1, y_1, th_1, ..., x_Nr, y_Nr, th_Nr), i.e. the vector the
    scripts assemble as x0 after every solve (centralized_six_robots_implementation.py:451)."""
    local = np.stack([np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64), yaw_from_quaternion_z(qz)], axis=-1)
    g = local_to_global(local, pose_init)
    return g.reshape(g.shape[:-2] + (-1,))

"""Sensor sequences of the reference's host API: Scan / Lidar (lidar.py:13-135)
and IMU (imu.py:9-26).  Same constructors, indexing and accessors; a Scan built
from ranges keeps them so the GPU path can upload the raw sweep."""
from collections.abc import Sequence
from math import cos, sin

import numpy as np

from .models import Pose, Position, Reading

_INT = (int, np.integer)


class Scan:
    """Cartesian beam endpoints of one sweep in the sensor frame."""

    def __init__(self, ranges, angles, timestamp=0):
        self._x = np.array([])
        self._y = np.array([])
        self._ranges = self._angles = None
        if isinstance(ranges, np.ndarray):
            self._ranges = np.asarray(ranges, dtype=np.float64)
            self._angles = np.asarray(angles, dtype=np.float64)
            self._x = np.array([r * cos(a) for r, a in zip(self._ranges, self._angles)])   # lidar.py:78
            self._y = np.array([r * sin(a) for r, a in zip(self._ranges, self._angles)])   # lidar.py:79
        self._timestamp = timestamp

    def x(self):
        return self._x

    def y(self):
        return self._y

    def ranges(self):
        return self._ranges

    def angles(self):
        return self._angles

    def timestamp(self):
        return self._timestamp

    def __getitem__(self, idx):
        if not isinstance(idx, _INT):
            raise Exception("Invalid attribute: " + str(idx))
        return Position(self._x[idx], self._y[idx])

    def __len__(self):
        return len(self._x)

    def __str__(self):
        return "Scan Class: %d points at timestamp: %s" % (len(self._x), self._timestamp)

    def from_global_reference(self, frame: Pose):
        """Endpoints in the global frame for a sensor at `frame` (lidar.py:111-128)."""
        c, s = cos(frame.theta()), sin(frame.theta())
        t = np.array([[c, -s, frame.x()], [s, c, frame.y()], [0.0, 0.0, 1.0]])
        pts = np.vstack((self._x, self._y, np.ones(len(self._x))))
        g = t @ pts
        out = Scan(None, None, self._timestamp)
        out._x, out._y = g[0], g[1]
        return out


class Lidar(Sequence):
    """Sequence of scans over a LidarData loader (lidar.py:13-51)."""

    def __init__(self, data, engine=None):
        self._scans = data.get_scans()
        self._times = data.get_times()
        self._angles = data.get_angles()
        self._matlab = engine                      # kept for signature compatibility; unused

    def __getitem__(self, idx):
        if not isinstance(idx, _INT):
            raise Exception("Invalid attribute: " + str(idx))
        return Scan(self._scans[idx], self._angles, self._times[idx])

    def timestamp_for_idx(self, idx):
        if not isinstance(idx, _INT):
            raise Exception("Invalid attribute: " + str(idx))
        return self._times[idx]

    def get_at_time(self, timestamp):
        idx = self._times.searchsorted(timestamp)
        if idx == 0 or idx == len(self._times):
            return None
        return self[int(idx)]

    def angles(self, idx):
        return self._angles[idx]

    def __len__(self):
        return len(self._scans)

    def __str__(self):
        return "Lidar Class: %d scans" % len(self._scans)


class IMU(Sequence):
    """Sequence of odometry readings over an IMUData loader (imu.py:9-26)."""

    def __init__(self, data):
        self._data = data.get_data()
        self._times = data.get_times()
        self._progress_fnc = data.progress_pose
        self._get_cov_input_uncertainty = data.get_cov_input_uncertainty
        self._get_cov_change_matrix = data.get_cov_change_matrix
        self._motion = getattr(data, "MOTION", None)

    def __getitem__(self, idx):
        if not isinstance(idx, _INT):
            raise Exception("Invalid attribute: %s (%s)" % (idx, type(idx)))
        return Reading(self._data[idx], self._times[idx], self._progress_fnc, self._get_cov_change_matrix,
                       self._get_cov_input_uncertainty, motion=self._motion)

    def __len__(self):
        return len(self._data)

    def __str__(self):
        return "IMU Class: %d readings" % len(self._data)

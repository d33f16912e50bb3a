"""Host value types that cross the reference's API: Position, Pose, Reading
(models.py:11-77 of the reference) and the 1e-4 s timestamp helpers
(models.py:5-9).  Pure data; no arithmetic of the hot path lives here.
"""
from dataclasses import dataclass
from typing import Any, Callable

TICKS_PER_SECOND = 1e4


def timestamp_to_time(timestamp):
    return timestamp / TICKS_PER_SECOND


def time_to_timestamp(time):
    return round(TICKS_PER_SECOND * time)


@dataclass
class Position:
    x: Any = 0
    y: Any = 0

    def __str__(self):
        return "(%s, %s)" % (self.x, self.y)


class Pose:
    """(x, y, theta) with the reference's accessor-method style (pose.x())."""

    __slots__ = ("_x", "_y", "_theta")

    def __init__(self, x=0.0, y=0.0, theta=0.0):
        self._x, self._y, self._theta = x, y, theta

    def x(self):
        return self._x

    def y(self):
        return self._y

    def theta(self):
        return self._theta

    def pos(self):
        return Position(self._x, self._y)

    def as_tuple(self):
        return (self._x, self._y, self._theta)

    def __str__(self):
        return "Pose: (%s, %s, %s)" % (self._x, self._y, self._theta)


class Reading:
    """One odometry/IMU sample plus the loader's three motion-model callbacks
    (reference models.py:44-77).  `motion` is what the GPU path consumes: the
    (family, noise parameters) pair of the loader that produced the reading."""

    def __init__(self, data, timestamp, progress_fnc: Callable, get_cov_change_matrix_fnc: Callable,
                 get_cov_input_uncertainty: Callable, motion=None):
        self._data = data
        self._timestamp = timestamp
        self._progress_fnc = progress_fnc
        self._get_cov_change_matrix_fnc = get_cov_change_matrix_fnc
        self._get_cov_input_uncertainty = get_cov_input_uncertainty
        self._dt = 0.0
        self.motion = motion

    def dt(self):
        return self._dt

    def set_dt(self, dt):
        self._dt = dt

    def timestamp(self):
        return self._timestamp

    def get_data(self):
        return self._data

    def get_moved_pose(self, pose):
        return self._progress_fnc(pose, self)

    def get_cov_change_matrix(self, pose):
        return self._get_cov_change_matrix_fnc(pose, self)

    def get_cov_input_uncertainty(self, pose):
        return self._get_cov_input_uncertainty(pose, self)

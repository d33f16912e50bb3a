"""Log loaders of the reference's host API (LidarData.py / IMUData.py and the
nine *LidarData.py / *IMUData.py pairs), table-driven.

Same class names, same `load_and_format()` results (times in 1e-4 s ticks, scans
[n, B] in metres, angles [B]; odometry rows + times) and the same three motion
callbacks per IMU loader, so `Lidar(IntelLidarData(), eng)` / `IMU(IntelIMUData())`
work as in main.py:84-85.  Each IMU loader additionally carries
MOTION = (family, noise parameters), which is what the CUDA motion kernel needs.

Parsing happens once on the host and is not part of the timed path.  `data_dir`
(default "./data", like the reference) lets tests point at fixtures.
"""
import os
from math import cos, pi, sin

import numpy as np

from .models import Pose

MOTION_ABSOLUTE, MOTION_VELOCITY, MOTION_UNICYCLE = 0, 1, 2
DATA_DIR = os.environ.get("THESIS_DATA_DIR", "./data")


def _carmen_rows(path, tag, with_lineno=False):
    with open(path) as fh:
        lines = fh.read().splitlines()
    if with_lineno:
        return [ln.split() + [str(n)] for n, ln in enumerate(lines) if ln.startswith(tag)]
    return [ln.split() for ln in lines if ln.startswith(tag)]


# ----------------------------------------------------------------- lidar --

class LidarData:
    """Loader base (LidarData.py:7-23)."""

    def __init__(self, data_dir=None):
        self._dir = data_dir or DATA_DIR
        self._times, self._scans, self._angles = self.load_and_format()

    def load_and_format(self):
        raise NotImplementedError

    def get_times(self):
        return self._times

    def get_scans(self):
        return self._scans

    def get_angles(self):
        return self._angles


class _CarmenLidar(LidarData):
    FILE = None
    BEAMS = 180
    DEDUP = True            # np.unique(times, return_index=True) like the *Raw / Freiburg loaders

    def _ticks(self, rows):
        return np.array([int(1000 * float(r[-1])) * 10 for r in rows])

    def load_and_format(self):
        rows = _carmen_rows(os.path.join(self._dir, self.FILE), "FLASER")
        scans = np.array([[float(v) for v in r[2:self.BEAMS + 2]] for r in rows])
        times = self._ticks(rows)
        angles = np.array([-pi / 2 + i * pi / (self.BEAMS - 1) for i in range(self.BEAMS)])
        if self.DEDUP:
            times, keep = np.unique(times, return_index=True)
            scans = scans[keep]
        return times, scans, angles


class IntelLidarData(_CarmenLidar):          # IntelLidarData.py:12-20: 0.1 s time quantisation, no de-duplication
    FILE, BEAMS, DEDUP = "intel.txt", 180, False

    def _ticks(self, rows):
        return np.array([int(10 * float(r[-3])) * 10 for r in rows])


class IntelRawLidarData(_CarmenLidar):
    FILE, BEAMS = "intel_raw.log", 180


class AcesLidarData(_CarmenLidar):
    FILE, BEAMS = "aces.txt", 180


class FreidLidarData(_CarmenLidar):
    FILE, BEAMS = "fr.txt", 360


class FreidCorrectLidarData(_CarmenLidar):
    FILE, BEAMS = "fr_correct.log", 360


class Freid101LidarData(_CarmenLidar):
    FILE, BEAMS = "fr101.log", 360


class OberoLidarData(_CarmenLidar):
    """OberoLidarData.py:8,16 asks for 360 readings per FLASER record, but the committed data/orebro.log
    holds 181 (a 180-degree SICK at 1 degree): the reference's loader runs into the record's trailing
    fields and dies on float('pippo').  Every FLASER record of that file is stamped 0 and the ODOM
    stamps are not times either, so the reference's np.unique would keep one sweep.  Declared
    deviations: the beam count is the one the record states and records are ordered by line number,
    0.1 s apart (like the CSAIL loader)."""
    FILE, BEAMS = "orebro.log", 360

    def load_and_format(self):
        rows = _carmen_rows(os.path.join(self._dir, self.FILE), "FLASER", with_lineno=True)
        self.BEAMS = int(rows[0][1])
        scans = np.array([[float(v) for v in r[2:self.BEAMS + 2]] for r in rows])
        times = np.array([int(r[-1]) * 1000 for r in rows])
        angles = np.array([-pi / 2 + i * pi / (self.BEAMS - 1) for i in range(self.BEAMS)])
        return times, scans, angles


class BeleLidarData(_CarmenLidar):
    FILE, BEAMS = "bele.log", 361

    def _ticks(self, rows):                  # BeleLidarData.py:17 (offset subtracted before scaling, kept)
        t0 = float(rows[0][-1])
        return np.array([int(1000 * float(r[-1]) - t0) * 10 for r in rows])


class CsailLidarData(_CarmenLidar):
    """data/csail_correct.log has no loader in the reference (SURVEY 8c data
    gaps); this one follows the FreidCorrect pattern with 361 beams."""
    FILE, BEAMS = "csail_correct.log", 361

    def load_and_format(self):
        # the log's stamps are printed as 1.13486e+09 (all equal): order records by
        # their line number instead, 0.1 s apart
        rows = _carmen_rows(os.path.join(self._dir, self.FILE), "FLASER", with_lineno=True)
        scans = np.array([[float(v) for v in r[2:self.BEAMS + 2]] for r in rows])
        times = np.array([int(r[-1]) * 1000 for r in rows])
        angles = np.array([-pi / 2 + i * pi / (self.BEAMS - 1) for i in range(self.BEAMS)])
        return times, scans, angles


class DefaultLidarData(LidarData):           # DefaultLidarData.py:10-19, UNSW .mat, 13-bit centimetre ranges
    def load_and_format(self):
        from scipy.io import loadmat

        m = loadmat(os.path.join(self._dir, "lidar"))
        raw = m["dataL"]["Scans"][0][0]
        scans = np.array([0.01 * (col & 0x1FFF) for col in raw]).transpose()
        times = np.array([t * 1e4 for t in m["dataL"]["times"][0][0][0]])
        angles = np.array([-pi / 2 + i * pi / 360 for i in range(361)])
        return times, scans, angles


# ------------------------------------------------------------------- imu --

class IMUData:
    """Loader base (IMUData.py:9-40) incl. its shape asserts."""
    MOTION = None

    def __init__(self, data_dir=None):
        self._dir = data_dir or DATA_DIR
        self._data, self._times = self.load_and_format()
        assert self._data.shape[0] > 2
        assert self._data.shape[1] < 5
        assert len(self._times.shape) == 1

    def load_and_format(self):
        raise NotImplementedError

    def get_data(self):
        return self._data

    def get_times(self):
        return self._times


class IntelIMUData(IMUData):
    """Absolute-set odometry (IntelIMUData.py:9-36), callbacks' contents swapped
    exactly as in the reference."""
    MOTION = (MOTION_ABSOLUTE, (0.0, 0.0, 0.0, 0.0))
    FILE = "intel.txt"

    def load_and_format(self):
        rows = _carmen_rows(os.path.join(self._dir, self.FILE), "ODOM")
        vals = np.array([[float(v) for v in r[1:4] + [r[7]]] for r in rows])
        times = np.array([int(10 * v[3]) * 10 for v in vals])
        return vals[:, :3].copy(), times

    @staticmethod
    def progress_pose(prev_pose, reading):
        d = reading.get_data()
        return Pose(d[0], d[1], d[2])

    @staticmethod
    def get_cov_input_uncertainty(prev_pose, reading):
        q = np.diag([1.0, 1.0, 1.0])
        q[0][2] = reading.get_data()[0] - prev_pose.x()
        q[1][2] = reading.get_data()[1] - prev_pose.y()
        return q

    @staticmethod
    def get_cov_change_matrix(prev_pose, reading):
        return np.abs(np.diag([0.01 ** 2, 0.01 ** 2, (0.2 * pi / 180) ** 2]))


class _VelocityIMU(IMUData):
    """Additive-velocity odometry: poses differenced into world-frame velocities
    (IntelRawIMUData.py:10-55; same shape in Aces/Freid*/Obero/Bele)."""
    FILE = None
    CALIB = 5                  # readings averaged for the zero offset
    FLIP_X = False             # FreidIMUData.py:16 negates x
    REL_TIME = False           # BeleIMUData.py:18 subtracts the first stamp
    NOISE = (0.02, 0.01, 0.2 * pi / 180, 0.02)     # a_xy, b_xy, a_th, b_th

    def load_and_format(self):
        rows = _carmen_rows(os.path.join(self._dir, self.FILE), "ODOM")
        vals = np.array([[float(v) for v in r[1:4] + [r[9]]] for r in rows])
        if self.FLIP_X:
            vals[:, 0] = -vals[:, 0]
        t0 = vals[0][3] if self.REL_TIME else 0.0
        ticks = np.array([int(1000 * (v[3] - t0)) * 10 for v in vals])
        times, keep = np.unique(ticks, return_index=True)
        pos = vals[:, :3] - np.mean(vals[0:self.CALIB, :3], axis=0)
        vel = 1e4 * np.diff(pos[keep], axis=0) / np.diff(np.column_stack((times, times, times)), axis=0)
        return np.vstack(([0.0, 0.0, 0.0], vel)), times

    @staticmethod
    def progress_pose(prev_pose, reading):
        d, dt = reading.get_data(), reading.dt() / 1e4
        return Pose(prev_pose.x() + d[0] * dt, prev_pose.y() + d[1] * dt, prev_pose.theta() + d[2] * dt)

    @staticmethod
    def get_cov_change_matrix(prev_pose, reading):
        return np.diag([1.0, 1.0, 1.0])

    @classmethod
    def _noise(cls, reading):
        a, b, at, bt = cls.NOISE
        d, dt = reading.get_data(), reading.dt() / 1e4
        return np.abs(np.diag([(a + b * abs(d[0]) * dt) ** 2, (a + b * abs(d[1]) * dt) ** 2,
                               (at + bt * abs(d[2]) * dt) ** 2]))


def _velocity_loader(name, file, noise=_VelocityIMU.NOISE, calib=5, flip_x=False, rel_time=False):
    def get_cov_input_uncertainty(prev_pose, reading, _noise=noise):
        a, b, at, bt = _noise
        d, dt = reading.get_data(), reading.dt() / 1e4
        return np.abs(np.diag([(a + b * abs(d[0]) * dt) ** 2, (a + b * abs(d[1]) * dt) ** 2,
                               (at + bt * abs(d[2]) * dt) ** 2]))

    return type(name, (_VelocityIMU,), dict(
        FILE=file, NOISE=noise, CALIB=calib, FLIP_X=flip_x, REL_TIME=rel_time,
        MOTION=(MOTION_VELOCITY, noise), get_cov_input_uncertainty=staticmethod(get_cov_input_uncertainty),
        __doc__="Velocity-family odometry over data/%s." % file))


IntelRawIMUData = _velocity_loader("IntelRawIMUData", "intel_raw.log", noise=(0.002, 0.05, 0.01 * pi / 180, 0.05))
AcesIMUData = _velocity_loader("AcesIMUData", "aces.txt", calib=1)
FreidIMUData = _velocity_loader("FreidIMUData", "fr.log", flip_x=True)
FreidCorrectIMUData = _velocity_loader("FreidCorrectIMUData", "fr_correct.log", flip_x=True)
Freid101IMUData = _velocity_loader("Freid101IMUData", "fr101.log")
BeleIMUData = _velocity_loader("BeleIMUData", "bele.log", rel_time=True)


class CsailIMUData(_VelocityIMU):
    """Odometry of data/csail_correct.log, stamped by line number like CsailLidarData."""
    FILE = "csail_correct.log"
    MOTION = (MOTION_VELOCITY, _VelocityIMU.NOISE)

    def load_and_format(self):
        rows = _carmen_rows(os.path.join(self._dir, self.FILE), "ODOM", with_lineno=True)
        vals = np.array([[float(v) for v in r[1:4]] for r in rows])
        times = np.array([int(r[-1]) * 1000 for r in rows])
        pos = vals - np.mean(vals[0:self.CALIB], axis=0)
        vel = 1e4 * np.diff(pos, axis=0) / np.diff(np.column_stack((times, times, times)), axis=0)
        return np.vstack(([0.0, 0.0, 0.0], vel)), times

    @staticmethod
    def get_cov_input_uncertainty(prev_pose, reading):
        return CsailIMUData._noise(reading)


class OberoIMUData(CsailIMUData):
    """Odometry of data/orebro.log (OberoIMUData.py): its ODOM stamps are not times (6.979, 1.961,
    3.4e-86, ...), so records are ordered by line number like OberoLidarData (declared)."""
    FILE = "orebro.log"

    @staticmethod
    def get_cov_input_uncertainty(prev_pose, reading):
        return OberoIMUData._noise(reading)


class DefaultIMUData(IMUData):
    """Unicycle model on UNSW speed + gyro .mat logs (DefaultIMUData.py:8-54)."""
    MOTION = (MOTION_UNICYCLE, (0.0, 0.0, 0.0, 0.0))
    NUM_REF_POINTS = 1000

    def load_and_format(self):
        from scipy.io import loadmat

        imu = loadmat(os.path.join(self._dir, "imu"))
        enc = loadmat(os.path.join(self._dir, "speed"))
        w_raw = imu["IMU"]["DATAf"][0][0][5]
        v_raw = enc["Vel"]["speeds"][0][0][0]
        w0 = sum(w_raw[0:self.NUM_REF_POINTS]) / self.NUM_REF_POINTS
        v0 = sum(v_raw[0:self.NUM_REF_POINTS]) / self.NUM_REF_POINTS
        omega = np.array([x - w0 for x in w_raw])
        speed = np.array([x - v0 for x in v_raw])
        times = np.array([t * 1e4 for t in imu["IMU"]["times"][0][0][0]])
        return np.vstack((speed, omega)).transpose(), times

    @staticmethod
    def progress_pose(prev_pose, reading):
        dt, d = reading.dt() / 1e4, reading.get_data()
        th = prev_pose.theta() + dt * d[1]
        return Pose(prev_pose.x() + dt * d[0] * cos(th), prev_pose.y() + dt * d[0] * sin(th), th)

    @staticmethod
    def get_cov_change_matrix(prev_pose, reading):
        f = np.diag([1.0, 1.0, 1.0])
        dt, d = reading.dt() / 1e4, reading.get_data()
        f[0][2] = dt * d[0] * cos(prev_pose.theta())
        f[1][2] = dt * d[0] * sin(prev_pose.theta())
        return f

    @staticmethod
    def get_cov_input_uncertainty(prev_pose, reading):
        dt = reading.dt() / 1e4
        g = np.array([[dt * cos(prev_pose.theta()), 0], [dt * sin(prev_pose.theta()), 0], [0, dt]])
        m = np.diag([0.05 ** 2, (pi / 180 / 2) ** 2])
        return np.abs(g @ m @ g.T) + np.abs(np.diag([0.01 ** 2, 0.01 ** 2, (0.2 * pi / 180) ** 2]))

"""Headless restatement of the reference's event loop (main.py:81-166, SURVEY
Appendix B): merge IMU and lidar timestamps, fan imu_update / map_update over the
particles, gate updates on particle 0's motion, resample.  Plotting
(main.py:101-108,167-182) and the shelve checkpoint (main.py:117-136,183-210) are
left out; loader timestamps are de-duplicated first.

Works with any object set that has the reference's particle API -- the GPU views
of thesis_b200.particles, or the reference's own Robot class in the tests.
"""
from math import pi, sqrt

import numpy as np

MAX_UPDATE_COUNT = 2        # main.py:41
ROT_THRESHOLD = pi / 9      # main.py:42
DIST_THRESHOLD = 0.33       # main.py:43


def dedup_times(seq):
    """np.unique(..., return_index=True) on a Lidar / IMU sequence's time axis."""
    t, keep = np.unique(seq._times, return_index=True)
    seq._times = t
    if hasattr(seq, "_scans"):
        seq._scans = seq._scans[keep]
    else:
        seq._data = seq._data[keep]
    return seq


def _fan(particles, call):
    """main.py:144,157-159 fan a call over the particle list.  The device-backed views of
    thesis_b200.particles run the batched kernels for every particle on the first call of a
    round, so one call is enough there (65,536 Python calls per stage would cost more than
    the kernels)."""
    if getattr(particles[0], "_shared", None) is not None:
        call(particles[0])
    else:
        [call(p) for p in particles]


def run_log(particles, lidar_data, imu_data, resample_fn, seed_fn=None, max_frames=None, frame0=0,
            on_frame=None):
    """Drive `particles` through the log.  Returns (particles, log) where log is a
    list of per-lidar-frame dicts (frame, updated, pose of particle 0)."""
    dedup_times(lidar_data)
    dedup_times(imu_data)
    if seed_fn is not None:
        seed_fn(particles, lidar_data[0])                         # main.py:89-90 (commented out there)
    prev_timestamp = imu_data[0].timestamp()
    imu_idx = lidar_idx = 0
    frame = frame0                                                # plotFrameNumber, main.py:109
    last_updated_pose = particles[0].get_latest_pose()
    last_scan = lidar_data[0].from_global_reference(last_updated_pose)
    update_count = 0
    times = np.unique(np.concatenate((imu_data._times, lidar_data._times)))   # main.py:114
    log = []
    for t in times:
        reading = imu_data[imu_idx]
        if reading.timestamp() == t:                              # main.py:139-145
            imu_idx = min(imu_idx + 1, len(imu_data) - 1)
            reading.set_dt(reading.timestamp() - prev_timestamp)
            _fan(particles, lambda p: p.imu_update(reading))
            prev_timestamp = reading.timestamp()
        if lidar_data.timestamp_for_idx(lidar_idx) == t:          # main.py:147-180
            scan = lidar_data[lidar_idx]
            lidar_idx = min(lidar_idx + 1, len(lidar_data) - 1)
            curr = particles[0].get_latest_pose()
            dist = sqrt((last_updated_pose.x() - curr.x()) ** 2 + (last_updated_pose.y() - curr.y()) ** 2)
            rot = abs(last_updated_pose.theta() - curr.theta())
            updated = False
            if update_count < MAX_UPDATE_COUNT or dist >= DIST_THRESHOLD or rot >= ROT_THRESHOLD:
                adj = not (frame % 5 < 2)                         # main.py:156-159
                _fan(particles, lambda p: p.map_update(scan, last_scan, adj))
                particles = resample_fn(particles)
                if dist >= DIST_THRESHOLD or rot >= ROT_THRESHOLD:
                    update_count = 0
                    last_updated_pose = curr
                elif update_count < MAX_UPDATE_COUNT:
                    update_count += 1
                if frame % 5 == 0:                                # main.py:167-168
                    last_scan = scan.from_global_reference(particles[0].get_latest_pose())
                updated = True
            p0 = particles[0].get_latest_pose()
            log.append(dict(frame=frame, updated=updated, pose=(p0.x(), p0.y(), p0.theta())))
            if on_frame is not None:
                on_frame(frame, particles)
            frame += 1                                            # main.py:180
            if max_frames is not None and frame - frame0 >= max_frames:
                break
    return particles, log

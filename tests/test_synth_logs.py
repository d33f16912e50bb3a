"""CPU: a synthetic CARMEN log written by thesis_b200.synth parses through the
reference-style loaders (stand-ins for the logs missing from the reference tree:
aces.txt, fr.log, ...) and drives the headless loop with the oracle."""
import numpy as np

import ref_adapter as RA
from thesis_b200 import harness, loaders, sensors, synth


def test_carmen_roundtrip_through_aces_loaders(tmp_path):
    w = synth.Workload(8, n_beams=180)
    synth.write_carmen_log(str(tmp_path / "aces.txt"), w)
    ld = sensors.Lidar(loaders.AcesLidarData(str(tmp_path)))
    im = sensors.IMU(loaders.AcesIMUData(str(tmp_path)))
    assert len(ld) == 8 and len(ld[0]) == 180
    assert np.allclose(ld._scans, np.round(w.ranges, 3))
    assert im[0].motion[0] == loaders.MOTION_VELOCITY
    # velocities integrate back to the odometry increments of the workload
    dt = np.diff(im._times) / 1e4
    total = np.sum(im._data[1:] * dt[:, None], axis=0)
    assert np.allclose(total, np.sum(w.odom, axis=0), atol=1e-4)


def test_oracle_runs_synthetic_freiburg_style_log(tmp_path):
    w = synth.Workload(6, n_beams=360)
    synth.write_carmen_log(str(tmp_path / "fr.txt"), w)
    synth.write_carmen_log(str(tmp_path / "fr.log"), w)           # the reference's two loaders disagree on the name
    ld = sensors.Lidar(loaders.FreidLidarData(str(tmp_path)))
    im = sensors.IMU(loaders.FreidIMUData(str(tmp_path)))
    np.random.seed(1)
    op = RA.OracleParticles(2, 360)
    parts, log = harness.run_log(op.views, ld, im, op.resample, seed_fn=op.seed, max_frames=4)
    assert len(log) == 4 and all(np.isfinite(l["pose"]).all() for l in log)
    # FreidIMUData negates x (FreidIMUData.py:16): the particle moves along -x
    assert parts[0].get_latest_pose().x() < 0

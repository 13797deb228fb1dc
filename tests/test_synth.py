"""CPU: the synthetic sequence generator is deterministic and geometrically sane."""
import numpy as np

from infinitam_b200 import synth


def test_deterministic_and_in_range():
    a = synth.sequence(2, 160, 120, noise=True)
    b = synth.sequence(2, 160, 120, noise=True)
    assert a.dtype == np.int16 and a.shape == (2, 120, 160) and np.array_equal(a, b)
    clean = synth.sequence(1, 160, 120)[0]
    assert clean.min() > 1000 and clean.max() <= 2700  # room is 4 m deep, camera ~0.6 m behind the centre
    assert (a == 0).mean() > 0.002  # dropped pixels exist in the noisy variant


def test_ground_truth_pose_is_rigid_and_starts_at_identity():
    assert np.allclose(synth.ground_truth_pose(0), np.eye(4))
    M = synth.ground_truth_pose(10)
    R = M[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-12
    step = np.linalg.norm(synth.ground_truth_pose(11)[:3, 3] - M[:3, 3])
    assert 0.002 < step < 0.03  # a few mm to cm per frame: inside the ICP basin

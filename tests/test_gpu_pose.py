"""GPU parity for row f4: surf_ransac_plane (hypotheses scored by k_ransac_score) against the oracle restatement of
vtkSlicerSurfFeaturesLogic::ransac (...Logic.cxx:2085-2188) -- indices, error sums and generator state identical --
and the whole chain matches -> points -> plane -> pose against the oracle."""
import numpy as np
import pytest

from oracle import pose_oracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from surffeatures_b200 import SURF

    e = SURF(100, 4, 3, max_width=64, max_height=64, max_batch=1, max_keypoints=256)
    yield e
    e.close()


def _cloud(n, seed, outlier_every=4):
    g = np.random.default_rng(seed)
    xy = g.uniform(-60, 60, size=(n, 2))
    pts = np.column_stack([xy, 0.3 * xy[:, 0] + 0.1 * xy[:, 1] + 7 + g.normal(size=n) * 0.6])
    pts[::outlier_every] += g.normal(size=pts[::outlier_every].shape) * 15
    return pts


@pytest.mark.parametrize("n,margin,iters,seed", [(4, 1.0, 1000, 1), (9, 1.0, 1000, 1), (150, 1.0, 1000, 1), (150, 2.5, 1000, 77),
                                                 (2000, 1.5, 1000, 1), (37, 0.0, 300, 5)])
def test_ransac_plane_equals_oracle(eng, n, margin, iters, seed):
    from surffeatures_b200 import pose

    pts = _cloud(n, n + seed)
    orng, prng = po.Rand(seed), pose.Rand(seed)
    rc, oinl, oplane, oerr, onout = po.ransac_plane(orng, pts, margin=margin, iterations=iters)
    assert rc == 0
    inl, plane, err, nout = pose.ransac_plane(eng, prng, pts, margin=margin, iterations=iters)
    assert inl.tolist() == oinl.tolist()
    assert plane.tolist() == oplane.tolist()
    assert err == oerr and nout == onout                    # bit-equal fp64 error sum
    assert prng.next() == orng.next()                       # same number of rand() calls consumed


def test_ransac_plane_degenerate_points_and_edge_cases(eng):
    from surffeatures_b200 import SurfError, pose

    # duplicated points: hypotheses that pick two coincident points are skipped by both
    pts = _cloud(30, 3)
    pts[10:20] = pts[0]
    orng, prng = po.Rand(1), pose.Rand(1)
    rc, oinl, oplane, oerr, onout = po.ransac_plane(orng, pts, iterations=400)
    inl, plane, err, nout = pose.ransac_plane(eng, prng, pts, iterations=400)
    assert rc == 0 and inl.tolist() == oinl.tolist() and plane.tolist() == oplane.tolist() and err == oerr and nout == onout
    # exactly three points: no draw, all three are the plane and the inliers
    inl, plane, err, nout = pose.ransac_plane(eng, pose.Rand(1), pts[:3])
    assert inl.tolist() == [0, 1, 2] and plane.tolist() == [0, 1, 2]
    # fewer than three (the reference returns 1), and no valid hypothesis at all (the reference reads an empty vector)
    with pytest.raises(SurfError):
        pose.ransac_plane(eng, pose.Rand(1), pts[:2])
    with pytest.raises(SurfError):
        pose.ransac_plane(eng, pose.Rand(1), np.ones((6, 3)), iterations=50)


def test_matches_to_pose_chain_equals_oracle(eng):
    """updateMatchNodeRansac's order of calls: match points, ransac (1000 iterations, margin 1.0), pose, distances."""
    from surffeatures_b200 import pose
    from test_oracle_pose import _scene

    m, qk, tk, off, T, I2P = _scene(n=55, seed=11)
    orng, prng = po.Rand(1), pose.Rand(1)
    otp, oqp = po.match_points(m, qk, tk, off, T, I2P)
    tp, qp = pose.match_points(m, qk, tk, off, T, I2P)
    assert np.array_equal(tp, otp) and np.array_equal(qp, oqp)
    rc, oinl, oplane, _, _ = po.ransac_plane(orng, otp, margin=1.0, iterations=1000)
    inl, plane, _, _ = pose.ransac_plane(eng, prng, tp, margin=1.0, iterations=1000)
    assert rc == 0 and inl.tolist() == oinl.tolist() and plane.tolist() == oplane.tolist()
    oest, ommd = po.estimate_pose(orng, oqp, otp, oinl, oplane)
    est, mmd = pose.estimate_pose(prng, qp, tp, inl, plane)
    # host fp64 with the same libm on both sides: equal; the stated tolerance covers another libm
    assert np.allclose(est, oest, rtol=1e-9, atol=1e-12) and abs(mmd - ommd) <= 1e-9 * max(1.0, abs(ommd))
    assert prng.next() == orng.next()

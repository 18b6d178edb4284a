"""MapMaker::AddPointEpipolar on the device (ptam_tracker_epipolar_search; SURVEY 8f rank 3, second half)
against the CPU oracle and against the reference itself (oracle/_ref).  Bar: the accepted set and the
winning corner of every candidate bit-exact; sub-pixel positions within 1e-9 (f64 warp reductions)."""
import numpy as np
import pytest

from ptam_cg_b200 import synth
from ptam_cg_b200.capi import Tracker, product_lib
from oracle.binding import oracle_lib, ref_lib

pytestmark = pytest.mark.gpu


def _case(lib, W, H, frames, poses, i_src, i_tgt, depth=(1.0, 0.3), wiggle=0.1, cands=None):
    t = Tracker(lib, W, H, 1)
    kf = t.add_keyframe(frames[i_src])
    if cands is None:
        t.make_keyframes([frames[i_src]])
        cands = [r[1] for r in t.keyframe_rest(0, 70.0)]
    t.make_keyframes([frames[i_tgt]])
    return cands, [t.epipolar_search(0, l, kf, poses[i_src], depth[0], depth[1], poses[i_tgt], wiggle, cands[l]) for l in range(4)]


@pytest.mark.parametrize("size", [(320, 240), (640, 480)], ids=["320x240", "640x480"])
@pytest.mark.parametrize("pair", [(0, 30), (30, 0), (10, 25)], ids=lambda p: f"src{p[0]}-tgt{p[1]}")
def test_epipolar_search_parity(size, pair):
    W, H = size
    frames, poses = synth.render_sequence(W, H, 40)
    cands, o = _case(oracle_lib(), W, H, frames, poses, *pair)
    _, p = _case(product_lib(), W, H, frames, poses, *pair, cands=cands)
    total = 0
    for l in range(4):
        (fo, bo, so), (fp, bp, sp) = o[l], p[l]
        assert np.array_equal(fo, fp), f"accepted set differs on level {l}"
        assert np.array_equal(bo, bp), f"winning corners differ on level {l}"
        np.testing.assert_allclose(sp, so, rtol=0, atol=1e-9)
        total += int(fo.sum())
    assert total > 200


def test_epipolar_search_early_exits_and_ragged_input():
    W, H = 320, 240
    frames, poses = synth.render_sequence(W, H, 40)
    prod, orc = product_lib(), oracle_lib()
    # zero baseline, depth range behind the camera, far depth range, and candidates on the image border
    border = [np.array([[0, 0], [3, 3], [4, 4], [5, 5], [W // 2 - 1, 5], [W - 6, H - 6], [W - 5, H - 5]], np.int32) // (1 << l) for l in range(4)]
    for args in [dict(i_src=5, i_tgt=5), dict(i_src=0, i_tgt=30, depth=(-3.0, 0.1), wiggle=-5.0), dict(i_src=0, i_tgt=30, depth=(50.0, 1.0)),
                 dict(i_src=0, i_tgt=20, cands=border)]:
        cands, o = _case(orc, W, H, frames, poses, **args)
        args2 = dict(args); args2["cands"] = cands
        _, p = _case(prod, W, H, frames, poses, **args2)
        for l in range(4):
            assert np.array_equal(o[l][0], p[l][0]) and np.array_equal(o[l][1], p[l][1])
            np.testing.assert_allclose(p[l][2], o[l][2], rtol=0, atol=1e-9)
    # empty candidate list
    t = Tracker(prod, W, H, 1)
    kf = t.add_keyframe(frames[0])
    t.make_keyframes([frames[3]])
    f, b, s = t.epipolar_search(0, 1, kf, poses[0], 1.0, 0.3, poses[3], 0.1, np.zeros((0, 2), np.int32))
    assert len(f) == 0


def test_epipolar_search_follows_the_reference():
    ref = ref_lib()
    if ref is None or not ref.has("tracker_epipolar_search"):
        pytest.skip("oracle/_ref/libref_ptam.so not present")
    W, H = 320, 240
    frames, poses = synth.render_sequence(W, H, 40)
    cands, r = _case(ref, W, H, frames, poses, 0, 30)
    _, p = _case(product_lib(), W, H, frames, poses, 0, 30, cands=cands)
    for l in range(4):
        assert np.array_equal(r[l][0], p[l][0])
        np.testing.assert_allclose(p[l][2], r[l][2], rtol=0, atol=1e-6)

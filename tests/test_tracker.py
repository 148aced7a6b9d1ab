"""Track bookkeeping and wire format against vectors the reference's own PointTracker.update /
get_tracks (src/demo.py:358-441) and to_ros_msg (src/yolopoint_ros.py:109-145) produced (oracle/make_golden.py tracker).  CPU only:
the match of every frame is supplied (as FramePipeline does from its device-side match); the GPU variant lives in
tests/test_gpu_postproc.py."""
import numpy as np
import pytest

from yolopoint_b200.tracker import PointTracker, keypoints_from_wire, keypoints_to_wire, objects_to_wire


def _frames(g):
    none = set(int(i) for i in g["none_frames"])
    f = 0
    while f"tracks{f}" in g.files:
        yield f, (None, None, None) if f in none else (g[f"pts{f}"], g[f"desc{f}"], g[f"matches{f}"])
        f += 1


def test_update_sequence_matches_reference(golden, capsys):
    g = golden("tracker.npz")
    trk = PointTracker(max_length=4, nn_thresh=0.7)
    n = 0
    for f, (p, d, m) in _frames(g):
        trk.update(p, d, matches=m)
        np.testing.assert_array_equal(trk.tracks, g[f"tracks{f}"])            # bit-identical incl. the running mean scores
        assert trk.track_count == int(g[f"count{f}"])
        np.testing.assert_array_equal(trk.get_tracks(2), g[f"long{f}"])
        np.testing.assert_array_equal(trk.get_offsets(), g[f"offsets{f}"])
        n += 1
    assert n == 8
    assert "no points were added" in capsys.readouterr().out                     # the dropped frame warns like the reference
    assert (g["tracks2"][:, 2:] != -1).sum(1).max() == 3                        # the sequence does exercise the running mean


def test_track_points_and_errors(golden):
    g = golden("tracker.npz")
    trk = PointTracker(4, 0.7)
    for f, (p, d, m) in _frames(g):
        if f > 2:
            break
        trk.update(p, d, matches=m)
    tracks = trk.get_tracks(3)
    assert tracks.shape[0] > 0
    for tr, xy in zip(tracks, trk.track_points(tracks)):
        assert xy.shape == (3, 2)
        np.testing.assert_array_equal(xy[-1], g["pts2"][:2, int(tr[-1] - trk.get_offsets()[-1])])
    with pytest.raises(ValueError):
        trk.get_tracks(0)
    with pytest.raises(ValueError):
        PointTracker(1, 0.7)
    with pytest.raises(AssertionError):
        trk.update(np.zeros((3, 4)), np.zeros((32, 5)), matches=np.zeros((3, 0)))


def test_wire_format_matches_reference(golden):
    g = golden("tracker.npz")
    w = keypoints_to_wire(g["w_pts"], g["w_desc"])
    for k in ("x", "y", "score", "desc_flat"):
        assert w[k].dtype == g[f"w_{k}"].dtype, k
        np.testing.assert_array_equal(w[k], g[f"w_{k}"])
    assert int(w["desc_len"]) == int(g["w_desc_len"]) == 64
    pts, desc = keypoints_from_wire(w)
    np.testing.assert_array_equal(pts[:2], np.floor(g["w_pts"][:2]))
    np.testing.assert_array_equal(desc, g["w_desc"].astype(float))
    objs = objects_to_wire(g["w_det"], list(g["w_names"]))
    assert [o["class_index"] for o in objs] == list(g["w_obj_index"]) and [o["class_name"] for o in objs] == list(g["w_obj_name"])
    np.testing.assert_array_equal([[o["bounding_box_min_x"], o["bounding_box_min_y"], o["bounding_box_max_x"], o["bounding_box_max_y"]] for o in objs],
                                  g["w_obj_box"])
    np.testing.assert_array_equal(np.array([o["class_probabilities"][0] for o in objs]), g["w_obj_prob"])
    assert all(o["class_count"] == 3 and o["is_instance"] for o in objs)


def test_wire_descriptor_width_256_wraps_like_the_uint8_field():
    pts = np.array([[5., 6.], [7., 8.], [.5, .25]])
    desc = np.arange(512, dtype=np.float32).reshape(256, 2)
    w = keypoints_to_wire(pts, desc)
    assert w["desc_len"].dtype == np.uint8 and int(w["desc_len"]) == 0 and w["desc_flat"].shape == (512,)
    p2, d2 = keypoints_from_wire(w)
    np.testing.assert_array_equal(d2, desc.astype(float))
    np.testing.assert_array_equal(p2, pts)
    e = keypoints_to_wire(np.zeros((3, 0)), np.zeros((64, 0)))
    p3, d3 = keypoints_from_wire(e)
    assert p3.shape == (3, 0) and d3.shape == (64, 0)

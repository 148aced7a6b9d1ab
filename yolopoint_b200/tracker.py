"""Track bookkeeping and wire format either side of the descriptor match.

``PointTracker`` mirrors the reference tracker's interface and state (src/demo.py:262-288 constructor, :343-356 ``get_offsets``,
:358-422 ``update``, :424-441 ``get_tracks``): same attributes (``maxl, nn_thresh, all_pts, last_desc, tracks, track_count,
max_score``), same ``tracks`` matrix -- one row per track ``[track id, running mean match score, point id at frame t-maxl+1, ...,
point id at frame t]``, point ids counted over the concatenation of the last ``maxl`` frames, ``-1`` = no observation -- so code
that reads ``tracker.tracks`` / ``get_tracks()`` keeps working.  What differs is how the work is done:

  * the two-way match is the CUDA kernel (``api.nn_match_two_way``), or -- when the frame came through ``FramePipeline`` -- the
    matches the pipeline already produced on the device are passed in (``update(pts, desc, matches=...)``) and nothing is recomputed;
  * the per-match Python loop of the reference (one ``np.argwhere`` over all tracks per match, O(tracks x matches)) is a single
    vectorised pass: a mutual match pairs every point at most once, so the track rows it touches are distinct and the updates
    commute; the arithmetic of the running score is the reference's, element for element (float64), hence bit-identical tracks.

``keypoints_to_wire`` / ``objects_to_wire`` restate the flattening ``to_ros_msg`` does for ``KeypointArray.msg`` /
``ObjectInstance2D`` (src/yolopoint_ros.py:109-145) without ROS types.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

MAX_SCORE = 9999


class PointTracker:
    def __init__(self, max_length=4, nn_thresh=0.7):
        if max_length < 2:
            raise ValueError("max_length must be greater than or equal to 2.")     # src/demo.py:264-265
        self.maxl = max_length
        self.nn_thresh = nn_thresh
        self.all_pts = [np.zeros((2, 0)) for _ in range(self.maxl)]
        self.last_desc = None
        self.tracks = np.zeros((0, self.maxl + 2))
        self.track_count = 0
        self.max_score = MAX_SCORE
        # state of the evaluation variant of the tracker (src/models/model_wrap.py:410-432, used by src/export_descriptor.py:72-132)
        self.matches = None
        self.last_pts = None
        self.mscores = None

    @staticmethod
    def nn_match_two_way(desc1, desc2, nn_thresh):
        from .api import nn_match_two_way
        return nn_match_two_way(desc1, desc2, nn_thresh)

    def get_offsets(self) -> np.ndarray:
        """Start of every stored frame in the concatenated point numbering (the newest frame's size is not needed)."""
        sizes = [0] + [p.shape[1] for p in self.all_pts[:-1]]
        return np.cumsum(np.array(sizes))

    def get_matches(self):
        """src/models/model_wrap.py:494: after ``update``: the [3,L] match rows of the first frame pair, afterwards the [4,L]
        coordinates (x1, y1, x2, y2) of the matched points in the previous / current frame."""
        return self.matches

    def get_mscores(self):
        """The raw [3,L] match rows (index in the previous frame, index in this frame, score) of the last update."""
        return self.mscores

    def clear_desc(self):
        """src/models/model_wrap.py:500: forget the previous descriptors (the next update starts new tracks only)."""
        self.last_desc = None

    def update(self, pts, desc, matches: Optional[np.ndarray] = None):
        """Add the observations of a new frame: pts [3,N] (x, y, conf), desc [D,N].  ``matches`` ([3,L] rows (index in the previous
        frame, index in this frame, score)) may be supplied when the match against the previous frame's descriptors has already been
        computed (FramePipeline does it on the device); otherwise it is computed here with the CUDA match."""
        if pts is None or desc is None:
            print("PointTracker: Warning, no points were added to tracker.")
            return
        assert pts.shape[1] == desc.shape[1]
        if self.last_desc is None:
            self.last_desc = np.zeros((desc.shape[0], 0))
        # slide the window: the oldest frame leaves, all point ids move down by its size
        gone = self.all_pts[0].shape[1]
        self.all_pts = self.all_pts[1:] + [pts]
        ids = self.tracks[:, 3:] - gone                     # column 2 (oldest frame) is dropped
        ids[ids < -1] = -1
        offsets = self.get_offsets()
        T = self.tracks.shape[0]
        tracks = np.concatenate((self.tracks[:, :2], ids, np.full((T, 1), -1.0)), axis=1)
        if matches is None:
            matches = self.nn_match_two_way(self.last_desc, desc, self.nn_thresh)
        matches = np.asarray(matches, dtype=np.float64).reshape(3, -1)
        self.matches = matches
        if self.last_desc.shape[1] and desc.shape[1]:     # model_wrap.py:475 keeps the raw [3,L] rows of the last non-trivial match
            self.mscores = matches
        if self.last_pts is not None:                       # src/models/model_wrap.py:536-541
            self.matches = np.concatenate((self.last_pts[:, matches[0].astype(int)], pts[:2, matches[1].astype(int)]), axis=0)
        N = pts.shape[1]
        matched = np.zeros(N, dtype=bool)
        if matches.shape[1] and T:
            prev_local = tracks[:, -2] - offsets[-2]        # index in the previous frame of every track's head (< 0: headless)
            n_prev = self.all_pts[-2].shape[1]
            row_of = np.full(max(n_prev, 1), -1, dtype=np.int64)
            heads = np.nonzero(tracks[:, -2] >= 0)[0]
            row_of[prev_local[heads].astype(np.int64)] = heads
            i1, i2, sc = matches[0].astype(np.int64), matches[1].astype(np.int64), matches[2]
            ok = i1 < n_prev
            rows = np.where(ok, row_of[np.minimum(i1, max(n_prev, 1) - 1)], -1)
            hit = rows >= 0
            rows, i2h, sch = rows[hit], i2[hit], sc[hit]
            matched[i2h] = True
            tracks[rows, -1] = i2h + offsets[-1]
            fresh = tracks[rows, 1] == self.max_score
            length = (tracks[rows, 2:] != -1).sum(axis=1) - 1.0
            with np.errstate(divide="ignore", invalid="ignore"):
                frac = 1.0 / length.astype(float)
                mean = (1.0 - frac) * tracks[rows, 1] + frac * sch
            tracks[rows, 1] = np.where(fresh, sch, mean)
        # every unmatched point starts a track
        new_ids = (np.arange(N) + offsets[-1])[~matched]
        born = np.full((new_ids.shape[0], self.maxl + 2), -1.0)
        born[:, -1] = new_ids
        born[:, 0] = self.track_count + np.arange(new_ids.shape[0])
        born[:, 1] = self.max_score
        tracks = np.vstack((tracks, born))
        self.track_count += new_ids.shape[0]
        self.tracks = tracks[np.any(tracks[:, 2:] >= 0, axis=1), :]
        self.last_desc = desc.copy()
        self.last_pts = pts[:2, :].copy()

    def get_tracks(self, min_length):
        """Tracks with at least ``min_length`` observations and one in the newest frame ([M, 2 + maxl] copy)."""
        if min_length < 1:
            raise ValueError("'min_length' too small.")
        long_enough = np.sum(self.tracks[:, 2:] != -1, axis=1) >= min_length
        has_head = self.tracks[:, -1] != -1
        return self.tracks[np.logical_and(long_enough, has_head), :].copy()

    def track_points(self, tracks: np.ndarray) -> List[np.ndarray]:
        """For every track the [k, 2] pixel coordinates of its observations, oldest first (what ``draw_tracks`` of the reference
        walks over, src/demo.py:443-480, without the drawing)."""
        offsets = self.get_offsets()
        out = []
        for tr in tracks:
            pts = []
            for i in range(self.maxl):
                if tr[i + 2] != -1:
                    pts.append(self.all_pts[i][:2, int(tr[i + 2] - offsets[i])])
            out.append(np.array(pts).reshape(-1, 2))
        return out


def keypoints_to_wire(pts: np.ndarray, desc: np.ndarray) -> dict:
    """Fields of ``KeypointArray.msg`` as ``to_ros_msg`` fills them (src/yolopoint_ros.py:110-117): note that the message's ``x`` is
    the ROW (``pts[1]``) and ``y`` the column (``pts[0]``), uint16 truncation of the float coordinates, float32 scores, the
    descriptor matrix [D, N] flattened row-major (all values of dimension 0 first)."""
    return {
        "x": pts[1, :].astype(np.uint16),
        "y": pts[0, :].astype(np.uint16),
        "score": pts[2, :].astype(np.float32),
        # the message field is uint8: D = 256 (YOLOPoint-L) does not fit and wraps to 0 (what numpy < 2 did for the reference's
        # np.array(256, dtype=np.uint8); numpy >= 2 raises there) -- consumers recover D from len(desc_flat) / len(x)
        "desc_len": np.array(desc.shape[0] & 0xFF, dtype=np.uint8),
        "desc_flat": desc.flatten().astype(float),
    }


def keypoints_from_wire(msg: dict):
    """Inverse of ``keypoints_to_wire`` for a consumer: (pts [3,N] float64, desc [D,N])."""
    n = len(msg["x"])
    pts = np.stack((np.asarray(msg["y"], np.float64), np.asarray(msg["x"], np.float64), np.asarray(msg["score"], np.float64)))
    flat = np.asarray(msg["desc_flat"])
    d = flat.size // n if n else int(msg["desc_len"])
    return pts, flat.reshape(d, n)


def objects_to_wire(det, names: Sequence[str]) -> List[dict]:
    """One dict per box with the fields ``to_ros_msg`` sets on ``ObjectInstance2D`` (src/yolopoint_ros.py:127-143): boxes in
    REVERSED order (lowest confidence first), integer-truncated corners."""
    out = []
    det = np.asarray(det.cpu() if hasattr(det, "cpu") else det)
    for row in det[::-1]:
        x0, y0, x1, y1, conf, cls = (float(v) for v in row)
        c = int(cls)
        out.append({"class_name": names[c], "class_index": c, "class_count": len(names), "class_probabilities": [float(conf)], "is_instance": True,
                    "bounding_box_min_x": int(x0), "bounding_box_min_y": int(y0), "bounding_box_max_x": int(x1), "bounding_box_max_y": int(y1)})
    return out

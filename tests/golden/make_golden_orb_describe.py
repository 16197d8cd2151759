"""Descriptor-stage fixture (build container only: needs /root/reference).

    python tests/golden/make_golden_orb_describe.py

``describe_sequence_orb.npz``: the UNMODIFIED reference's ``OrbFeatureDetector.detect_and_compute``
(`/root/reference/feature_detectors.py:18-26`) on three seeded synthetic images
(``slam_experiments_b200.synth.textured_image``: 640 x 480 gray n_features=2000, 752 x 480 BGR n_features=500,
517 x 333 gray n_features=1000): keypoint position / angle / octave and the 32-byte descriptors cv2 produced.  The
images are regenerated from their seeds by the tests, so only keypoints and descriptors are stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
from feature_detectors import OrbFeatureDetector  # noqa: E402  (the reference's module, unmodified)

# the package itself needs a GPU to import its detector; the image generator is plain numpy
import importlib.util  # noqa: E402
spec = importlib.util.spec_from_file_location("hm_synth", os.path.join(ROOT, "slam_experiments_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

CASES = (("gray640", 480, 640, 7, 1, 2000), ("bgr752", 480, 752, 11, 3, 500), ("gray517", 333, 517, 13, 1, 1000))


def main():
    out = {"cases": np.array([c[0] for c in CASES]), "params": np.array([c[1:] for c in CASES], np.int64)}
    for name, h, w, seed, ch, nf in CASES:
        img = synth.textured_image(h, w, seed, ch)
        kps, desc = OrbFeatureDetector(n_features=nf).detect_and_compute(img, None)
        out[f"{name}_xy"] = np.array([k.pt for k in kps], np.float32)
        out[f"{name}_angle"] = np.array([k.angle for k in kps], np.float32)
        out[f"{name}_octave"] = np.array([k.octave for k in kps], np.int32)
        out[f"{name}_desc"] = np.ascontiguousarray(desc)
        print(name, len(kps), "keypoints, octaves", np.bincount(out[f"{name}_octave"]).tolist())
    np.savez_compressed(os.path.join(HERE, "describe_sequence_orb.npz"), **out)


if __name__ == "__main__":
    main()

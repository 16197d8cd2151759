"""Descriptor stage timing (8f rank 3): cv2.ORB.detect / compute on the host cores next to hm_frame_put_orb.
One JSON line.  python tools/time_orb.py [n_features]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np
import torch
import slam_experiments_b200 as sx
from slam_experiments_b200 import synth
from slam_experiments_b200.feature_detectors import keypoint_arrays

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
img = synth.textured_image(480, 752, 11)
orb = cv2.ORB.create(nfeatures=nf)
kps = orb.detect(img, None)
_, ref = orb.compute(img, kps)
det = sx.OrbFeatureDetector(n_features=nf)
store = sx.FrameDescriptorStore()
got = store.put_image("f", img, kps, want_descriptors=True)
reps = 50


def timed(fn):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps * 1e3


def put_sync():
    store.put_image("f", img, kps)
    torch.cuda.synchronize()


xy, ang, octv = keypoint_arrays(kps)
ctx = store._ctx


def put_arrays_sync():           # without the Python loop over cv2.KeyPoint objects
    ctx.frame_put_orb(0, img, xy, ang, octv, 8, None, False)
    torch.cuda.synchronize()


print(json.dumps({
    "image": "752x480 synthetic texture", "keypoints": len(kps), "bit_identical_to_cv2": bool(np.array_equal(got, ref)),
    "cv2_detect_ms": timed(lambda: orb.detect(img, None)), "cv2_compute_ms": timed(lambda: orb.compute(img, kps)),
    "cv2_detect_and_compute_ms": timed(lambda: orb.detectAndCompute(img, None)),
    "device_put_image_ms": timed(put_sync), "device_put_arrays_ms": timed(put_arrays_sync),
    "host_threads": cv2.getNumThreads()}))

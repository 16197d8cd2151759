"""B200-native brute-force Hamming matcher: drop-in for the feature-matcher plugin of
ViV99/slam-experiments (`/root/reference/feature_matchers.py`).

The package holds only what the hot path needs: ``csrc/`` (hand-written sm_100a CUDA
kernels and the C ABI of ``include/hm_matcher.h``), the ctypes binding, and the host-side
mirror of the reference's plugin interface.
"""
from .feature_detectors import FeatureDetector, OrbFeatureDetector
from .feature_matchers import (BFMatcher, BruteForceFeatureMatcher, DMatch, FeatureMatcher,
                               MatcherError, NORM_HAMMING)
from .frontend_glue import (FrameDescriptorStore, KeyframeWindow, get_descriptors, get_featured_detection_mask, keypoint_array,
                            match_features, matched_point_arrays, propagate_map_points)
from .keyframe_db import ShardedKeyframeDatabase, shard_ranges
from ._native import NativeError

__all__ = ["FeatureDetector", "OrbFeatureDetector", "FrameDescriptorStore", "KeyframeWindow", "get_descriptors", "get_featured_detection_mask", "keypoint_array", "match_features", "matched_point_arrays",
           "propagate_map_points", "BFMatcher", "BruteForceFeatureMatcher", "DMatch", "FeatureMatcher", "MatcherError",
           "NORM_HAMMING", "NativeError", "ShardedKeyframeDatabase", "shard_ranges"]

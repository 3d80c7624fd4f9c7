"""monocular_slam_b200 -- B200-native ORB front-end (extraction + Hamming matching) for the Monocular_SLAM pipeline.

Only the hot path named by BASELINE.json is here: csrc/ holds the sm_100a CUDA kernels and the C ABI
(include/orbx.h -> liborbx.so); orb.py mirrors the reference's call surface on top of it; sharded.py shards the
path over the GPUs of one box; synthetic.py generates the benchmark inputs.  Nothing in this package runs the
algorithm on the CPU.
"""
from ._lib import (CAMERAS_DTYPE, DMATCH_DTYPE, FAST_SCORE, HARRIS_SCORE, KEYPOINT_DTYPE, LIB_PATH, TOP2_DTYPE, OrbxError, Params, build)  # noqa: F401
from .orb import (NORM_HAMMING, ORB, BFMatcher, DataManager, FeatureExtractor, Features, Frame, FundamentalFilter,  # noqa: F401
                  ORB_create, OrbDescriptorExtractor, OrbFeatureDetector, Triangulator, match_features, popc_peak)
from .bow import Vocabulary  # noqa: F401
from .ingest import JpegDecoder, JpegIngest  # noqa: F401

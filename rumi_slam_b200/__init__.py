"""rumi_slam_b200 -- B200-native (sm_100a) ORB feature front-end of RUMI-SLAM behind the reference's own
ORBextractor / ORBmatcher interface.  Product code = csrc/ (CUDA + C ABI, include/rumi_orb.h) and the thin host
mirrors in extractor.py / matcher.py / bow.py / flow.py / sharding.py.  No CPU fallback."""
from ._lib import KP_DTYPE, LIB_PATH, RumiError  # noqa: F401
from .extractor import ORBextractor  # noqa: F401
from .matcher import FrameGrid, ORBmatcher  # noqa: F401
from .bow import ORBVocabulary  # noqa: F401
from .flow import KFDSample, SparsePyrLK  # noqa: F401

"""vins_fast_b200 -- B200-native (sm_100a) stereo visual front-end with the API of vins-fast's
estimator::FeatureTracker.  See DESIGN.md; the C ABI is include/vins_fast_b200.h."""
from ._lib import FEATURE_DTYPE, VfCamera, VfConfig, VfError, VfFeature, load  # noqa: F401
from .tracker import (Batch, FeatureTracker, default_config, feature_frame, load_camera_yaml,  # noqa: F401
                      load_config_yaml, make_config)

__version__ = "0.1.0"

"""sfx_b200 -- host side of the sm_100a batched speech feature extractor (libsfx_b200.so)."""
from .extractor import (NoCudaDeviceError, SpeechFeatureExtractor, extract_features_batch,  # noqa: F401
                        get_extractor)
from ._lib import SfxError  # noqa: F401

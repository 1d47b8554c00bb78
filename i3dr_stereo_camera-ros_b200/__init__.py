"""B200-native semi-global block matching disparity engine (drop-in for the reference's SGBM matcher path)."""
from .params import SGBMParams, CParams, Config, CONFIGS, C5_CAMERA, MODE_SGBM, MODE_HH  # noqa: F401
from . import synth  # noqa: F401
from .engine import Engine, B200SGMError, load_library, LIB_PATH, SYMBOLS, STAGES, alu_peak, reproject_from_camera, CReproject  # noqa: F401
from . import stream  # noqa: F401

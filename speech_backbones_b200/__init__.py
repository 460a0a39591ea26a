"""Import shim: the product package lives in `speech-backbones_b200/` (a name Python
cannot import directly); this module makes it importable as `speech_backbones_b200`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "speech-backbones_b200")
__path__.insert(0, _real)

from .spec import UNetConfig, estimator_param_spec, synthetic_state_dict, synthetic_inputs, synthetic_noise  # noqa: E402,F401

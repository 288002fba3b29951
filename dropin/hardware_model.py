"""sys.path shim: `import hardware_model` resolves here when <repo>/dropin precedes the reference
(noisynet.py:14, models/resnet.py:11, main.py:37, chip_mnist.py:9)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from noisynet_b200.hardware_model import *  # noqa: F401,F403,E402
from noisynet_b200.hardware_model import (AddNoise, NoisyConv2d, NoisyLinear, QuantMeasure, UniformQuantize,  # noqa: F401,E402
                                          add_noise_calculate_power, distort_tensor)

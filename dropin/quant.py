"""sys.path shim: `import quant` / `from quant import QuantMeasure` (models/mobilenet.py:9, models/resnet.py:9)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from noisynet_b200.quant import QuantMeasure, UniformQuantize  # noqa: F401,E402

"""B200-native hot path of kenm-28/Semilayer-Wise-Mixed-Precision-Quantization.

The reference is a flat directory of modules (``import functions``, ``import resnet``); this
directory mirrors that: put it on ``sys.path`` in place of the reference and the unmodified
``resnetXX_main.py`` scripts run on the sm_100a kernels.  Importing this package does exactly that.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

import slq_build  # noqa: E402
import slq_lib  # noqa: E402
import imagenet  # noqa: E402,F401
import resnet  # noqa: E402
import functions  # noqa: E402
import slq_engine  # noqa: E402

__all__ = ["slq_build", "slq_lib", "resnet", "functions", "slq_engine", "imagenet"]

"""misa_md_b200 -- B200-native EAM force hot path for MISA-MD (CUDA library + thin Python test/bench harness).

The product is misa_md_b200/libmisa_b200.so (C ABI: include/misa_b200.h). Nothing in this package falls back to
a CPU implementation: importing works everywhere, computing needs the built library and a CUDA device.
"""
import os

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
SETFL_PATH = os.path.join(DATA_DIR, "FeCuNi.synthetic.eam.alloy")

from . import synth  # noqa: E402,F401
from . import capi  # noqa: E402,F401
from .capi import Context, MisaError, load, device_count  # noqa: E402,F401

"""Forwarder: the import shim for the unmodified reference lives in ``oracle/ref_shim.py`` (the CPU arm
of bench.py uses it too).  The golden generators in this directory run in the dev container only and
read the mounted checkout."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("SFB_REFERENCE_ROOT", "/root/reference")

from oracle.ref_shim import *  # noqa: E402,F401,F403
from oracle.ref_shim import import_reference, install, reference_root  # noqa: E402,F401

REFERENCE_ROOT = "/root/reference"

"""Shim: ``from supres import DSen2_20, DSen2_60`` (testing/supres.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dsen2_b200.supres import *  # noqa: F401,F403,E402
from dsen2_b200.supres import DSen2_20, DSen2_60, MDL_PATH, SCALE  # noqa: F401,E402

"""Shim: ``from utils.patches import get_test_patches, get_test_patches60, recompose_images`` (utils/patches.py)."""
from dsen2_b200.patches import (get_test_patches, get_test_patches60, interp_patches,  # noqa: F401
                                recompose_images)

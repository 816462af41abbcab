"""Shim: ``from utils.patches import get_test_patches, get_test_patches60, recompose_images`` (utils/patches.py)."""
from dsen2_b200.patches import (OpenDataFiles, OpenDataFilesTest, get_test_patches,  # noqa: F401
                                get_test_patches60, interp_patches, recompose_images, splitTrainVal)

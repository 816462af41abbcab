"""Shim: ``from utils.patches import ...`` (utils/patches.py) -- test-time tiling, training-data helpers, loaders."""
from dsen2_b200.patches import (OpenDataFiles, OpenDataFilesTest, downPixelAggr, get_test_patches,  # noqa: F401
                                get_test_patches60, interp_patches, recompose_images, save_random_patches,
                                save_random_patches60, save_test_patches, save_test_patches60, splitTrainVal)

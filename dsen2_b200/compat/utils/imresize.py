"""Shim: ``from utils.imresize import imresize`` (utils/imresize.py)."""
from dsen2_b200.imresize import imresize  # noqa: F401

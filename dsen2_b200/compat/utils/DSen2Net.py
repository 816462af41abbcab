"""Shim: ``from utils.DSen2Net import s2model`` (utils/DSen2Net.py)."""
from dsen2_b200.DSen2Net import S2Model, resBlock, s2model  # noqa: F401

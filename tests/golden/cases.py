"""Seeded synthetic inputs shared by make_golden.py (reference side) and the tests (oracle / CUDA side)."""
import numpy as np

CASES20 = {'a': (300, 412), 'b': (560, 560), 'c': (112, 112), 'd': (226, 118)}
CASES60 = {'e': (192, 300), 'f': (336, 168), 'g': (600, 348)}


def synth(tag):
    h, w = {**CASES20, **CASES60}[tag]
    rng = np.random.RandomState(1000 + ord(tag))
    d10 = rng.randint(0, 9000, size=(h, w, 4)).astype(np.float32)
    d20 = rng.randint(0, 9000, size=(h // 2, w // 2, 6)).astype(np.float32)
    d60 = rng.randint(0, 9000, size=(h // 6, w // 6, 2)).astype(np.float32) if tag in CASES60 else None
    return d10, d20, d60


def synth_pred(tag, n, c, p):
    rng = np.random.RandomState(2000 + ord(tag))
    return rng.rand(n, c, p, p).astype(np.float32)

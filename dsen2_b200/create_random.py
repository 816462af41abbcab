"""Validation split -- mirror of ``training/create_random.py``: ``val_index.npy`` marks a random 10 % of the training patches
as validation samples (read back by ``patches.splitTrainVal``, ``utils/patches.py:274-285``).  It must be re-created whenever
the number of training patches changes.  The reference hard-codes ``size = 45 * 8000`` and ``path = '../data/train/'``
(``create_random.py:11,21``) and uses ``np.bool`` / ``np.int``, which numpy >= 1.24 no longer has; here both are options.

    python -m dsen2_b200.create_random [--tiles 45] [--patches_per_tile 8000] [--ratio 0.1] [--path ../data/train/]
"""
import argparse
from random import randrange

import numpy as np


def make_val_index(size, ratio=0.1):
    """``create_random.py:13-19``: draw positions with ``randrange`` until ``int(size * ratio)`` distinct ones are set
    (the same sequence of draws as the reference for a given ``random.seed``).  -> (boolean index, number of draws)."""
    nb = int(size * ratio)
    index = np.zeros(size, dtype=bool)
    draws = 0
    count = 0
    while count < nb:
        x = randrange(0, size)
        if not index[x]:
            index[x] = True
            count += 1
        draws += 1
    return index, draws


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__.split('\n')[0], formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument('--tiles', type=int, default=45, help='number of Sentinel-2 tiles')
    p.add_argument('--patches_per_tile', type=int, default=8000, help='8000 for the 20 m network, 500 for the 60 m one')
    p.add_argument('--ratio', type=float, default=0.1)
    p.add_argument('--path', default='../data/train/', help="'../data/train60/' for the 60 m network")
    args = p.parse_args(argv)
    size = args.tiles * args.patches_per_tile
    index, draws = make_val_index(size, args.ratio)
    np.save(args.path + 'val_index.npy', index)
    print('Full no of samples: {}'.format(size))
    print('Validation samples: {}'.format(int(np.sum(index))))
    print("Number of iterations: {}".format(draws))
    return 0


if __name__ == '__main__':
    raise SystemExit(main())

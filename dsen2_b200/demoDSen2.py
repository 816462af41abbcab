"""Mirror of ``testing/demoDSen2.py``: ``readh5`` (:14-28), ``RMSE`` (:31-35) and the demo driver (:38-167, without the
matplotlib figures).  The `.mat` scenes are MATLAB v7.3 (HDF5) files read by the pure-Python ``dsen2_b200.hdf5``.

    python -m dsen2_b200.demoDSen2 [--data ../data/] [--models ../models/] [--random-weights]

Scenes or weight files that are absent (the three ground-truth scenes and both weight files are missing blobs in this
checkout) are skipped with a message; ``--random-weights`` runs the pipeline with he_uniform weights instead so that
the 20 m / 60 m paths and the bicubic baseline can still be exercised end to end.
"""
import argparse
import os
import sys

import numpy as np

from .hdf5 import File

DATA_PATH = '../data/'


def readh5(fname, im60=False, imGT=False):
    """demoDSen2.py:14-28 -- arrays come back transposed to (H, W, C) like ``f[name][()].transpose()``."""
    f = File(DATA_PATH + fname)
    d10 = f['im10'][()].transpose()
    d20 = f['im20'][()].transpose()
    out = [d10, d20]
    if im60:
        out.append(f['im60'][()].transpose())
    if imGT:
        out.append(f['imGT'][()].transpose())
    return tuple(out)


def RMSE(x1, x2):
    """demoDSen2.py:31-35."""
    diff = np.asarray(x1).astype(np.float64) - np.asarray(x2).astype(np.float64)
    rms = np.sqrt(np.mean(np.power(diff, 2)))
    print('RMSE: {:.4f}'.format(rms))
    return rms


def main(argv=None):
    global DATA_PATH
    from . import supres
    from .DSen2Net import s2model
    from .imresize import imresize
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('--data', default=DATA_PATH)
    ap.add_argument('--models', default=supres.MDL_PATH)
    ap.add_argument('--random-weights', action='store_true')
    args = ap.parse_args(argv)
    DATA_PATH, supres.MDL_PATH = args.data, args.models
    models = {}

    def model_for(run_60):
        if not args.random_weights:
            return None                                     # -> supres loads MDL_PATH + s2_03x_lr_*.hdf5
        if run_60 not in models:
            shape = ((4, None, None), (6, None, None)) + (((2, None, None),) if run_60 else ())
            models[run_60] = s2model(shape, num_layers=6, feature_size=128, seed=0)
        return models[run_60]

    # (scene, 60 m path?, has ground truth?) in the order of demoDSen2.py:42-129
    scenes = [('S2B_MSIL1C_20170725_T43WFQ.mat', False, True), ('S2A_MSIL1C_20171028_T34HCH.mat', True, True),
              ('S2B_MSIL1C_20170928_T18TWL.mat', False, True), ('S2A_MSIL1C_20170527_T33UUB.mat', False, False),
              ('S2A_MSIL1C_20170527_T33UUB.mat', True, False), ('S2B_MSIL1C_20171022_T49JGM.mat', False, False),
              ('S2B_MSIL1C_20171022_T49JGM.mat', True, False)]
    done = 0
    for name, run_60, has_gt in scenes:
        if not os.path.exists(DATA_PATH + name):
            print('skipping %s: not in %s' % (name, DATA_PATH))
            continue
        try:
            data = readh5(name, im60=run_60, imGT=has_gt)
            print('%s  (%s m -> 10 m)' % (name, 60 if run_60 else 20))
            if run_60:
                sr = supres.DSen2_60(data[0], data[1], data[2], deep=False, model=model_for(True))
                low, scale = data[2], 6
            else:
                sr = supres.DSen2_20(data[0], data[1], deep=False, model=model_for(False))
                low, scale = data[1], 2
        except OSError as e:
            print('skipping %s: %s' % (name, e))
            continue
        if has_gt:
            print('DSen2:')
            RMSE(sr, data[-1])
            print('Bicubic:')
            RMSE(imresize(low, scale), data[-1])
        else:
            print('super-resolved %s, mean %.3f' % (sr.shape, float(sr.mean())))
        done += 1
    return 0 if done else 1


if __name__ == '__main__':
    sys.exit(main())

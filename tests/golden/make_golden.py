"""Generate golden vectors from the reference's OWN code (run in the build container only).

Imports ``/root/reference/utils/patches.py`` (with ``skimage.transform`` stubbed:
only ``interp_patches`` touches it) and ``/root/reference/utils/imresize.py``
unmodified, runs them on the two scenes present in the mount and on small
synthetic inputs, and writes

* ``tests/golden/scene_malmo_u16.npz``, ``scene_shark_u16.npz`` -- the two scenes present in the mount (Malmo,
  Shark Bay) as uint16 (CC BY 4.0, Copernicus / ESA; values are integer DN so the cast is lossless),
* ``tests/golden/reference_golden.npz``  -- small outputs of the reference code,
* ``tests/golden/fingerprints.json``     -- sha1 of the large outputs.

Usage:  python tests/golden/make_golden.py
"""
import hashlib
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = '/root/reference'


def load_reference():
    stub = types.ModuleType('skimage')
    stub.transform = types.ModuleType('skimage.transform')
    stub.transform.resize = lambda *a, **k: (_ for _ in ()).throw(RuntimeError('skimage absent'))
    sys.modules['skimage'] = stub
    sys.modules['skimage.transform'] = stub.transform
    sys.path.insert(0, REF)
    from utils import patches as ref_patches, imresize as ref_imresize
    return ref_patches, ref_imresize


def sha(a):
    return hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    from dsen2_b200.hdf5 import File
    rp, ri = load_reference()
    fp, small = {}, {}
    scenes = {'malmo': 'S2A_MSIL1C_20170527_T33UUB.mat', 'shark': 'S2B_MSIL1C_20171022_T49JGM.mat'}
    for name, fn in scenes.items():
        f = File(os.path.join(REF, 'data', fn))
        raw = {k: f[k][()] for k in ('im10', 'im20', 'im60')}
        d10, d20, d60 = (raw[k].transpose() for k in ('im10', 'im20', 'im60'))
        for k in raw:
            fp['%s.%s' % (name, k)] = hashlib.sha1(np.ascontiguousarray(raw[k]).tobytes()).hexdigest()[:12]
        for a in (d10, d20, d60):                  # both shipped scenes travel to the GPU box as uint16 fixtures
            assert np.array_equal(a, a.astype(np.uint16))
        np.savez_compressed(os.path.join(HERE, 'scene_%s_u16.npz' % name),
                            im10=d10.astype(np.uint16), im20=d20.astype(np.uint16), im60=d60.astype(np.uint16))
        p10, p20 = rp.get_test_patches(d10, d20, 128, 8, interp=False)
        q10, q20, q60 = rp.get_test_patches60(d10, d20, d60, 192, 12, interp=False)
        fp.update({name + '.p10': sha(p10), name + '.p20': sha(p20), name + '.q10': sha(q10),
                   name + '.q20': sha(q20), name + '.q60': sha(q60)})
        with redirect_stdout(io.StringIO()):
            r20 = rp.recompose_images(p10, 8, d10.shape)
            r60 = rp.recompose_images(q10, 12, d10.shape)
        assert np.array_equal(r20, d10) and np.array_equal(r60, d10)
        b2 = ri.imresize(d20, 2)
        b6 = ri.imresize(d60, 6)
        fp.update({name + '.bic2': sha(b2), name + '.bic6': sha(b6),
                   name + '.bic2.mean': float(b2.mean()), name + '.bic6.mean': float(b6.mean())})
        if name == 'malmo':
            small['malmo_bic2_crop'] = b2[:40, :40].copy()
            small['malmo_bic6_crop'] = b6[-40:, -40:].copy()
            small['malmo_p10_last'] = p10[24, :, :16, :16].copy()
    # synthetic cases through the reference code: ragged sizes, exact multiples, 60 m grid
    from cases import CASES20, CASES60, synth, synth_pred
    for tag in CASES20:
        d10, d20, _ = synth(tag)
        p10, p20 = rp.get_test_patches(d10, d20, 128, 8, interp=False)
        fp['syn_%s.p10' % tag], fp['syn_%s.p20' % tag] = sha(p10), sha(p20)
        fp['syn_%s.n' % tag] = int(p10.shape[0])
        pred = synth_pred(tag, p10.shape[0], 3, 128)
        with redirect_stdout(io.StringIO()):
            rec = rp.recompose_images(pred, 8, d10.shape)
        fp['syn_%s.rec' % tag] = sha(rec)
    for tag in CASES60:
        d10, d20, d60 = synth(tag)
        q10, q20, q60 = rp.get_test_patches60(d10, d20, d60, 192, 12, interp=False)
        fp['syn_%s.q10' % tag], fp['syn_%s.q20' % tag], fp['syn_%s.q60' % tag] = sha(q10), sha(q20), sha(q60)
        fp['syn_%s.n' % tag] = int(q10.shape[0])
        pred = synth_pred(tag, q10.shape[0], 2, 192)
        with redirect_stdout(io.StringIO()):
            rec = rp.recompose_images(pred, 12, d10.shape)
        fp['syn_%s.rec' % tag] = sha(rec)
    rng = np.random.RandomState(7)
    # bicubic known answers on a small random image (float64 outputs stored in full)
    img = rng.rand(9, 7, 2).astype(np.float32) * 4000
    small['bic_in'] = img
    small['bic_out2'] = ri.imresize(img, 2)
    small['bic_out6'] = ri.imresize(img, 6)
    small['bic_out_shape'] = ri.imresize(img, output_shape=(20, 11))
    np.savez_compressed(os.path.join(HERE, 'reference_golden.npz'), **small)
    with open(os.path.join(HERE, 'fingerprints.json'), 'w') as fh:
        json.dump(fp, fh, indent=1, sort_keys=True)
    print('wrote', len(small), 'arrays and', len(fp), 'fingerprints')


if __name__ == '__main__':
    main()

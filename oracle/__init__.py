"""CPU oracle for the DSen2 super-resolution hot path.  TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``dsen2_b200/``) imports this directory.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU baseline -- never as the thing shipped.

Pinning status (see DESIGN.md "Oracle"):

* ``patches_oracle`` (extract / stitch indexing): PINNED -- compared against the
  reference's own ``utils/patches.py`` executed in the build container
  (``tests/golden/make_golden.py``), fingerprints + small arrays committed.
* ``imresize_oracle`` (MATLAB bicubic): PINNED -- bit-identical to the reference's
  own ``utils/imresize.py`` on both shipped scenes and on synthetic inputs.
* ``patches_oracle.interp_patches`` (bilinear, mirror): parity unpinned by the
  reference (scikit-image is absent and not vendored); pinned instead against
  ``scipy.ndimage.zoom(order=1, mode='mirror', grid_mode=True)``, which is what
  scikit-image >= 0.19 ``resize(mode='reflect')`` dispatches to.
* ``dsen2net_oracle`` (Keras graph, fp32): PARITY UNPINNED -- Keras/TensorFlow
  and the shipped ``models/*.hdf5`` / ground-truth scenes are absent from this
  mount, so no reference artefact pins the CNN arithmetic.  It restates
  ``utils/DSen2Net.py:9-43`` with ``torch.nn.functional.conv2d`` in fp32.
"""

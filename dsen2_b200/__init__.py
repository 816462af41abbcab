"""dsen2_b200 -- B200-native DSen2 / VDSen2 super-resolution inference path.

Drop-in mirror of the reference's Python interface (ACMEAtronOmatic/DSen2):

* ``dsen2_b200.supres``    <-> ``testing/supres.py``    (``DSen2_20``, ``DSen2_60``)
* ``dsen2_b200.DSen2Net``  <-> ``utils/DSen2Net.py``    (``s2model``)
* ``dsen2_b200.patches``   <-> ``utils/patches.py``     (``get_test_patches[60]``, ``interp_patches``, ``recompose_images``)
* ``dsen2_b200.imresize``  <-> ``utils/imresize.py``    (``imresize``)

All arithmetic runs in hand-written sm_100a CUDA behind the C ABI declared in
``include/dsen2_b200.h`` (``dsen2_b200/_lib/libdsen2_b200.so``).  There is no CPU
fallback: importing the compute modules without the built library, or calling them
without a CUDA device, raises.
"""
__version__ = "0.1.0"

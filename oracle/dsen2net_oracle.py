"""fp32 CPU restatement of the DSen2 / VDSen2 Keras graph.  PARITY UNPINNED.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Follows
``/root/reference/utils/DSen2Net.py``: ``resBlock`` :9-15, ``s2model`` :18-43, with the
Keras semantics they rely on (third-party, Keras 2.x + TensorFlow 1.x, not in
``/root/reference`` and not installable here): ``Conv2D(padding='same')`` = zero
padding + cross-correlation, kernels stored HWIO ``(3,3,Cin,Cout)``,
``channels_first`` activations, ``he_uniform`` init with zero bias.  No reference
artefact available in this mount pins the arithmetic (weights + GT scenes are
missing blobs), hence "parity unpinned".
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def he_uniform_weights(in_ch, out_ch, num_layers, feature_size, seed=0):
    """[(kernel HWIO (3,3,Cin,Cout) f32, bias (Cout,) f32)] in Keras topological order."""
    rng = np.random.RandomState(seed)
    shapes = [(in_ch, feature_size)] + [(feature_size, feature_size)] * (2 * num_layers) + [(feature_size, out_ch)]
    out = []
    for cin, cout in shapes:
        lim = math.sqrt(6.0 / (9 * cin))
        out.append((rng.uniform(-lim, lim, size=(3, 3, cin, cout)).astype(np.float32),
                    np.zeros((cout,), np.float32)))
    return out


def _conv(x, kernel_hwio, bias):
    w = torch.from_numpy(np.ascontiguousarray(kernel_hwio.transpose(3, 2, 0, 1)))
    return F.conv2d(x, w, torch.from_numpy(bias), padding=1)


def forward(inputs, weights, scale=0.1):
    """inputs: list of (N,C,P,P) float32 arrays [x10, x20up(, x60up)]; returns (N,Cout,P,P) float32."""
    with torch.no_grad():
        xs = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) for a in inputs]
        x = torch.cat(xs, dim=1)                                   # DSen2Net.py:24,26
        x = torch.relu(_conv(x, *weights[0]))                      # :29
        n_res = (len(weights) - 2) // 2
        for l in range(n_res):                                     # :31-32 -> resBlock :9-15
            t = torch.relu(_conv(x, *weights[1 + 2 * l]))
            t = _conv(t, *weights[2 + 2 * l])
            x = x + t * scale
        x = _conv(x, *weights[-1])                                 # :35
        x = x + xs[-1]                                             # :38 / :41
        return x.numpy()


def predict(inputs, weights, batch_size=32):
    """``model.predict`` (Keras default batch 32, supres.py:65)."""
    n = inputs[0].shape[0]
    outs = [forward([a[i:i + batch_size] for a in inputs], weights) for i in range(0, n, batch_size)]
    return np.concatenate(outs, axis=0)


def DSen2_20(d10, d20, weights, scale=2000.0):
    """supres.py:15-30 with the oracle pieces."""
    from . import patches_oracle as po
    p10, p20 = po.get_test_patches(d10, d20, patchSize=128, border=8)
    p10 /= np.float32(scale)
    p20 /= np.float32(scale)
    pred = predict([p10, p20], weights)
    img = po.recompose_images(pred, border=8, size=d10.shape)
    return img * np.float32(scale)


def DSen2_60(d10, d20, d60, weights, scale=2000.0):
    """supres.py:33-50 with the oracle pieces."""
    from . import patches_oracle as po
    p10, p20, p60 = po.get_test_patches60(d10, d20, d60, patchSize=192, border=12)
    p10 /= np.float32(scale)
    p20 /= np.float32(scale)
    p60 /= np.float32(scale)
    pred = predict([p10, p20, p60], weights)
    img = po.recompose_images(pred, border=12, size=d10.shape)
    return img * np.float32(scale)

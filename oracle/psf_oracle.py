"""ORACLE (test infrastructure only -- nothing under torchoptics_b200/ imports this): CPU
restatement of the reference's Gaussian soft-histogram PSF.

    compute_psf    /root/reference/torchlens/ray_tracing.py:206-270   (the TensorFlow original; the
                   function is absent from ray_tracing_lite.py, its consumer sample_psfs is
                   commented out at optics_simulator_lite.py:656-677)

Statement by statement in numpy, float32 like TF.  Pinned (tests/test_psf_oracle.py) to
tests/golden/tf/psf.npz, which tests/golden/make_golden_tf.py produced by executing the reference's
own source file with a numpy stand-in for the `tensorflow` module -- TF itself cannot be installed
here, so TF's own rounding of exp() is the one thing not pinned.

Quirks of the reference kept on purpose: with `increment` given the y extent uses the number of x
bins (rt_tf:228: `y_size = increment * n_x_bins`); without it the function only runs for a single
grid (rt_tf:267 compares per-ray arrays with per-grid sizes); only the non-negative half of the x
bins is evaluated and mirrored (rays are expected with their x-mirror images, rt_tf:238-243, :257-261).
"""
import numpy as np


def compute_psf(x, y, n_bins=(21, 21), increment=None, y_target=None):
    """x, y: [n_lens, n_fields, n_channels, n_rays] -> (x_size, y_size, y_target [n_grids],
    kernels [n_grids, n_channels, n_y_bins, n_x_bins] of unit mass, accounted [n_lens, n_fields])."""
    x = np.asarray(x, np.float32)
    y = np.asarray(y, np.float32)
    nw = x.shape[-2]
    n_grids = x.shape[0] * x.shape[1]                                            # rt_tf:213
    n_x_bins, n_y_bins = n_bins
    if y_target is None:
        y_target = np.mean(y.reshape(n_grids, -1), axis=1, dtype=np.float32)     # rt_tf:218
    y_target = np.asarray(y_target, np.float32)
    y = y - y_target[:, None, None]                                              # rt_tf:221
    if increment is not None:                                                    # rt_tf:223-226
        x_incr = y_incr = np.ones(n_grids, np.float32) * np.float32(increment)
        x_size = increment * n_x_bins
        y_size = increment * n_x_bins
    else:                                                                        # rt_tf:228-235
        y_min = y.reshape(n_grids, -1).min(axis=1)
        y_max = y.reshape(n_grids, -1).max(axis=1)
        x_size = x.reshape(n_grids, -1).max(axis=1)
        y_size = 2 * np.maximum(y_max - y_target, y_target - y_min)
        x_incr = x_size / n_x_bins
        y_incr = y_size / n_y_bins
    if n_x_bins % 2 == 1:                                                        # rt_tf:239-242
        grid_x = np.arange(n_x_bins // 2 + 1, dtype=np.float32)[None, :] * x_incr[:, None]
    else:
        grid_x = (np.arange(n_x_bins // 2, dtype=np.float32) + 0.5)[None, :] * x_incr[:, None]
    grid_y = (np.arange(n_y_bins, dtype=np.float32) + 0.5 - n_y_bins / 2)[None, :] * y_incr[:, None]
    sigma_x = x_incr / 2                                                         # rt_tf:246-247
    sigma_y = y_incr / 2
    dx2 = (x.reshape(n_grids, nw, 1, 1, -1) - grid_x.reshape(n_grids, 1, 1, -1, 1)) ** 2
    dy2 = (y.reshape(n_grids, nw, 1, 1, -1) - grid_y.reshape(n_grids, 1, -1, 1, 1)) ** 2
    gx = np.exp(-(dx2 / sigma_x.reshape(-1, 1, 1, 1, 1) ** 2) / 2)
    gy = np.exp(-(dy2 / sigma_y.reshape(-1, 1, 1, 1, 1) ** 2) / 2)
    kernels = (gx * gy).sum(axis=-1)                                             # rt_tf:254
    if n_x_bins % 2 == 1:                                                        # rt_tf:257-260
        kernels = np.concatenate((np.flip(kernels[..., 1:], axis=-1), kernels), axis=-1)
    else:
        kernels = np.concatenate((np.flip(kernels, axis=-1), kernels), axis=-1)
    kernels = kernels / kernels.sum(axis=(-1, -2), keepdims=True)                # rt_tf:263
    accounted = (np.abs(y) < y_size / 2) & (np.abs(x) < x_size / 2)              # rt_tf:266-267
    accounted = accounted.astype(np.float32).mean(axis=(-1, -2), dtype=np.float32)
    return x_size, y_size, y_target, kernels, accounted

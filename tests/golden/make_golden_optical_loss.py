"""Golden vectors for the batched training-loss entry: the reference's
`Optical_Loss.optical_loss_unsupervised_single` (optical_loss.py:20-96) -> `RaytracedOptics.do_ray_tracing`
(optics_simulator_lite.py:456-493) -> `compute_loss_out` (:430-450), run ONE SAMPLE AT A TIME through the
reference's own source files, exactly as its per-sample loop (optical_loss.py:99-122) does.

Neither file can be imported as it stands in this image, so their missing imports are stood in for --
nothing of their own code is altered:
  * `preprocessing.process_dataframe` (absent from the reference repository): `sequence_encoder` /
    `sequence_decoder` are INFERRED from their use at optical_loss.py:14-16 (the code's digit count is the
    number of surfaces, its digit sum the number of glasses): 'G' -> 1, 'A' -> 0, read as a decimal number;
  * `matplotlib.pyplot`, `utils.w2rgb` (plotting only), `shapely` (unused): empty stand-ins;
  * `torch.Tensor.cuda` is the identity (there is no GPU here; optical_loss.py:75,83 call it
    unconditionally) and `numpy.loadtxt` returns a two-glass table for the glass catalogue path
    (optical_loss.py:90: a Colab path; the catalogue feeds only the commented-out glass penalty).

    python tests/golden/make_golden_optical_loss.py        # build container only (needs /root/reference)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, 'optical_loss')
REF = '/root/reference'


def sequence_encoder(sequence):
    return int(''.join('1' if ch == 'G' else '0' for ch in sequence))


def sequence_decoder(code):
    return ''.join('G' if ch == '1' else 'A' for ch in str(int(code)))


def import_reference():
    stubs = {}
    for name in ('matplotlib', 'matplotlib.pyplot', 'utils', 'utils.w2rgb', 'shapely', 'shapely.geometry',
                 'preprocessing', 'preprocessing.process_dataframe'):
        stubs[name] = types.ModuleType(name)
    stubs['matplotlib'].pyplot = stubs['matplotlib.pyplot']
    stubs['utils'].w2rgb = stubs['utils.w2rgb']
    stubs['utils.w2rgb'].wavelength_to_rgb = lambda *a, **k: (0, 0, 0)
    stubs['shapely'].geometry = stubs['shapely.geometry']
    stubs['shapely.geometry'].Polygon = object
    stubs['preprocessing'].process_dataframe = stubs['preprocessing.process_dataframe']
    stubs['preprocessing.process_dataframe'].sequence_encoder = sequence_encoder
    stubs['preprocessing.process_dataframe'].sequence_decoder = sequence_decoder
    sys.modules.update(stubs)
    sys.path[:0] = [REF, os.path.join(REF, 'torchlens')]
    torch.Tensor.cuda = lambda self, *a, **k: self
    real_loadtxt = np.loadtxt

    def loadtxt(path, *args, **kwargs):
        if str(path).endswith('selected_ohara_glass.csv'):
            return np.asarray([[1.5168, 64.17], [1.7847, 25.68]], dtype=np.float32)
        return real_loadtxt(path, *args, **kwargs)
    np.loadtxt = loadtxt
    import optical_loss
    import lens_modeling
    import ray_tracing_lite
    return optical_loss, lens_modeling, ray_tracing_lite


def make_samples(lens_type, n_samples, seed, lm, rtl):
    """Plausible normalised designs (EFL = 1 by construction of compute_last_curvature): the network input /
    output vectors of optical_loss.py:22-37."""
    rng = np.random.default_rng(seed)
    n_surf = len(lens_type)
    n_glass = lens_type.count('G')
    inputs, outputs = [], []
    for _ in range(n_samples):
        nd = rng.uniform(1.50, 1.80, n_glass).astype(np.float32)
        v = rng.uniform(30.0, 62.0, n_glass).astype(np.float32)
        g = lm.g_from_n_v(torch.tensor(nd), torch.tensor(v)).reshape(-1, 2).numpy()
        sign = 1.0
        c_wo_last = []
        for k in range(n_surf - 1):
            c_wo_last.append(sign * rng.uniform(0.4, 1.6))
            sign = -sign if lens_type[k] == 'A' else sign * rng.choice([1.0, -0.6])
        t = [rng.uniform(0.04, 0.12) if ch == 'G' else rng.uniform(0.02, 0.1) for ch in lens_type]
        stop_idx = 1 if lens_type[0] == 'G' and n_surf > 1 else 0
        # image distance = the paraxial back focal length of the finished lens, a little defocused
        structure = lm.Structure(stop_idx=np.asarray([stop_idx]), sequence=np.array([lens_type]), default_device='cpu')
        c_full = rtl.compute_last_curvature(structure, torch.tensor(c_wo_last, dtype=torch.float32),
                                            torch.tensor(t, dtype=torch.float32), torch.tensor(nd))
        lens = lm.Lens(structure, c_full, torch.tensor(t, dtype=torch.float32), torch.tensor(nd), torch.tensor(v))
        _, bfl = rtl.get_first_order(lens)
        t[-1] = float(bfl[0]) * rng.uniform(0.97, 1.01)
        epd = rng.uniform(0.12, 0.25)
        hfov = rng.uniform(8.0, 20.0)
        inp = [epd, hfov] + [0.0] * (2 * n_surf) + [float(sequence_encoder(lens_type)), float(stop_idx), -1.0, -1.0]
        out = list(g.reshape(-1)) + c_wo_last + t
        inputs.append(inp)
        outputs.append(out)
    return np.asarray(inputs, np.float32), np.asarray(outputs, np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    ol, lm, rtl = import_reference()
    for lens_type, n_samples, seed in (('GA', 6, 0), ('GAGA', 6, 1), ('GGA', 5, 2), ('GAGAGA', 5, 3)):
        inputs, outputs = make_samples(lens_type, n_samples, seed, lm, rtl)
        loss = ol.Optical_Loss(lens_type)
        rec = {'lens_type': np.asarray(lens_type), 'inputs': inputs, 'outputs': outputs, 'penalty_rate': np.float32(0.2)}
        per = {k: [] for k in ('loss', 'rms', 'penalty')}
        grads = []
        for i in range(n_samples):
            x = torch.tensor(inputs[i])
            y = torch.tensor(outputs[i], requires_grad=True)
            lu, rms, pen = loss.optical_loss_unsupervised_single(x, y, 0.2, device='cpu')
            per['loss'].append(float(lu))
            per['rms'].append(float(rms))
            per['penalty'].append(float(pen))
            g, = torch.autograd.grad(lu, y, allow_unused=True)
            grads.append(np.zeros_like(outputs[i]) if g is None else g.numpy())
        for k, v in per.items():
            rec['per_sample_' + k] = np.asarray(v, np.float32)
        rec['grad_outputs'] = np.asarray(grads, np.float32)
        # the SAME reference code in float64 (default dtype switched; the yardstick for fp32 noise)
        # (compute_last_curvature hard-codes float32 buffers, rtl:741-743: for this run only it is replaced by
        # its own statements, rtl:735-766, on buffers of the input's dtype)
        def last_curvature_any_dtype(structures, c, t, nd):
            mask = structures.mask_torch
            rows = torch.arange(mask.shape[0])
            n_surf = mask.sum(dim=1)
            air_air = ~structures.mask_G_torch[rows, n_surf - 2]
            last_c_idx = n_surf - 1 - air_air.long()
            c_mask = mask.clone()
            c_mask[rows, n_surf - 1] = False
            c2d = torch.zeros(mask.shape, dtype=c.dtype).masked_scatter(c_mask, c)
            t2d = torch.zeros(mask.shape, dtype=c.dtype).masked_scatter(mask, t)
            n2d = torch.ones(mask.shape, dtype=c.dtype).masked_scatter(structures.mask_G_torch, nd)
            n2d = torch.cat((torch.ones_like(n2d[:, 0:1]), n2d), dim=1)
            selection = c_mask.clone()
            selection[rows, last_c_idx] = False
            abcd = rtl.interface_propagation_abcd(c2d, t2d, n2d)
            eye = torch.eye(2, dtype=c.dtype)[None, None, ...]
            abcd = rtl.reduce_abcd(torch.where(selection[..., None, None].expand_as(abcd), abcd, eye))
            last_n = n2d[rows, last_c_idx]
            last_c = -(1 + last_n * abcd[:, 1, 0]) / (abcd[:, 0, 0] * (last_n - 1))
            c2d = c2d.clone()
            c2d[rows, last_c_idx] = last_c
            return c2d[mask]
        torch.set_default_dtype(torch.float64)
        saved = ol.compute_last_curvature
        ol.compute_last_curvature = last_curvature_any_dtype
        try:
            per64 = {k: [] for k in ('loss', 'rms', 'penalty')}
            grads64 = []
            for i in range(n_samples):
                x = torch.tensor(inputs[i], dtype=torch.float64)
                y = torch.tensor(outputs[i], dtype=torch.float64, requires_grad=True)
                lu, rms, pen = loss.optical_loss_unsupervised_single(x, y, 0.2, device='cpu')
                assert lu.dtype == torch.float64
                for k, v in zip(('loss', 'rms', 'penalty'), (lu, rms, pen)):
                    per64[k].append(float(v.detach()))
                g, = torch.autograd.grad(lu, y, allow_unused=True)
                grads64.append(np.zeros_like(outputs[i], dtype=np.float64) if g is None else g.numpy())
        finally:
            torch.set_default_dtype(torch.float32)
            ol.compute_last_curvature = saved
        for k, v in per64.items():
            rec['f64_per_sample_' + k] = np.asarray(v, np.float64)
        rec['f64_grad_outputs'] = np.asarray(grads64, np.float64)
        mean = loss.optical_loss_unsupervised(torch.tensor(inputs), torch.tensor(outputs), 0.2, device='cpu')
        rec['batch_mean'] = np.asarray([float(v) for v in mean], np.float32)
        # the supervised loss (ol:136-176) on perturbed copies of the designs as "labels"
        labels = (outputs * np.random.default_rng(seed + 50).uniform(0.9, 1.1, outputs.shape)).astype(np.float32)
        rec['supervised_labels'] = labels
        rec['supervised_loss'] = np.float32(float(loss.optical_loss_supervised(torch.tensor(labels), torch.tensor(outputs), device='cpu')))
        np.savez_compressed(os.path.join(OUT, f'{lens_type}.npz'), **rec)
        rel = np.abs(rec['grad_outputs'] - rec['f64_grad_outputs']).max(axis=1) / np.abs(rec['f64_grad_outputs']).max(axis=1)
        print(lens_type, 'rms', per['rms'], 'penalty', per['penalty'], 'nan grads', int(np.isnan(rec['grad_outputs']).sum()),
              '\n   reference fp32 vs fp64: loss', np.abs(np.asarray(per['loss']) / np.asarray(per64['loss']) - 1).max(),
              'rms', np.abs(np.asarray(per['rms']) / np.asarray(per64['rms']) - 1).max(), 'grads', rel)


if __name__ == '__main__':
    main()

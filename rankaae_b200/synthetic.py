"""Seeded synthetic spectra in the reference's CSV schema (BASELINE.json configs: "synthetic 100k x 256",
"synthetic 1M spectra x 256 points").  The real dataset (feff_V_CT_CN_OCN_RSTD_MOOD_7000.csv) is a missing blob in
the reference (.MISSING_LARGE_BLOBS:1), so benchmarks and tests run on this generator: descriptors ~ N(0, 1) with
column 1 an integer in {4, 5, 6} (a coordination number, sc/report/analysis.py:245 assumes min CN 4); every descriptor
modulates a distinct spectral feature of a smooth positive 256-point curve (edge position, white-line height, second
peak position, oscillation amplitude and frequency) so that the rank constraint has something to learn.

Data generation only — not part of the path that is measured or checked.
"""
import numpy as np


def synthetic_dataset(n, n_aux=5, dim=256, seed=0, dtype=np.float32):
    """Returns (spec [n, dim], aux [n, n_aux])."""
    rng = np.random.default_rng(seed)
    K = int(n_aux)
    d = rng.standard_normal((n, max(K, 1)))
    if K > 1:
        d[:, 1] = rng.integers(4, 7, size=n)
    grid = np.linspace(0.0, 1.0, dim)[None, :]
    dd = d.copy()
    if K > 1:
        dd[:, 1] = dd[:, 1] - 5.0
    g = lambda k: dd[:, k % max(K, 1)][:, None]
    edge = 1.0 / (1.0 + np.exp(-(grid - 0.18 - 0.02 * g(0)) * 40.0))
    peak1 = (0.9 + 0.25 * g(1)) * np.exp(-0.5 * ((grid - 0.27) / 0.035) ** 2)
    peak2 = 0.35 * np.exp(-0.5 * ((grid - 0.5 - 0.04 * g(2)) / 0.06) ** 2)
    osc = (0.12 + 0.04 * g(3)) * np.sin(2 * np.pi * (grid - 0.3) * (3.0 + 0.3 * g(4))) * (grid > 0.3)
    spec = edge * (1.0 + osc) + peak1 * (grid > 0.1) + peak2
    spec = spec + 0.02 * rng.standard_normal(spec.shape)
    spec = np.clip(spec, 0.0, None)
    return spec.astype(dtype), d[:, :K].astype(dtype)


def synthetic_dataset_chunked(n, n_aux=5, dim=256, seed=0, dtype=np.float32, chunk=100_000):
    """Large sets (the 1 M-row data-parallel configuration) generated chunk by chunk (chunk c uses seed + c) so that the
    float64 intermediates stay small."""
    specs, auxs = [], []
    for c, lo in enumerate(range(0, n, chunk)):
        s, a = synthetic_dataset(min(chunk, n - lo), n_aux, dim, seed + c, dtype)
        specs.append(s)
        auxs.append(a)
    return np.concatenate(specs), np.concatenate(auxs)


def synthetic_dataset_torch(n, n_aux=5, dim=256, seed=0, device="cuda:0"):
    """The same generator evaluated on the device with torch's generator (float32; the draws differ from the numpy
    version): for the 1 M-row data-parallel configuration, where a host-side float64 pass would dominate the set-up."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    K = int(n_aux)
    d = torch.randn(n, max(K, 1), device=device, generator=g)
    if K > 1:
        d[:, 1] = torch.randint(4, 7, (n,), device=device, generator=g).float()
    grid = torch.linspace(0.0, 1.0, dim, device=device)[None, :]
    dd = d.clone()
    if K > 1:
        dd[:, 1] -= 5.0
    gk = lambda k: dd[:, k % max(K, 1)][:, None]
    edge = torch.sigmoid((grid - 0.18 - 0.02 * gk(0)) * 40.0)
    peak1 = (0.9 + 0.25 * gk(1)) * torch.exp(-0.5 * ((grid - 0.27) / 0.035) ** 2)
    peak2 = 0.35 * torch.exp(-0.5 * ((grid - 0.5 - 0.04 * gk(2)) / 0.06) ** 2)
    osc = (0.12 + 0.04 * gk(3)) * torch.sin(2 * torch.pi * (grid - 0.3) * (3.0 + 0.3 * gk(4))) * (grid > 0.3)
    spec = edge * (1.0 + osc) + peak1 * (grid > 0.1) + peak2
    spec = spec + 0.02 * torch.randn(spec.shape, device=device, generator=g)
    return spec.clamp_(min=0.0).contiguous(), d[:, :K].contiguous()


AUX_NAMES = ["AUX_CT", "AUX_CN", "AUX_OCN", "AUX_RSTD", "AUX_MOOD", "AUX_X5", "AUX_X6", "AUX_X7"]


def write_csv(path, spec, aux, grid=None):
    """Writes spectra + descriptors in the reference CSV schema: two index columns, AUX_* x n_aux, ENE_<energy> x dim
    (sc/clustering/dataloader.py:12-25).  Values are printed with repr(float(v)), which round-trips the double exactly."""
    n, dim = spec.shape
    n_aux = aux.shape[1]
    if grid is None:
        grid = np.linspace(5460.0, 5520.0, dim)
    cols = ["mp_id", "site"] + AUX_NAMES[:n_aux] + [f"ENE_{e:.3f}" for e in grid]
    with open(path, "w") as f:
        f.write("# synthetic spectra, reference CSV schema\n")
        f.write(",".join(cols) + "\n")
        for i in range(n):
            vals = [f"mp-{i}", "0"] + [repr(float(v)) for v in aux[i]] + [repr(float(v)) for v in spec[i]]
            f.write(",".join(vals) + "\n")

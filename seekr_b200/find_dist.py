"""Background distribution of SEEKR r values behind the reference's ``find_dist`` function.

Drop-in for ``seekr.find_dist.find_dist`` (seekr/find_dist.py:82-294).  The GPU part is the hot path the
reference spends its time in (find_dist.py:141-169): normalisation vectors of the background set, its
standardised counts, ``pearson(counts, counts)``, the strict upper triangle and a random subset of it.

  * without subsetting the symmetric GEMM fills the matrix and ``skr_triu_extract`` flattens the triangle on
    the device (row-major, the order of ``np.triu_indices(n, k=1)``);
  * with subsetting only the sampled pairs are computed at all: ``subset_size`` distinct positions of the
    triangle are drawn (uniformly, without replacement, as ``np.random.choice(..., replace=False)`` does) and
    ``skr_pearson_pairs`` evaluates just those dot products -- 100 000 x 4^k instead of n^2 x 4^k products.

Fitting the candidate distributions is scipy's work in the reference and stays scipy's work here.
"""

import os
import warnings

import numpy as np

from . import _lib, device
from . import pearson as skr_pearson
from .kmer_counts import BasicCounter

COMMON10 = ['cauchy', 'chi2', 'expon', 'exponpow', 'gamma', 'lognorm', 'norm', 'pareto', 'rayleigh', 'uniform']


def triu_pairs(n, flat):
    """Row-major position(s) in the strict upper triangle of an n x n matrix -> (i, j) with i < j."""
    flat = np.asarray(flat, dtype=np.int64)
    # rows before i hold i*(n-1) - i*(i-1)/2 entries; invert with a float estimate and fix it up exactly
    b = 2 * n - 1
    i = np.floor((b - np.sqrt(np.maximum(b * b - 8.0 * flat, 0.0))) / 2).astype(np.int64)
    i = np.clip(i, 0, max(n - 2, 0))
    off = lambda r: r * (n - 1) - r * (r - 1) // 2  # noqa: E731
    for _ in range(3):
        i = np.where(off(i) > flat, i - 1, i)
        i = np.where(off(i + 1) <= flat, i + 1, i)
    j = flat - off(i) + i + 1
    return i, j


def triu_flat_device(sim):
    """Strict upper triangle of a square device matrix, flattened row-major, as a device tensor."""
    lib = _lib.load()
    torch = device.require_cuda()
    n = int(sim.shape[0])
    out = device.empty((int(lib.skr_triu_count(n)),), sim.dtype)
    _lib.check(lib.skr_triu_extract(device.ptr(sim), int(sim.dtype == torch.float64), n, sim.stride(0), device.ptr(out),
                                    device.stream_ptr(None)))
    return out


def pearson_pairs(prepared_a, prepared_b, i, j):
    """r of the listed (i, j) pairs from prepared planes (device); returns a host float32 array."""
    lib = _lib.load()
    torch = device.require_cuda()
    di = device.to_device(np.ascontiguousarray(i, dtype=np.int64))
    dj = device.to_device(np.ascontiguousarray(j, dtype=np.int64))
    out = device.empty((int(di.numel()),), torch.float32)
    _lib.check(lib.skr_pearson_pairs(device.ptr(prepared_a.hi), device.ptr(prepared_a.lo), device.ptr(prepared_a.scale),
                                     device.ptr(prepared_b.hi), device.ptr(prepared_b.lo), device.ptr(prepared_b.scale),
                                     prepared_a.K, device.ptr(di), device.ptr(dj), int(di.numel()), 1.0 / prepared_a.K,
                                     device.ptr(out), device.stream_ptr(None)))
    return device.to_host(out, pinned=False)


def background_r(counts, subsetting=True, subset_size=100000, rng=None, return_pairs=False):
    """find_dist.py:160-169 on the device: the (sub-sampled) upper triangle of pearson(counts, counts)."""
    device.require_cuda()
    prepared = skr_pearson.prepare(counts)
    n = prepared.rows
    total = n * (n - 1) // 2
    if subsetting and total > subset_size:
        rng = np.random.default_rng() if rng is None else rng
        flat = rng.choice(total, size=subset_size, replace=False)
        i, j = triu_pairs(n, flat)
        values = pearson_pairs(prepared, prepared, i, j)
        return (values, i, j) if return_pairs else values
    if subsetting:
        print("subset_size is larger than the actual data size, use the actual data size instead")
    sim = skr_pearson.pearson_device(prepared, prepared)
    values = device.to_host(triu_flat_device(sim), pinned=False)
    if return_pairs:
        i, j = triu_pairs(n, np.arange(total, dtype=np.int64))
        return values, i, j
    return values


def find_dist(inputseq='default', k_mer=4, log2='Log2.post', models='common10', subsetting=True, subset_size=100000,
              fit_model=True, statsmethod='ks', progress_bar=False, plotfit=None, outputname=None):
    """Same arguments and return values as seekr/find_dist.py:82-294 (plots need matplotlib)."""
    if inputseq == 'default':
        raise FileNotFoundError("the reference's bundled default background FASTA is not shipped with seekr_b200; "
                                "pass the path of a background FASTA as inputseq")
    if fit_model:
        from scipy import stats
        from scipy.stats import kstest

        if models == 'common10':
            names = list(COMMON10)
        else:
            names = [d for d in dir(stats) if isinstance(getattr(stats, d), (stats.rv_continuous, stats.rv_discrete))]
            names = [d for d in names if d[0] != '_' and d not in ('levy_stable', 'studentized_range')]
            if models != 'all':
                wanted = list(models)
                names_ok = [d for d in wanted if d in names]
                if len(names_ok) < len(wanted):
                    print("Please enter valid distribution names available in scipy.stats. refer to https://docs.scipy.org/doc/scipy/reference/stats.html#continuous-distributions")
                    print(f"Excluding invalid distributions for fitting: {[d for d in wanted if d not in names]}")
                names = names_ok
        distributions = [getattr(stats, d) for d in names]
        with_fit = [d for d in distributions if 'fit' in dir(d)]
        if len(with_fit) < len(distributions) and models not in ('all', 'common10'):
            print(f"Excluding distributions do not have a 'fit' method: {[d.name for d in distributions if d not in with_fit]}")
        distributions = with_fit

    # normalisation vectors of the background set, saved where the reference saves them (find_dist.py:141-147)
    bkg_norm_counter = BasicCounter(inputseq, log2=log2, k=k_mer, silent=True)
    bkg_norm_counter.get_norm_vectors()
    mean_path = f'bkg_mean_{k_mer}mers.npy'
    std_path = f'bkg_std_{k_mer}mers.npy'
    np.save(mean_path, bkg_norm_counter.mean)
    np.save(std_path, bkg_norm_counter.std)

    bkg_counter = BasicCounter(inputseq, mean=mean_path, std=std_path, k=k_mer, silent=True)
    bkg_counter._device_only = True  # the counts are only pearson's input (find_dist.py:152-160)
    bkg_counter.make_count_file()
    device_counts = getattr(bkg_counter, "counts_device", None)  # still on the device after get_counts()
    sim_triu = background_r(device_counts if device_counts is not None else bkg_counter.counts,
                            subsetting=subsetting, subset_size=subset_size)

    if not fit_model:
        if plotfit:
            print('No plot will be produced as fit_model is set to False, please set fit_model=True to plot the fitted distributions vs the actual data')
        if outputname:
            np.savetxt(f'{outputname}.csv', sim_triu, delimiter=",")
        return sim_triu

    if len(distributions) > 50 and len(sim_triu) > 5000000 and subsetting is False:
        print("The input sequence count and distribution number for fitting are both large, subsetting is recommended to save time")
    results = []
    iterable = distributions
    if progress_bar:
        from .my_tqdm import my_tqdm

        iterable = my_tqdm()(distributions)
    for distribution in iterable:
        with warnings.catch_warnings():
            warnings.filterwarnings('ignore')
            try:
                params = distribution.fit(sim_triu)
                continuous = isinstance(distribution, stats.rv_continuous)
                if statsmethod == 'mse':
                    if continuous:
                        synthetic = distribution.rvs(*params, size=len(sim_triu))
                    else:
                        synthetic = distribution.rvs(*params[:-2], loc=params[-2], scale=params[-1], size=len(sim_triu))
                    D = np.mean((sim_triu - synthetic) ** 2)
                elif statsmethod in ('aic', 'bic'):
                    if continuous:
                        loglik = np.sum(distribution.logpdf(sim_triu, *params))
                    else:
                        loglik = np.sum(distribution.logpmf(sim_triu, *params[:-2], loc=params[-2], scale=params[-1]))
                    D = (2 if statsmethod == 'aic' else np.log(len(sim_triu))) * len(params) - 2 * loglik
                else:
                    if statsmethod != 'ks':
                        print("Please enter a valid statsmethod: 'ks', 'mse', 'aic', or 'bic'. Use default 'ks' now.")
                    D, _ = kstest(sim_triu, distribution.name, args=params)
            except Exception as e:  # a family that cannot be fitted is left out, as in the reference
                print(f"Could not fit {distribution.name} because {e}, excluding it from the results")
                continue
            results.append((distribution.name, D, params))
    results.sort(key=lambda x: x[1])

    if plotfit:
        import matplotlib.pyplot as plt

        n = len(results)
        n_cols = min(5, n)
        n_rows = n // n_cols + (n % n_cols > 0)
        fig, axes = plt.subplots(n_rows, n_cols, figsize=(n_cols * 3, n_rows * 3))
        axes = np.ravel(axes)
        for idx, (ax, (dist_name, dval, params)) in enumerate(zip(axes, results)):
            x = np.linspace(min(sim_triu), max(sim_triu), 1000)
            ax.hist(sim_triu, bins=100, density=True, alpha=0.6, color='skyblue')
            ax.plot(x, getattr(stats, dist_name).pdf(x, *params), 'r--', linewidth=2)
            ax.set_title(f'{idx + 1}: {dist_name} (Dev={dval:.3f})')
        for extra in range(len(results), len(axes)):
            fig.delaxes(axes[extra])
        plt.tight_layout()
        plt.savefig(f'{plotfit}.pdf', dpi=300)

    if outputname:
        import pandas as pd

        pd.DataFrame(results, columns=['distribution_name', 'D_statistics', 'params']).to_csv(f'{outputname}.csv', index=False)
    return results

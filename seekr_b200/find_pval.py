"""p-values of SEEKR Pearson correlations on the GPU behind the reference's ``find_pval`` function.

Drop-in for ``seekr.find_pval.find_pval`` (seekr/find_pval.py:70-183).  The reference counts both FASTA
files, forms ``pearson(t1.counts, t2.counts)`` and then walks the m x n matrix in a Python double loop,
either ``1 - distribution.cdf(r)`` (fitres = output list of find_dist) or ``np.sum(fitres > r) / len(fitres)``
(fitres = 1-D array of background r values).  Here the r matrix never leaves the device: the counts, the
GEMM and the p-value pass (``skr_pval_empirical`` / ``skr_pval_dist``) run on the GPU and only the p-values
come back.
"""

import numpy as np

from . import _lib, device
from . import pearson as skr_pearson
from .fasta_reader import Reader
from .kmer_counts import BasicCounter

# scipy.stats families with a closed-form CDF evaluated on the device: name -> (SKR_DIST_* code, has a shape)
FAMILIES = {
    "norm": (0, False), "lognorm": (1, True), "cauchy": (2, False), "expon": (3, False),
    "rayleigh": (4, False), "uniform": (5, False), "pareto": (6, True), "exponpow": (7, True),
    "gamma": (8, True), "chi2": (9, True),
}


def is_float_type(x):
    return isinstance(x, float) or np.isscalar(x)


def check_tuple_format(tup):
    """(distribution name, deviance, parameters) as find_dist returns them (find_pval.py:57-64)."""
    if not (isinstance(tup, tuple) and len(tup) == 3):
        return False
    return isinstance(tup[0], str) and is_float_type(tup[1]) and isinstance(tup[2], tuple) and \
        all(is_float_type(x) for x in tup[2])


def check_main_list(main_list):
    return all(check_tuple_format(tup) for tup in main_list)


def _sorted_background(fitres):
    """Ascending device copy of the background (float32 stays float32, everything else binary64)."""
    torch = device.require_cuda()
    arr = np.ascontiguousarray(fitres, dtype=np.float32 if fitres.dtype == np.float32 else np.float64)
    dev = device.to_device(arr)
    return torch.sort(dev).values.contiguous()


def pval_empirical_device(r, background_sorted, out=None, stream=None):
    """p = count(background > r) / N for a device matrix r (float32 / float64); returns a device tensor."""
    torch = device.require_cuda()
    lib = _lib.load()
    if out is None:
        out = device.empty(tuple(r.shape), r.dtype)
    m, n = int(r.shape[0]), int(r.shape[1])
    _lib.check(lib.skr_pval_empirical(device.ptr(r), int(r.dtype == torch.float64), m, n, r.stride(0),
                                      device.ptr(background_sorted), int(background_sorted.dtype == torch.float64),
                                      int(background_sorted.numel()), device.ptr(out), out.stride(0),
                                      device.stream_ptr(stream)))
    return out


def pval_dist_device(r, distname, params, out=None, stream=None):
    """p = 1 - scipy.stats.<distname>(*params).cdf(r) for a device matrix r; returns a device tensor."""
    torch = device.require_cuda()
    lib = _lib.load()
    if distname not in FAMILIES:
        raise NotImplementedError(
            "distribution %r has no closed-form CDF on the device (supported: %s); evaluate "
            "1 - scipy.stats.%s(*params).cdf(r) on the host for this family" % (distname, ", ".join(sorted(FAMILIES)), distname))
    code, has_shape = FAMILIES[distname]
    params = tuple(float(x) for x in params)
    if len(params) != (3 if has_shape else 2):
        raise TypeError("%s expects %d parameters (shapes, loc, scale), got %d"
                        % (distname, 3 if has_shape else 2, len(params)))
    shape = params[0] if has_shape else 0.0
    loc, scale = params[-2], params[-1]
    if distname in ("gamma", "chi2") and shape * (0.5 if distname == "chi2" else 1.0) > 1e6:
        # series / continued fraction need ~10 sqrt(a) terms per value: beyond a = 1e6 that is minutes per 1e9 values
        raise NotImplementedError("%s with shape %g: the incomplete-gamma evaluation on the device is limited to "
                                  "a <= 1e6 (a fit this close to a normal distribution is better served by 'norm')"
                                  % (distname, shape))
    if out is None:
        out = device.empty(tuple(r.shape), r.dtype)
    m, n = int(r.shape[0]), int(r.shape[1])
    _lib.check(lib.skr_pval_dist(device.ptr(r), int(r.dtype == torch.float64), m, n, r.stride(0), code, shape, loc, scale,
                                 device.ptr(out), out.stride(0), device.stream_ptr(stream)))
    return out


_BLOCK_BYTES = 1 << 31  # r and p staging per row block


def _pvalues(counts1, counts2, transform):
    """Row blocks of pearson(counts1, counts2) -> transform(block) -> host matrix."""
    torch = device.require_cuda()
    pa = skr_pearson.prepare(counts1)
    pb = skr_pearson.prepare(counts2)
    m, n, K = pa.rows, pb.rows, pa.K
    out = device.pinned_empty((m, n), np.float32)
    if m == 0 or n == 0:
        return out
    block = max(128, min(m, (_BLOCK_BYTES // (n * 4)) // 128 * 128))
    r = device.empty((min(block, m), n), torch.float32)
    p = device.empty((min(block, m), n), torch.float32)
    for row0 in range(0, m, block):
        nrows = min(block, m - row0)
        skr_pearson.gemm_block(pa, row0, nrows, pb, r, 1.0 / K)
        transform(r[:nrows], p[:nrows])
        device.d2h(out[row0:row0 + nrows], p[:nrows])
        device.sync()
    return out


def _device_counts(counter):
    """The count matrix where get_counts() left it on the device (no second upload); the host copy otherwise."""
    dev = getattr(counter, "counts_device", None)
    if dev is not None and (counter.counts is None or tuple(dev.shape) == tuple(counter.counts.shape)):
        return dev
    return counter.counts


def find_pval(seq1file, seq2file, mean_path, std_path, k_mer, fitres, log2='Log2.post', bestfit=1, outputname=None,
              progress_bar=True):
    """Same arguments, printed diagnostics and return values as seekr/find_pval.py:70-183."""
    import pandas as pd

    meanfile = np.load(mean_path)
    stdfile = np.load(std_path)
    # the reference's test, operator precedence included (find_pval.py:76)
    if len(meanfile) != 4 ** (k_mer) | len(stdfile) != 4 ** (k_mer):
        print('k_mer size is not compatible with the normalization mean and/or std files.')
        print('Please make sure the normalization mean and std files are generated using the same kmer size as specified here in k_mer.')
        print('No p value is calculated. The output is None.')
        return None

    t1 = BasicCounter(seq1file, mean=mean_path, std=std_path, k=k_mer, log2=log2, silent=True)
    t2 = BasicCounter(seq2file, mean=mean_path, std=std_path, k=k_mer, log2=log2, silent=True)
    t1._device_only = t2._device_only = True  # the counts are only pearson's input (find_pval.py:96-100)
    t1.make_count_file()
    t2.make_count_file()

    header1 = [i[1:] for i in t1._headers()]  # Reader(seq1file).get_headers() without a second parse
    header2 = [i[1:] for i in t2._headers()]
    if len(header1) != len(set(header1)):
        print('The headers of seq1file is not unique.')
        print('Be carefule during further analysis as there are potential indexing problems.')
    if len(header2) != len(set(header2)):
        print('The headers of seq2file is not unique.')
        print('Be carefule during further analysis as there are potential indexing problems.')

    if isinstance(fitres, list):
        if not check_main_list(fitres):
            print('The format of fitres is wrong.')
            print('fitres should be a list consisting of tuples (string, number, tuple of numbers) corresponds to (distribution name, deviance, parameters)')
            print('fitres should be the output of find_dist.')
            print('No p value is calculated. The output is None.')
            return None
        distname, _, params = fitres[bestfit - 1]
        p_values = _pvalues(_device_counts(t1), _device_counts(t2), lambda r, p: pval_dist_device(r, distname, params, out=p))
    elif isinstance(fitres, np.ndarray):
        if len(fitres.shape) != 1:
            print('The dimension of fitres as a numpy array is wrong. fitres should be a 1D numpy array.')
            print('fitres should be the output of find_dist.')
            print('No p value is calculated. The output is None.')
            return None
        background = _sorted_background(fitres)
        p_values = _pvalues(_device_counts(t1), _device_counts(t2), lambda r, p: pval_empirical_device(r, background, out=p))
    else:
        print('fitres should be the output of find_dist. It should be either a list of distributions or a numpy array.')
        print('No p value is calculated. The output is None.')
        return None

    pval_df = pd.DataFrame(p_values, index=header1, columns=header2)
    if outputname:
        pval_df.to_csv(f'{outputname}.csv')
    return pval_df

"""Pearson error against binary64 for sparse count-like rows (dev tool): sparsity x K sweep."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import seekr_oracle as po  # noqa: E402
from seekr_b200.pearson import pearson  # noqa: E402


def main():
    rng = np.random.default_rng(0)
    for K in (4096, 16384, 65536):
        for lam in (0.8, 0.2, 0.05, 0.01):
            m = 192
            x = (rng.poisson(lam, size=(m, K)) * rng.uniform(0.1, 3, size=(m, 1))).astype(np.float32)
            x[:, 0] += 1e-3  # no constant rows
            z = np.log2(x + 1.0).astype(np.float32) if lam > 0.5 else x
            got = pearson(z, z)
            want = po.pearson_f64(z, z)
            err = np.abs(got - want)
            print("K=%6d lambda=%.2f  max |r - f64| = %.2e  max |diag - 1| = %.2e  mean signed diag err = %+.2e"
                  % (K, lam, err.max(), np.abs(np.diag(got) - 1).max(), float(np.mean(np.diag(got) - 1))))


if __name__ == "__main__":
    main()

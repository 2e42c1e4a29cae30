#!/usr/bin/env python
"""Golden fixtures for the consumers of the r matrix, from the UNMODIFIED reference.

Run in the build container only (imports /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_pval.py

Writes tests/golden/pval/:
  mean_k3.npy, std_k3.npy   the vectors find_dist saves for medium.fa (bkg_mean_3mers.npy / bkg_std_3mers.npy)
  triu_k3.npy               reference find_dist(medium.fa, k_mer=3, subsetting=False, fit_model=False):
                            the full upper triangle of pearson(self, self), row-major (find_dist.py:148-163)
  pval.npz                  reference find_pval(small.fa, medium.fa, ...) p-value frames:
                              emp            fitres = triu_k3 (float32 ndarray)            find_pval.py:157-159
                              emp64          fitres = triu_k3 as float64 with perturbations
                              <family>_<n>   fitres = [(family, 0.0, params)]               find_pval.py:126-128
                            plus sim (the reference's r matrix for the same pair of files) and the headers
  MANIFEST.json
"""

import hashlib
import json
import os
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import pandas as pd  # noqa: E402
import scipy  # noqa: E402

for name in ("matplotlib", "matplotlib.pyplot"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            mod = types.ModuleType(name)
            mod.__path__ = []
            mod.__getattr__ = lambda attr, _n=name: types.SimpleNamespace()
            sys.modules[name] = mod
            if "." in name:
                parent, child = name.rsplit(".", 1)
                setattr(sys.modules[parent], child, mod)

from seekr.find_dist import find_dist  # noqa: E402  (the reference)
from seekr.find_pval import find_pval  # noqa: E402
from seekr.kmer_counts import BasicCounter  # noqa: E402
from seekr.pearson import pearson  # noqa: E402

# (family, params) as scipy orders them: shapes..., loc, scale.  Chosen so that r in [-1, 1] falls inside,
# below and above the support, plus invalid parameters (NaN everywhere).
FAMILIES = [
    ("norm", (0.02, 0.11)),
    ("norm", (-0.3, 0.5)),
    ("lognorm", (0.35, -0.6, 0.55)),
    ("lognorm", (1.2, 0.05, 0.2)),
    ("cauchy", (0.01, 0.07)),
    ("expon", (-0.25, 0.2)),
    ("rayleigh", (-0.4, 0.3)),
    ("uniform", (-0.2, 0.7)),
    ("pareto", (3.5, -1.6, 1.2)),
    ("exponpow", (1.7, -0.5, 0.9)),
    ("norm", (0.0, -1.0)),
    ("pareto", (-2.0, 0.0, 1.0)),
    ("gamma", (2.5, -0.4, 0.12)),
    ("gamma", (120.0, -2.1, 0.0175)),
    ("gamma", (0.6, -0.3, 0.2)),
    ("gamma", (4200.0, -7.0, 0.00167)),
    ("chi2", (7.0, -0.5, 0.07)),
    ("chi2", (300.0, -3.2, 0.0107)),
    ("gamma", (-1.0, 0.0, 1.0)),
]


def main():
    out_dir = os.path.join(HERE, "pval")
    os.makedirs(out_dir, exist_ok=True)
    small = os.path.join(HERE, "small.fa")
    medium = os.path.join(HERE, "medium.fa")
    written = []
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)  # find_dist writes bkg_mean_3mers.npy / bkg_std_3mers.npy into the working directory
        try:
            triu = find_dist(inputseq=medium, k_mer=3, log2="Log2.post", subsetting=False, fit_model=False)
            mean = np.load("bkg_mean_3mers.npy")
            std = np.load("bkg_std_3mers.npy")
        finally:
            os.chdir(cwd)
    assert triu.dtype == np.float32 and triu.ndim == 1
    np.save(os.path.join(out_dir, "triu_k3.npy"), triu)
    np.save(os.path.join(out_dir, "mean_k3.npy"), mean)
    np.save(os.path.join(out_dir, "std_k3.npy"), std)
    written += ["triu_k3.npy", "mean_k3.npy", "std_k3.npy"]
    mean_path, std_path = os.path.join(out_dir, "mean_k3.npy"), os.path.join(out_dir, "std_k3.npy")

    t1 = BasicCounter(small, mean=mean_path, std=std_path, k=3, log2="Log2.post", silent=True)
    t2 = BasicCounter(medium, mean=mean_path, std=std_path, k=3, log2="Log2.post", silent=True)
    t1.make_count_file()
    t2.make_count_file()
    sim = pearson(t1.counts, t2.counts)

    out = {"sim": sim}
    kw = dict(seq1file=small, seq2file=medium, mean_path=mean_path, std_path=std_path, k_mer=3, log2="Log2.post",
              progress_bar=False)
    frame = find_pval(fitres=triu, **kw)
    out["emp"] = frame.to_numpy()
    out["rows"] = np.array(list(frame.index))
    out["cols"] = np.array(list(frame.columns))
    rng = np.random.default_rng(7)
    bg64 = triu.astype(np.float64) + rng.normal(0, 1e-9, triu.shape)  # values no float32 can represent
    np.save(os.path.join(out_dir, "bg64.npy"), bg64)
    written.append("bg64.npy")
    out["emp64"] = find_pval(fitres=bg64, **kw).to_numpy()
    for n, (family, params) in enumerate(FAMILIES):
        frame = find_pval(fitres=[(family, 0.0, tuple(float(x) for x in params))], bestfit=1, **kw)
        out[f"{family}_{n}"] = frame.to_numpy()
    np.savez_compressed(os.path.join(out_dir, "pval.npz"), **out)
    written.append("pval.npz")
    with open(os.path.join(out_dir, "families.json"), "w") as handle:
        json.dump([[f, list(p)] for f, p in FAMILIES], handle)
    written.append("families.json")

    manifest = {"numpy": np.__version__, "pandas": pd.__version__, "scipy": scipy.__version__,
                "reference": "CalabreseLab/seekr 2.0.2", "files": {}}
    for rel in written:
        with open(os.path.join(out_dir, rel), "rb") as handle:
            manifest["files"][rel] = hashlib.sha256(handle.read()).hexdigest()
    with open(os.path.join(out_dir, "MANIFEST.json"), "w") as handle:
        json.dump(manifest, handle, indent=1, sort_keys=True)
    print("wrote", written, "sim", sim.shape, "triu", triu.shape)


if __name__ == "__main__":
    main()

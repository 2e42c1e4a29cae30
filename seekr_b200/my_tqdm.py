"""Progress-bar selection with the reference's names (seekr/my_tqdm.py:16-32)."""

import sys

from tqdm import tqdm, trange


def _is_kernel():
    if "IPython" not in sys.modules:
        return False
    from IPython import get_ipython

    return getattr(get_ipython(), "kernel", None) is not None


def my_tqdm():
    if _is_kernel():
        from tqdm import tqdm_notebook

        return tqdm_notebook
    return tqdm


def my_trange():
    if _is_kernel():
        from tqdm import tnrange

        return tnrange
    return trange

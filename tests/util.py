import importlib

import numpy as np


def pkg():
    return importlib.import_module("libgrape-lite_b200")


def have_gpu():
    try:
        pkg().device_info()
        return True
    except Exception:
        return False


def rmat_graph(scale, seed=1, weight_mode=0, edgefactor=16):
    """(n, src, dst, w) from the product's host generator (synthetic input)."""
    src, dst, w = pkg().rmat_edges_host(scale, edgefactor, seed, weight_mode)
    return 1 << scale, src, dst, w


def gather_inner(frags_results):
    """Concatenate per-fragment inner results in fid order (contiguous partition
    => global index order)."""
    return np.concatenate(frags_results)

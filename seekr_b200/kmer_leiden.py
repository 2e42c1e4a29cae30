"""Similarity graph behind the reference's ``kmer_leiden`` (seekr/kmer_leiden.py:64-107), on the device.

The reference forms ``pearson(counts, counts)`` on the host, zeroes every r below ``pearsoncutoff`` and the
diagonal, wraps the dense n x n matrix in a DataFrame and hands networkx / igraph a Python list of lists
(kmer_leiden.py:88-104) -- at 50 000 transcripts that is a 10 GB matrix turned into 2.5e9 Python objects.
Here the r matrix never leaves the GPU: the symmetric GEMM fills it, and what travels to the host is the edge
list igraph needs (``skr_sim_edge_offsets`` + ``skr_sim_edge_fill``, row-major = ``np.nonzero`` order) or, for
callers that do want the DataFrame, the thresholded matrix (``skr_sim_threshold``).

The community detection itself (leidenalg on the igraph graph), the layout and the plots are library work in
the reference and stay with those libraries: ``leiden_inputs`` returns exactly what they are given.
"""

import numpy as np

from . import _lib, device
from . import pearson as skr_pearson
from .kmer_counts import BasicCounter


def _device_matrix(sim):
    torch = device.require_cuda()
    if isinstance(sim, torch.Tensor):
        if sim.dim() != 2 or sim.dtype not in (torch.float32, torch.float64) or sim.stride(1) != 1:
            raise ValueError("similarity matrix must be a 2-D float32 / float64 tensor with unit column stride")
        return sim, True
    arr = np.asarray(sim)
    if arr.ndim != 2:
        raise ValueError("similarity matrix must be 2-D")
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float64)
    return device.to_device(np.ascontiguousarray(arr)), False


def threshold_similarity(sim, pearsoncutoff=0, zero_diagonal=True, row0=0):
    """kmer_leiden.py:91-94: ``sim[sim < pearsoncutoff] = 0; np.fill_diagonal(sim, 0)``.

    A device tensor is modified in place and returned; a host array is left alone and a new host array returned.
    ``row0``: ``sim`` holds rows row0, row0+1, ... of the whole matrix (its diagonal is at column row0 + i)."""
    torch = device.require_cuda()
    lib = _lib.load()
    dev, on_device = _device_matrix(sim)
    m, n = int(dev.shape[0]), int(dev.shape[1])
    _lib.check(lib.skr_sim_threshold(device.ptr(dev), int(dev.dtype == torch.float64), m, n, dev.stride(0) if m else n,
                                     int(row0), float(pearsoncutoff), int(bool(zero_diagonal)), device.stream_ptr(None)))
    return dev if on_device else device.to_host(dev, pinned=False)


def similarity_edges(sim, pearsoncutoff=0, upper_only=False, return_offsets=False, with_sources=True, row0=0,
                     offsets=None):
    """Edges of the thresholded, zero-diagonal similarity matrix without forming it (kmer_leiden.py:91-104).

    Returns ``(rows, cols, weights)`` as host arrays (int32, int32, sim's dtype) in row-major order: the order of
    ``np.nonzero(adj > 0)`` and ``adj[adj > 0]``.  ``upper_only`` keeps j > i (one entry per undirected edge).
    ``return_offsets`` appends the CSR row offsets (int64, m + 1).  ``row0``: ``sim`` is a block of rows of the whole
    matrix starting at row row0 (a rank's shard, or one block of a result produced block by block); ``rows`` are
    whole-matrix indices.  ``offsets``: the (row, slice) offsets when they exist already (the Pearson GEMM counts the
    edges in its epilogue, ``similarity_matrix_and_offsets``): the counting pass over ``sim`` is skipped."""
    torch = device.require_cuda()
    lib = _lib.load()
    dev, _ = _device_matrix(sim)
    m, n = int(dev.shape[0]), int(dev.shape[1])
    is64 = int(dev.dtype == torch.float64)
    ld = dev.stride(0) if m else n
    stream = device.stream_ptr(None)
    slices = _lib.SIM_SLICES  # offsets per (row, column slice); every slices-th value is the CSR row offset
    if offsets is None:
        offsets = device.empty((m * slices + 1,), torch.int64)
        _lib.check(lib.skr_sim_edge_offsets(device.ptr(dev), is64, m, n, ld, int(row0), float(pearsoncutoff),
                                            int(bool(upper_only)), device.ptr(offsets), stream))
    total = int(offsets[m * slices].item())  # the one host round trip: sizes the edge arrays
    cols = device.empty((total,), torch.int32)
    rows = device.empty((total,), torch.int32) if with_sources else None
    weights = device.empty((total,), dev.dtype)
    if total:
        _lib.check(lib.skr_sim_edge_fill(device.ptr(dev), is64, m, n, ld, int(row0), float(pearsoncutoff), int(bool(upper_only)),
                                         device.ptr(offsets), device.ptr(rows) if with_sources else None,
                                         device.ptr(cols), device.ptr(weights), stream))
    # pageable destination: fresh pinned slabs for a 2.7 GB edge list cost more than they save (measured 1 199 ms
    # against 780 ms for 224 M edges, 49 against 46 ms for 2.9 M)
    out = (device.to_host(rows, pinned=False) if with_sources else None, device.to_host(cols, pinned=False),
           device.to_host(weights, pinned=False))
    if return_offsets:
        out = out + (device.to_host(offsets[::slices].contiguous(), pinned=False),)
    return out


_SIM_MAX_BYTES = 96 << 30  # r matrices up to this size are formed whole on the device


def similarity_matrix_and_offsets(pa, row0, nrows, pb, out, pearsoncutoff, upper_only, symmetric=False):
    """Rows [row0, row0 + nrows) of pearson(a, b) into ``out`` with the edge counts of kmer_leiden.py:91-104 taken in
    the GEMM's epilogue (skr_pearson_gemm_edges): returns the (row, slice) offsets ``similarity_edges`` would
    otherwise obtain from a pass over the matrix."""
    import ctypes

    torch = device.require_cuda()
    lib = _lib.load()
    kp = pa.hi.shape[1]
    offsets = device.empty((nrows * _lib.SIM_SLICES + 1,), torch.int64)
    _lib.check(lib.skr_pearson_gemm_edges(
        ctypes.c_void_p(pa.hi.data_ptr() + row0 * kp * 2), ctypes.c_void_p(pa.lo.data_ptr() + row0 * kp * 2),
        ctypes.c_void_p(pa.scale.data_ptr() + row0 * 4), nrows, device.ptr(pb.hi), device.ptr(pb.lo), device.ptr(pb.scale),
        pb.rows, pa.K, 1.0 / pa.K, device.ptr(out), out.stride(0), int(bool(symmetric)), int(row0), float(pearsoncutoff),
        int(bool(upper_only)), device.ptr(offsets), device.stream_ptr(None)))
    return offsets


def leiden_inputs(inputfile, mean, std, k, pearsoncutoff=0, upper_only=True, dense=False, block_bytes=None):
    """kmer_leiden.py:70-104 up to the hand-over to igraph: counts with the given vectors -> r -> graph.

    Returns a dict: ``names`` (headers without '>'), ``rows`` / ``cols`` / ``weights`` (the edge list; with
    ``upper_only=False`` the order and multiplicity of ``df.values[df.values > 0]``), ``offsets`` (CSR), and with
    ``dense=True`` also ``adjacency``, the thresholded host matrix the reference wraps in a DataFrame.
    When the r matrix does not fit on the device (or exceeds ``block_bytes``) it is produced, scanned and dropped in
    row blocks.  Returns None (after the reference's messages) when the vectors do not match 4**k (kmer_leiden.py:74-78)."""
    meanfile = np.load(mean) if isinstance(mean, str) else np.asarray(mean)
    stdfile = np.load(std) if isinstance(std, str) else np.asarray(std)
    # the reference's test, operator precedence included (kmer_leiden.py:74): a chained comparison around `|`
    if len(meanfile) != 4 ** (k) | len(stdfile) != 4 ** (k):
        print('kmer size is not compatible with the normalization mean and/or std files.')
        print('Please make sure the normalization mean and std files are generated using the same kmer size as specified here in k.')
        print('No Leiden community is calculated or plotted. The output is None.')
        return None
    device.require_cuda()
    counter = BasicCounter(inputfile, mean=mean, std=std, k=k, silent=True)
    counter._device_only = True  # t1.counts is only pearson's input here (kmer_leiden.py:79-88)
    counter.make_count_file()
    names = [h[1:] for h in counter._headers()]  # Reader(inputfile).get_headers() without a second parse
    counts = getattr(counter, "counts_device", None)  # still on the device after get_counts()
    prepared = skr_pearson.prepare(counts if counts is not None else counter.counts)
    del counts
    counter.counts_device = None
    n = prepared.rows
    torch = device.require_cuda()
    budget = int(block_bytes) if block_bytes else min(_SIM_MAX_BYTES, torch.cuda.mem_get_info()[0] // 2)
    if n * n * 4 <= budget:
        # the whole r matrix fits: symmetric GEMM (upper tiles + mirror), one extraction
        if upper_only:
            # the edge counts come out of the GEMM's epilogue: the matrix is read once (to fill the edge arrays)
            sim = device.empty((n, n), torch.float32)
            pre = similarity_matrix_and_offsets(prepared, 0, n, prepared, sim, pearsoncutoff, True, symmetric=True)
        else:
            sim, pre = skr_pearson.pearson_device(prepared, prepared), None
        rows, cols, weights, offsets = similarity_edges(sim, pearsoncutoff, upper_only=upper_only, return_offsets=True,
                                                        offsets=pre)
        out = {"names": names, "rows": rows, "cols": cols, "weights": weights, "offsets": offsets}
        if dense:
            out["adjacency"] = device.to_host(threshold_similarity(sim, pearsoncutoff), pinned=False)
        return out
    # r does not fit (250 000 transcripts: 250 GB): row blocks of r are formed, scanned for edges and dropped
    block = max(128, (budget // (n * 4)) // 128 * 128)
    buf = device.empty((min(block, n), n), torch.float32)
    parts, row_offsets, dense_rows, total = [], [np.zeros(1, dtype=np.int64)], [], 0
    for row0 in range(0, n, block):
        nrows = min(block, n - row0)
        pre = similarity_matrix_and_offsets(prepared, row0, nrows, prepared, buf, pearsoncutoff, upper_only)
        r, c, w, off = similarity_edges(buf[:nrows], pearsoncutoff, upper_only=upper_only, return_offsets=True, row0=row0,
                                        offsets=pre)
        parts.append((r, c, w))
        row_offsets.append(off[1:] + total)
        total += int(off[-1])
        if dense:
            dense_rows.append(device.to_host(threshold_similarity(buf[:nrows], pearsoncutoff, row0=row0).contiguous(),
                                             pinned=False))
    out = {"names": names, "rows": np.concatenate([p[0] for p in parts]), "cols": np.concatenate([p[1] for p in parts]),
           "weights": np.concatenate([p[2] for p in parts]), "offsets": np.concatenate(row_offsets)}
    if dense:
        out["adjacency"] = np.concatenate(dense_rows, axis=0)
    return out


def kmer_leiden(inputfile, mean, std, k, algo='RBERVertexPartition', rs=1.0, pearsoncutoff=0, setseed=False,
                edgecolormethod='gradient', edgethreshold=0.1, labelfontsize=12, plotname=None, csvfile=None):
    """Same arguments as the reference's ``kmer_leiden`` (seekr/kmer_leiden.py:64-300).

    The numeric front half -- counts, all-pairs r, threshold, edge list -- runs on the device (``leiden_inputs``);
    the community detection is leidenalg's, on an igraph graph built from the edge list instead of from n x n Python
    lists.  Both libraries are the reference's own dependencies and are imported here, when the function is called.
    ``csvfile`` writes the reference's two files (``_nodes_leiden.csv``: Id, Label, Color = community number;
    ``_edges_leiden.csv``: every upper-triangle pair with its thresholded weight); plotting (``plotname``) follows the
    reference's networkx / matplotlib drawing when those are installed.  Both need the dense matrix on the host.
    Returns None like the reference."""
    import pandas as pd

    inputs = leiden_inputs(inputfile, mean, std, k, pearsoncutoff=pearsoncutoff, upper_only=True,
                           dense=bool(plotname or csvfile))
    if inputs is None:
        return None
    try:
        import igraph as ig
        import leidenalg
    except ImportError as exc:  # not a dependency of the hot path: say what is missing instead of failing obscurely
        raise ImportError("kmer_leiden needs python-igraph and leidenalg for the community detection (%s); "
                          "leiden_inputs() returns the graph they would be given" % exc)
    names = inputs["names"]
    graph = ig.Graph(n=len(names), edges=list(zip(inputs["rows"].tolist(), inputs["cols"].tolist())), directed=False)
    graph.es["weight"] = inputs["weights"].tolist()
    algos = {name: getattr(leidenalg, name) for name in
             ("ModularityVertexPartition", "RBConfigurationVertexPartition", "RBERVertexPartition", "CPMVertexPartition",
              "SurpriseVertexPartition", "SignificanceVertexPartition")}
    seed = 1 if setseed is True else None
    if algo == "SignificanceVertexPartition":
        partition = leidenalg.find_partition(graph, algos[algo], seed=seed)
    elif algo in ("SurpriseVertexPartition", "ModularityVertexPartition"):
        partition = leidenalg.find_partition(graph, algos[algo], weights="weight", seed=seed)
    else:
        partition = leidenalg.find_partition(graph, algos[algo], weights="weight", resolution_parameter=rs, seed=seed)
    if plotname:
        _plot_communities(inputs, partition, edgecolormethod, edgethreshold, labelfontsize, plotname)
    if csvfile:
        labels, colors = [], []
        for i, community in enumerate(partition):  # kmer_leiden.py:318-328
            for node_index in community:
                labels.append(names[node_index])
                colors.append(i + 1)
        pd.DataFrame({"Id": labels, "Label": labels, "Color": colors}).to_csv(f"{csvfile}_nodes_leiden.csv", index=False)
        df = pd.DataFrame(inputs["adjacency"], columns=names, index=names)
        mask = np.triu(np.ones(df.shape), k=1).astype(bool)
        df_triu = df.where(mask).stack().reset_index()
        df_triu.columns = ["Source", "Target", "Weight"]
        df_triu.to_csv(f"{csvfile}_edges_leiden.csv", index=False)
    return None


def _plot_communities(inputs, partition, edgecolormethod, edgethreshold, labelfontsize, plotname):
    """The reference's drawing (kmer_leiden.py:150-297): networkx spring layout, edges shaded by weight or by a
    threshold, nodes coloured by community.  Library work on the host; needs networkx and matplotlib."""
    import matplotlib.pyplot as plt
    import networkx as nx
    import pandas as pd

    names = inputs["names"]
    df = pd.DataFrame(inputs["adjacency"], columns=names, index=names)
    graph = nx.from_pandas_adjacency(df)
    weights = inputs["weights"]
    if edgecolormethod == "threshold":
        edge_colors = ["black" if w > edgethreshold else "grey" for w in weights]
        edge_widths = [4 if w > edgethreshold else 1 for w in weights]
    else:
        if edgecolormethod != "gradient":
            print("edgecolormethod must be either 'gradient' or 'threshold', use default 'gradient' now")
        normalized = (weights - weights.min()) / (weights.max() - weights.min())
        mapped = 0.1 + 0.9 * normalized
        edge_colors = [(1 - w, 1 - w, 1 - w) for w in mapped]
        edge_widths = [(1 + 3 * w) for w in mapped]
    community_colors = plt.cm.rainbow(np.linspace(0, 1, max(partition.membership) + 1))
    node_colors = [community_colors[c] for c in partition.membership]
    pos = nx.spring_layout(graph, weight="weight")
    plt.figure(figsize=(15, 15))
    plt.gca().axis("off")
    nx.draw_networkx_nodes(graph, pos, node_color=node_colors, node_size=500)
    nx.draw_networkx_edges(graph, pos, edge_color=edge_colors, width=edge_widths)
    nx.draw_networkx_labels(graph, pos, font_size=labelfontsize, font_family="sans-serif")
    plt.tight_layout()
    plt.savefig(f"{plotname}.pdf")

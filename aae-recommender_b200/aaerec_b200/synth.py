"""Synthetic item-set data of the shapes SURVEY.md section 8(d) names.

The reference ships no datasets (only the shapes logged in ``nmi.txt``), so benchmarks
and tests use seeded synthetic sets: set lengths ``clip(Poisson(mean_len))``, items
drawn from a Zipf(``zipf``) popularity law, duplicates inside a set dropped (the
reference requires binary sets: ``Bags.load_tabcomma_format(..., unique=True)``,
main.py:63).  Output is what ``BagsWithVocab.tocsr()`` (datasets.py:459-470) hands to
the recommender: a scipy CSR matrix of ones with sorted column indices.
"""
import numpy as np
import scipy.sparse as sp


def synth_sets(n, V, mean_len, min_len=2, max_len=None, seed=0, zipf=1.0, dtype=np.float32, len_choices=None):
    """``len_choices``: draw each set's length uniformly from these values instead of clip(Poisson(mean_len)) -- the
    MPD challenge's query sets hold 1 / 5 / 10 / 25 / 100 seed tracks (eval/mpd/create_dev_set.py:16-17)."""
    rs = np.random.RandomState(seed)
    if len_choices is not None:
        lens = rs.choice(np.asarray(len_choices, dtype=np.int64), size=n)
    else:
        lens = rs.poisson(mean_len, n)
        lens = np.clip(lens, min_len, max_len if max_len else None)
    lens = np.minimum(lens, V).astype(np.int64)
    w = 1.0 / np.arange(1, V + 1, dtype=np.float64) ** zipf
    cdf = np.cumsum(w)
    cdf /= cdf[-1]
    total = int(lens.sum())
    rows = np.repeat(np.arange(n, dtype=np.int64), lens)
    items = np.minimum(np.searchsorted(cdf, rs.random_sample(total)), V - 1).astype(np.int64)
    # popularity rank -> item id through a fixed permutation so that popular items are spread
    # over the whole id range (as in a real vocabulary) instead of clustered at id 0
    perm = np.random.RandomState(seed + 7919).permutation(V)
    items = perm[items]
    key = np.unique(rows * V + items)           # sorted by (row, item), duplicates dropped
    rows_u = key // V
    items_u = (key % V).astype(np.int32)
    counts = np.bincount(rows_u, minlength=n)
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    data = np.ones(items_u.shape[0], dtype=dtype)
    idx_dtype = np.int32 if indptr[-1] < 2 ** 31 else np.int64
    X = sp.csr_matrix((data, items_u, indptr.astype(idx_dtype)), shape=(n, V))
    X.has_sorted_indices = True
    return X


def synth_condition(n, dim=300, seed=2, scale=0.1):
    """Title-embedding-shaped condition rows (random, word2vec-sized)."""
    return (np.random.RandomState(seed).randn(n, dim) * scale).astype(np.float32)


SHAPES = {
    # name: (n_sets, n_items, mean_len, min_len, max_len, seed)   -- SURVEY 8(d)
    "econbiz": (61607, 4587, 5, 2, 30, 0),
    "pubmed": (50000, 200000, 16, 2, 200, 1),
    "mpd": (20000, 2000000, 66, 5, 250, 3),
}


def synth_named(name, n=None):
    n0, V, mean_len, lo, hi, seed = SHAPES[name]
    return synth_sets(n or n0, V, mean_len, lo, hi, seed)

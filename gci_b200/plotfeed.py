"""Plot feed: the data behind the reference's depth plots, without matplotlib.

`sliding_window_average_depth` is `GCI.py:660-705` restated with numpy instead of a per-base Python loop:
same arguments, same stderr warning, same return value — the
positions in Mbp as a list and the averaged depths as an array — so the reference's plotting code can be fed
from it unchanged.  Plotting itself stays out of scope.  Host code: nothing here is on the scored hot path.
"""
from __future__ import annotations

import sys

import numpy as np


def sliding_window_average_depth(depths=[], window_size=50000, max_depth=None, start=0, target=None):
    """GCI.py:660-705.  Every zero-depth base is a point of its own and restarts the window; a window that is
    still open when a zero (or the end of the region) arrives is averaged over what it holds and reported at the
    position of its last base; averages above `max_depth` are clipped to it."""
    d = np.asarray(depths)
    n = int(d.shape[0])
    if n < window_size:
        print(f'Warning!!! The length ({len(depths)}) of plotting region ({target}:{start}-{start + len(depths)}) is '
              f'less than the window size ({window_size}), and therefore the window size will be 1 bp', file=sys.stderr)
        window_size = 1
    if n == 0:
        return [], np.array([])
    d = d.astype(np.int64, copy=False)
    zero = d == 0
    # maximal runs of non-zero depth
    edge = np.diff(np.concatenate(([0], (~zero).astype(np.int8), [0])))
    run_s = np.flatnonzero(edge == 1)
    run_e = np.flatnonzero(edge == -1)                     # exclusive
    run_len = run_e - run_s
    csum = np.concatenate(([0], np.cumsum(d)))
    n_full = run_len // window_size
    rest = run_len - n_full * window_size
    # full windows: run r, window k ends at index run_s[r] + (k + 1) * window_size - 1
    rep = np.repeat(np.arange(len(run_s)), n_full)
    k = np.arange(int(n_full.sum())) - np.repeat(np.cumsum(n_full) - n_full, n_full)
    w_lo = run_s[rep] + k * window_size
    w_hi = w_lo + window_size
    full_idx = w_hi - 1
    full_sum = csum[w_hi] - csum[w_lo]
    # the open window at the end of a run
    has_rest = rest > 0
    r_hi = run_e[has_rest]
    r_lo = r_hi - rest[has_rest]
    rest_idx = r_hi - 1
    rest_sum = csum[r_hi] - csum[r_lo]
    rest_n = rest[has_rest]
    zero_idx = np.flatnonzero(zero)
    # merge the three kinds of points by base index (each index carries at most one point)
    idx = np.concatenate((full_idx, rest_idx, zero_idx))
    kind = np.concatenate((np.zeros(len(full_idx), np.int8), np.ones(len(rest_idx), np.int8),
                           np.full(len(zero_idx), 2, np.int8)))
    num = np.concatenate((full_sum, rest_sum, np.zeros(len(zero_idx), np.int64)))
    den = np.concatenate((np.full(len(full_idx), window_size, np.int64), rest_n, np.ones(len(zero_idx), np.int64)))
    order = np.argsort(idx, kind="stable")
    idx, kind, num, den = idx[order], kind[order], num[order], den[order]
    positions = ((idx + start) / 1e6).tolist()
    # the value list is rebuilt with the reference's Python types (float averages, int zeros, max_depth as given):
    # np.array() of it then picks the same dtype as the reference's
    avg = (num / den).tolist()
    values = [0 if kd == 2 else (max_depth if a > max_depth else a) for kd, a in zip(kind.tolist(), avg)]
    return positions, np.array(values)


def sliding_window_average_depth_gpu(ctx, track, contig, lo, hi, window_size=50000, max_depth=None, start=0, target=None):
    """The same result for depth[lo:hi] of one contig of a GPU depth track (`Context.sliding_window`: zero runs from the
    run extraction, window sums from a device prefix sum); only the points come back — the division, the clip and the
    Mbp positions are finished here with the reference's Python types, so values and dtype are the reference's."""
    n = int(hi) - int(lo)
    if n < window_size:
        print(f'Warning!!! The length ({n}) of plotting region ({target}:{start}-{start + n}) is '
              f'less than the window size ({window_size}), and therefore the window size will be 1 bp', file=sys.stderr)
        window_size = 1
    if n <= 0:
        return [], np.array([])
    idx, num, den, kind = ctx.sliding_window(track, contig, lo, hi, window_size)
    positions = ((idx + start) / 1e6).tolist()
    avg = (num / den).tolist()
    values = [0 if kd == 2 else (max_depth if a > max_depth else a) for kd, a in zip(kind.tolist(), avg)]
    return positions, np.array(values)

"""The oracle's vectorised restatements against scalar, loop-level restatements of the same definitions on small
random cases (pure Python / numpy, no torch broadcasting tricks): inverse-CDF resampling, the interlevel outer
measure, the distortion integral, compositing weights, median depth and the trilinear hash interpolation."""

import math

import numpy as np
import pytest
import torch

from oracle import nerfstudio_math as M


def _sorted_bins(rng, n):
    b = np.sort(rng.random(n + 1))
    b[0], b[-1] = 0.0, 1.0
    return b


def test_pdf_resampling_against_scalar_inverse_cdf():
    rng = np.random.default_rng(0)
    for S_prev, S_new, jit in ((8, 4, None), (33, 17, 0.37), (256, 96, 0.9), (5, 11, 0.0)):
        w = rng.random(S_prev) ** 4
        w[rng.integers(0, S_prev, S_prev // 3)] = 0.0  # empty space: only the 0.01 padding remains
        bins = _sorted_bins(rng, S_prev)
        t_rand = None if jit is None else torch.tensor([[jit]], dtype=torch.float32)
        got = M.pdf_resample_bins(torch.tensor(w[None], dtype=torch.float32), torch.tensor(bins[None], dtype=torch.float32),
                                  S_new, t_rand)[0].double().numpy()
        # scalar restatement in float64
        wp = w.astype(np.float32).astype(np.float64) + 0.01
        tot = wp.sum()
        pad = max(1e-5 - tot, 0.0)
        wp, tot = wp + pad / S_prev, tot + pad
        cdf = np.concatenate([[0.0], np.minimum(1.0, np.cumsum(wp / tot))])
        nb = S_new + 1
        want = np.empty(nb)
        for j in range(nb):
            u = j * (1.0 - 1.0 / nb) / (nb - 1) + (jit / nb if jit is not None else 1.0 / (2 * nb))
            ind = int(np.searchsorted(cdf, u, side="right"))
            below, above = min(max(ind - 1, 0), S_prev), min(max(ind, 0), S_prev)
            den = cdf[above] - cdf[below]
            t = 0.0 if den == 0 else min(max((u - cdf[below]) / den, 0.0), 1.0)
            want[j] = bins[below] + t * (bins[above] - bins[below])
        # float32 vs float64 arithmetic: a sample inside an empty bin divides by a tiny CDF increment
        assert np.all(np.diff(got) >= -1e-6)
        assert np.max(np.abs(got - want)) < 5e-4, (S_prev, S_new, np.max(np.abs(got - want)))


def test_interlevel_outer_measure_against_interval_loops():
    rng = np.random.default_rng(1)
    for S, Sp in ((6, 4), (48, 96), (17, 33)):
        c, cp = _sorted_bins(rng, S), _sorted_bins(rng, Sp)
        if S == 6:
            cp[2] = c[3]  # a shared edge: the searchsorted sides decide which proposal bins count
        w, wp = rng.random(S) / S, rng.random(Sp) / Sp
        got = M.lossfun_outer(torch.tensor(c[None]), torch.tensor(w[None]), torch.tensor(cp[None]),
                              torch.tensor(wp[None]))[0].numpy()
        for i in range(S):
            lo = max(sum(1 for j in range(Sp) if cp[j] <= c[i]) - 1, 0)          # bin holding the start
            hi = min(sum(1 for j in range(Sp) if cp[j + 1] <= c[i + 1]), Sp - 1)  # first bin ending after the end
            w_outer = sum(wp[j] for j in range(lo, hi + 1))
            want = max(w[i] - w_outer, 0.0) ** 2 / (w[i] + 1e-7)
            assert got[i] == pytest.approx(want, rel=1e-9, abs=1e-12), (S, i)
            # the outer measure really bounds the mass of every proposal bin that overlaps the interval
            overlap = sum(wp[j] for j in range(Sp) if cp[j] < c[i + 1] and cp[j + 1] > c[i])
            assert w_outer >= overlap - 1e-12


def test_distortion_loss_against_the_double_integral():
    """loss = integral integral w(u) w(v) |u - v| du dv for the piecewise-constant density w_i / delta_i."""
    rng = np.random.default_rng(2)
    S = 5
    t = _sorted_bins(rng, S)
    w = rng.random(S)
    got = float(M.distortion_loss([torch.tensor(w[None, :, None])], [torch.tensor(t[None])]))
    n = 4000
    u = (np.arange(n) + 0.5) / n
    dens = np.zeros(n)
    for i in range(S):
        m = (u >= t[i]) & (u < t[i + 1])
        dens[m] = w[i] / (t[i + 1] - t[i])
    want = float((dens[:, None] * dens[None, :] * np.abs(u[:, None] - u[None, :])).sum() / n / n)
    assert got == pytest.approx(want, rel=2e-3)


def test_weights_and_median_depth_against_loops():
    rng = np.random.default_rng(3)
    S = 12
    edges = np.cumsum(rng.random(S + 1))
    deltas, sigma = np.diff(edges), rng.random(S) * 3
    w = M.get_weights(torch.tensor(deltas[None, :, None]), torch.tensor(sigma[None, :, None]))[0, :, 0].numpy()
    T, cum, median = 1.0, 0.0, None
    for i in range(S):
        alpha = 1.0 - math.exp(-deltas[i] * sigma[i])
        assert w[i] == pytest.approx(alpha * T, rel=1e-9)
        T *= math.exp(-deltas[i] * sigma[i])
        cum += alpha * (T / math.exp(-deltas[i] * sigma[i]))
        if median is None and cum >= 0.5:
            median = 0.5 * (edges[i] + edges[i + 1])
    if median is None:
        median = 0.5 * (edges[S - 1] + edges[S])
    got = float(M.render_depth_median(torch.tensor(w[None, :, None]), torch.tensor(edges[None, :-1, None]),
                                      torch.tensor(edges[None, 1:, None])))
    assert got == pytest.approx(median, rel=1e-9)


def test_hash_encoding_against_scalar_trilinear_interpolation():
    rng = np.random.default_rng(4)
    L, T = 5, 12
    scal = M.hash_scalings(L, 16, 128)
    table = torch.tensor(rng.standard_normal((L << T, 2)), dtype=torch.float32)
    x = torch.tensor(rng.random((7, 3)), dtype=torch.float32)
    got = M.hash_encode(x, table, scal, T).double().numpy()
    P1, P2 = 2654435761, 805459861
    for n in range(x.shape[0]):
        for l in range(L):
            s = np.float32(x[n].numpy()) * np.float32(scal[l])
            f, c = np.floor(s).astype(np.int64), np.ceil(s).astype(np.int64)
            off = (s - np.floor(s)).astype(np.float64)
            acc = np.zeros(2)
            for bx in (0, 1):
                for by in (0, 1):
                    for bz in (0, 1):
                        v = (c[0] if bx else f[0], c[1] if by else f[1], c[2] if bz else f[2])
                        idx = ((int(v[0]) * 1) ^ (int(v[1]) * P1) ^ (int(v[2]) * P2)) % (1 << T) + l * (1 << T)
                        wgt = (off[0] if bx else 1 - off[0]) * (off[1] if by else 1 - off[1]) * (off[2] if bz else 1 - off[2])
                        acc += wgt * table[idx].double().numpy()
            assert np.allclose(got[n, 2 * l:2 * l + 2], acc, atol=2e-6), (n, l)

"""Size-independent properties of the oracle's building blocks on randomly drawn inputs (hypothesis): the invariants
the GPU full-size tests rely on (tests/test_fullsize_gpu.py) hold for the checker itself."""

import math

import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import nerfstudio_math as M

SETTINGS = dict(max_examples=40, deadline=None, derandomize=True)  # the same examples on every run


def _rand(seed, *shape):
    return torch.rand(*shape, generator=torch.Generator().manual_seed(seed))


@settings(**SETTINGS)
@given(seed=st.integers(0, 10**6), S_in=st.integers(1, 64), S_out=st.integers(1, 64), sparsity=st.floats(0, 1),
       jittered=st.booleans())
def test_pdf_resampling_yields_sorted_bins_inside_the_parent_range(seed, S_in, S_out, sparsity, jittered):
    R = 5
    w = _rand(seed, R, S_in) * (_rand(seed + 1, R, S_in) >= sparsity)  # some (or all) bins empty
    edges = torch.sort(_rand(seed + 2, R, S_in + 1), dim=-1).values
    tr = _rand(seed + 3, R, 1) if jittered else None
    bins = M.pdf_resample_bins(w, edges, S_out, tr)
    assert bins.shape == (R, S_out + 1)
    assert torch.isfinite(bins).all()
    assert (bins[:, 1:] >= bins[:, :-1] - 1e-7).all()                       # non-decreasing
    assert (bins >= edges[:, :1] - 1e-6).all() and (bins <= edges[:, -1:] + 1e-6).all()


@settings(**SETTINGS)
@given(seed=st.integers(0, 10**6), S=st.integers(1, 96), scale=st.floats(1e-3, 50))
def test_weights_are_a_sub_probability_with_the_closed_form_mass(seed, S, scale):
    R = 4
    deltas = _rand(seed, R, S, 1) * 0.1
    sigma = _rand(seed + 1, R, S, 1) * scale
    w = M.get_weights(deltas, sigma)
    assert (w >= 0).all() and torch.isfinite(w).all()
    total = w.sum(dim=-2)
    expected = 1 - torch.exp(-(deltas * sigma).sum(dim=-2))                  # telescoping product of transmittances
    assert (total <= 1 + 1e-5).all()
    assert torch.allclose(total, expected, atol=2e-5)
    assert torch.allclose(M.render_accumulation(w), total)


@settings(**SETTINGS)
@given(seed=st.integers(0, 10**6), mag=st.floats(1e-3, 1e4))
def test_contraction_maps_into_the_radius_two_cube_and_fixes_the_unit_cube(seed, mag):
    x = (_rand(seed, 64, 3) * 2 - 1) * mag
    y = M.contract_linf(x)
    n = x.abs().amax(dim=-1)
    assert (y.abs().amax(dim=-1) < 2 + 1e-6).all()
    inside = n <= 1
    assert torch.equal(y[inside], x[inside])
    # outside: direction preserved, L-inf norm 2 - 1/|x|
    out = ~inside
    if out.any():
        assert torch.allclose(y[out].abs().amax(-1), 2 - 1 / n[out], atol=1e-5)
        assert torch.allclose(y[out] / y[out].abs().amax(-1, keepdim=True), x[out] / n[out, None], atol=1e-5)


@settings(**SETTINGS)
@given(seed=st.integers(0, 10**6), log2=st.integers(4, 19), levels=st.integers(1, 16))
def test_hash_indices_stay_inside_their_level(seed, log2, levels):
    g = torch.Generator().manual_seed(seed)
    coords = torch.randint(0, 2049, (32, levels, 3), generator=g, dtype=torch.int32)
    idx = M.hash_indices(coords, log2, levels)
    size = 1 << log2
    lvl = torch.arange(levels)[None, :]
    assert ((idx >= lvl * size) & (idx < (lvl + 1) * size)).all()
    # neighbouring cells along x differ by the xor with 1 * prime_x = 1: the low bit flips, nothing else moves
    shifted = coords.clone()
    shifted[..., 0] ^= 1
    assert torch.equal((M.hash_indices(shifted, log2, levels) - lvl * size) ^ 1, idx - lvl * size)


@settings(**SETTINGS)
@given(x=st.floats(1e-3, 1e4))
def test_spacing_function_round_trip_and_monotonicity(x):
    t = torch.tensor([x, x * 1.01], dtype=torch.float64)
    s = M.spacing_fn(t)
    assert torch.allclose(M.spacing_fn_inv(s), t, rtol=1e-9)
    assert s[1] > s[0] and 0 < s[0] < 1
    assert math.isclose(float(M.spacing_fn(torch.tensor(1.0, dtype=torch.float64))), 0.5)


@settings(**SETTINGS)
@given(seed=st.integers(0, 10**6), S=st.integers(2, 48))
def test_losses_are_non_negative_and_vanish_in_their_degenerate_cases(seed, S):
    R = 3
    edges = torch.sort(_rand(seed, R, S + 1), dim=-1).values
    w = _rand(seed + 1, R, S, 1)
    w = w / w.sum(dim=-2, keepdim=True) * 0.9
    # a level bounded by itself has no excess mass: the interlevel term is zero; any proposal >= it as well
    assert float(M.interlevel_loss([w, w], [edges, edges])) <= 1e-10
    assert float(M.interlevel_loss([w * 1.1, w], [edges, edges])) <= 1e-10
    assert float(M.interlevel_loss([w * 0.5, w], [edges, edges])) > 0
    d = float(M.distortion_loss([w], [edges]))
    assert d >= 0
    # all the mass in one vanishing interval: both distortion terms vanish with its width
    one = torch.zeros(R, S, 1)
    one[:, 0] = 1.0
    tight = torch.cat([torch.zeros(R, 1), torch.full((R, 1), 1e-6), torch.linspace(0.5, 1, S - 1).expand(R, -1)], -1)
    assert float(M.distortion_loss([one], [tight])) < 1e-5

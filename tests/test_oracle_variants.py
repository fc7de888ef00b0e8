"""The unanchored choices of libnoise 1.1.2 are switches (oracle/vg_oracle.h VGO_VAR_*): each one
changes exactly what it says and nothing else, the default stays what the committed fixtures were
made with.  CPU only; the CUDA side of the same switches is tests/test_gpu_variants.py."""
import numpy as np
import pytest

import oracle


@pytest.fixture(autouse=True)
def _reset_variant():
    oracle.set_noise_variant()
    yield
    oracle.set_noise_variant()


def _fbm_samples(flags=0, order2=None, order3=None):
    oracle.set_noise_variant(flags, order2, order3)
    s = oracle.Simplex(12345)
    f2 = oracle.Fbm(s, 6, 0.008, 1.8, 0.45)
    f3 = oracle.Fbm(s, 8, 0.04, 1.2, 0.35)
    pts = [(1000000.0 + 13.0 * i, 999871.0 - 7.0 * i, float(20 + i % 90)) for i in range(200)]
    out = (np.array([f2.sample2(x, y) for x, y, _ in pts]), np.array([f3.sample3(x, y, z) for x, y, z in pts]))
    oracle.set_noise_variant()
    return out


def test_default_is_variant_zero():
    assert oracle.noise_variant() == 0


def test_fbm_coordinate_form_differs_in_the_last_bits_only():
    """(p * f) * l vs p * (f * l): same noise to ~1e-9 at world coordinates ~1e6, different bits --
    enough to move a floor or a threshold somewhere in a region, which is why it is a switch."""
    a2, a3 = _fbm_samples(0)
    b2, b3 = _fbm_samples(oracle.VAR_FBM_RUNNING_FREQ)
    assert np.abs(a2 - b2).max() < 1e-6 and np.abs(a3 - b3).max() < 1e-6
    assert (a2 != b2).any() and (a3 != b3).any()


def test_tie_breaks_are_value_neutral():
    """The corners a tie chooses between differ by a lattice vector of the form (1, -1) / (1, -1, 0),
    which unskewing leaves at length sqrt(2): a point that ties is equidistant from both, so at
    distance >= sqrt(2) / 2, i.e. at r^2 >= 0.5 = R_SQUARED -- the tied corner contributes exactly 0
    whichever one is taken.  So the tie-breaks of the simplex traversal, unanchored as they are,
    cannot change a value: generic points and exact ties (x == y, x == y == z, two equal) alike."""
    a2, a3 = _fbm_samples(0)
    b2, b3 = _fbm_samples(oracle.VAR_TIE2_OTHER | oracle.VAR_TIE3_OTHER)
    assert np.array_equal(a2, b2) and np.array_equal(a3, b3)
    s = oracle.Simplex(7)
    rng = np.random.default_rng(3)
    ties2 = [(v, v) for v in rng.uniform(-50, 50, 300)]
    ties3 = [(v, v, v) for v in rng.uniform(-50, 50, 100)]
    ties3 += [(v, v, w) for v, w in rng.uniform(-50, 50, (100, 2))]
    ties3 += [(w, v, v) for v, w in rng.uniform(-50, 50, (100, 2))]
    d2 = np.array([s.sample2(*p) for p in ties2])
    d3 = np.array([s.sample3(*p) for p in ties3])
    oracle.set_noise_variant(oracle.VAR_TIE2_OTHER | oracle.VAR_TIE3_OTHER)
    e2 = np.array([s.sample2(*p) for p in ties2])
    e3 = np.array([s.sample3(*p) for p in ties3])
    assert np.array_equal(d2, e2) and np.array_equal(d3, e3)
    assert np.abs(d2).max() > 0.1 and np.abs(d3).max() > 0.1  # not trivially zero


def test_skew_digits_and_gradient_order_change_samples():
    a2, a3 = _fbm_samples(0)
    b2, b3 = _fbm_samples(oracle.VAR_SKEW_EXACT)
    assert (a2 != b2).any() and np.array_equal(a3, b3)  # only the 2-D constants have two spellings
    rot2 = np.roll(np.arange(24, dtype=np.uint8), 1)
    rot3 = np.roll(np.arange(12, dtype=np.uint8), 5)
    c2, c3 = _fbm_samples(0, rot2, None)
    d2, d3 = _fbm_samples(0, None, rot3)
    assert (a2 != c2).any() and np.array_equal(a3, c3)
    assert np.array_equal(a2, d2) and (a3 != d3).any()
    # the identity permutation is the default
    i2, i3 = _fbm_samples(0, np.arange(24, dtype=np.uint8), np.arange(12, dtype=np.uint8))
    assert np.array_equal(a2, i2) and np.array_equal(a3, i3)


@pytest.mark.parametrize("flags", [1, 6, 8, 15])
def test_generation_runs_under_every_variant_and_default_is_restored(flags):
    seed = oracle.hash_world_name("test")
    xy = oracle.region_xy(62499, 62499, 2, 2)
    base, base_h, _, _ = oracle.generate_areas(seed, xy)
    oracle.set_noise_variant(flags)
    vox, mh, _, _ = oracle.generate_areas(seed, xy)
    assert vox.shape == base.shape and (vox == 9).any()  # Stone everywhere below the surface
    for i in range(len(xy)):
        assert np.array_equal(mh[i], oracle.update_all_column_heights(vox[i]).reshape(-1))
    oracle.set_noise_variant()
    again, again_h, _, _ = oracle.generate_areas(seed, xy)
    assert np.array_equal(again, base) and np.array_equal(again_h, base_h)

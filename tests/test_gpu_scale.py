"""GPU parity at the BASELINE sizes (-m gpu): the engine's -o / -s / -i texts, hashed, against the digests of the
UNMODIFIED reference's outputs on the same seeded inputs (tests/golden/scale_hashes.json, written by
tests/golden/make_scale_hashes.py in the build container).  One test per BASELINE config at its real size, plus the
tie-heavy set of SURVEY.md §8d and the smaller sizes of the same generator streams."""
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _check(name, **opt):
    if name not in helpers.scale_hashes():
        pytest.skip(f"no committed digests for {name}")
    got = helpers.engine_case_outputs(name, **opt)
    assert helpers.compare_case(name, got) == [], (name, got, helpers.scale_hashes()[name])


@pytest.mark.parametrize("name", ["c2_1m", "c3_1m", "tie1m", "tie1m_f", "c4_100k"])
def test_reference_digests_small(built, name):
    _check(name)


@pytest.mark.parametrize("opt", [{"tile_rows": 1, "cluster_kernel": 4}, {"cluster_kernel": 6}, {"join_kernel": 2, "cluster_kernel": 3}, {"join_kernel": 1}, {"enum_mode": 1, "cluster_kernel": 2}])
def test_reference_digests_other_kernels(built, opt):
    _check("tie1m", **opt)


def test_c2_10m_d1(built):
    """BASELINE configs[1]: 10 M x 150 bp, d=1"""
    _check("c2")


def test_c3_10m_fastidious(built):
    """BASELINE configs[2]: 10 M x 150 bp, d=1 --fastidious (reference run with -t 1)"""
    _check("c3")


def test_c4_1m_400bp_d2(built):
    """BASELINE configs[3]: 1 M x 400 bp, d=2"""
    _check("c4")

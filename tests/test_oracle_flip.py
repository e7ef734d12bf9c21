"""CPU (not gpu): the FLIP particle <-> grid plugins (plugin/flip.cpp: markFluidCells, gridParticleIndex, unionParticleLevelset, mapPartsToMAC,
mapMACToParts, flipVelocityUpdate -- SURVEY 8f-4) restated in the oracle ahead of their device versions: bit for bit the golden vectors of
the unmodified reference, and the reference itself on another seed."""
import numpy as np
import pytest

import helpers
from helpers import FLIP_SCENES, load_golden, run_flip_plugins


@pytest.mark.parametrize("prec", [4, 8])
@pytest.mark.parametrize("name", list(FLIP_SCENES))
def test_port_reproduces_flip_golden(name, prec, port32, port64):
    g = load_golden("step_" + name, prec)
    out = run_flip_plugins(port32 if prec == 4 else port64, name, prec)
    assert set(out) == set(g)
    for key in out:
        assert np.array_equal(out[key], g[key]), (name, prec, key)
    # the fixtures are not vacuous
    assert (g["mark"] & 1).sum() > 50 and len(g["index_sys"]) > 300 and (g["union"] < 0).sum() > 50 and np.abs(g["map_vel"]).max() > 0.5
    assert not np.array_equal(g["pic"], g["flip"])


@pytest.mark.parametrize("prec", [4, 8])
def test_port_equals_reference_on_another_scene(prec, port32, port64, ref32, ref64, monkeypatch):
    P, R = (port32, ref32) if prec == 4 else (port64, ref64)
    monkeypatch.setitem(helpers.FLIP_SCENES, "other", (9, 11, 13))
    a, b = run_flip_plugins(P, "other", prec), run_flip_plugins(R, "other", prec)
    for key in a:
        assert np.array_equal(a[key], b[key]), key

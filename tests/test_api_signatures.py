"""CPU (not gpu): the Python mirror keeps the NAME, ORDER and DEFAULTS of every PYTHON() plugin it replaces (unknown or
reordered keyword arguments would break existing scenes).  The signatures are parsed from the reference sources when /root/reference
is present (and the committed fixture tests/golden/plugin_signatures.json is checked against them); without the reference tree the
fixture alone is used."""
import inspect
import json
import os
import re

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "plugin_signatures.json")
REF = "/root/reference/source"
PLUGINS = {     # plugin -> reference file
    "solvePressure": "plugin/pressure.cpp", "computePressureRhs": "plugin/pressure.cpp", "solvePressureSystem": "plugin/pressure.cpp",
    "correctVelocity": "plugin/pressure.cpp", "releaseMG": "plugin/pressure.cpp",
    "setWallBcs": "plugin/extforces.cpp", "addGravity": "plugin/extforces.cpp", "addGravityNoScale": "plugin/extforces.cpp", "addBuoyancy": "plugin/extforces.cpp",
    "advectSemiLagrange": "plugin/advection.cpp", "cgSolveDiffusion": "conjugategrad.cpp", "cgSolveWE": "plugin/waves.cpp",
    "PD_fluid_guiding": "plugin/fluidguiding.cpp",
    "extrapolateMACSimple": "fastmarch.cpp", "extrapolateLsSimple": "fastmarch.cpp", "extrapolateVec3Simple": "fastmarch.cpp", "extrapolateMACFromWeight": "fastmarch.cpp",
    "getLaplacian": "plugin/flip.cpp", "getCurvature": "plugin/flip.cpp",
    "updateFractions": "plugin/initplugins.cpp", "setObstacleFlags": "plugin/initplugins.cpp",
    "markFluidCells": "plugin/flip.cpp", "gridParticleIndex": "plugin/flip.cpp", "unionParticleLevelset": "plugin/flip.cpp", "mapPartsToMAC": "plugin/flip.cpp",
    "mapMACToParts": "plugin/flip.cpp", "flipVelocityUpdate": "plugin/flip.cpp", "pushOutofObs": "plugin/flip.cpp",
    "addForcePvel": "plugin/ptsplugins.cpp", "updateVelocityFromDeltaPos": "plugin/ptsplugins.cpp", "eulerStep": "plugin/ptsplugins.cpp", "setPartType": "plugin/ptsplugins.cpp",
    "markIsolatedFluidCell": "grid.cpp",
}


def parse_reference(name, path):
    txt = open(os.path.join(REF, path)).read()
    txt = re.sub(r"//[^\n]*", "", txt)
    m = re.search(r"PYTHON\(\)\s+void\s+%s\s*\(" % re.escape(name), txt)
    assert m, name
    i, depth, start = m.end(), 1, m.end()
    while depth:
        depth += {"(": 1, ")": -1}.get(txt[i], 0)
        i += 1
    body = txt[start:i - 1]
    params, depth, cur = [], 0, ""
    for ch in body:
        if ch == "," and depth == 0:
            params.append(cur); cur = ""
        else:
            depth += {"(": 1, ")": -1}.get(ch, 0); cur += ch
    if cur.strip():
        params.append(cur)
    out = []
    for p in params:
        p = " ".join(p.split())
        default = None
        if "=" in p:
            p, default = (t.strip() for t in p.split("=", 1))
        pname = re.findall(r"[A-Za-z_][A-Za-z_0-9]*", p)[-1]
        out.append([pname, default])
    return out


def norm_default(d):
    """C++ default -> comparable Python value"""
    if d is None:
        return inspect.Parameter.empty
    d = d.strip()
    if d in ("NULL", "0", "nullptr") or d == "0":
        return 0
    if d in ("true", "false"):
        return d == "true"
    if d == "PcMIC":
        return 1
    try:
        return float(d.rstrip("f"))
    except ValueError:
        return d


def load_signatures():
    if os.path.isdir(REF):
        sigs = {n: parse_reference(n, p) for n, p in PLUGINS.items()}
        if os.environ.get("MP_WRITE_SIGNATURE_FIXTURE"):
            json.dump(sigs, open(FIXTURE, "w"), indent=1)
        assert json.load(open(FIXTURE)) == sigs, "tests/golden/plugin_signatures.json is stale: regenerate with MP_WRITE_SIGNATURE_FIXTURE=1"
        return sigs
    return json.load(open(FIXTURE))


@pytest.mark.parametrize("name", sorted(PLUGINS))
def test_plugin_keeps_reference_signature(name):
    import mantaflow_b200 as mf      # importing the package does not need a GPU (the library is loaded on first use)
    ref = load_signatures()[name]
    sig = inspect.signature(getattr(mf, name))
    mine = list(sig.parameters.values())
    assert [p.name for p in mine] == [p[0] for p in ref], (name, [p.name for p in mine], [p[0] for p in ref])
    for p, (rname, rdef) in zip(mine, ref):
        want = norm_default(rdef)
        got = p.default
        if want is inspect.Parameter.empty:
            assert got is inspect.Parameter.empty, (name, rname, "must stay a required argument")
            continue
        if want == 0 and got is None:       # null pointer default <-> None
            continue
        if isinstance(want, bool) or isinstance(got, bool):
            assert bool(got) == bool(want), (name, rname, got, want)
        elif isinstance(want, float):
            assert got is not None and float(got) == pytest.approx(want), (name, rname, got, want)
        else:
            assert got == want, (name, rname, got, want)

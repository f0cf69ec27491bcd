"""Host-side generators and config vs the reference's own outputs (golden vectors).  CPU only."""

import json
import os

import numpy as np
import pytest

from mpc_bipedal.config import MPCConfig
from mpc_bipedal.generators import CoPGenerator, State, generate_footsteps
from mpc_bipedal.models import LIPMModel, zmp_preview_matrices, velocity_preview_matrices
from oracle import lipm_oracle as orc

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STATE_CODE = {State.STANDING: 0, State.DOUBLE_SUPPORT: 1, State.SINGLE_SUPPORT: 2}


def config_for(meta):
    meta = json.loads(str(meta))
    values = {}
    if meta["base"] == "json":
        with open(os.path.join(REPO, "configs", "default.json")) as fh:
            values = json.load(fh)["mpc"]
    values.update(meta["overrides"])
    return MPCConfig.from_dict(values)


def walk_tags(g):
    return sorted({k.split("/")[0] for k in g.files})


def test_generators_bit_identical_to_reference(golden):
    g = golden("ref_generators")
    assert len(walk_tags(g)) == 8
    for tag in walk_tags(g):
        cfg = config_for(g[f"{tag}/config"])
        steps = generate_footsteps(cfg.distance, cfg.step_length, cfg.foot_spread)
        got = np.array([[c.x, c.y, *c.z_max, *c.z_min] for c in steps])
        assert np.array_equal(got, g[f"{tag}/footsteps"]), tag
        z_max, z_min, states = CoPGenerator(cfg).generate_cop_trajectory(save_footsteps=False)
        assert np.array_equal(z_max, g[f"{tag}/z_max"]), tag
        assert np.array_equal(z_min, g[f"{tag}/z_min"]), tag
        assert np.array_equal([STATE_CODE[s] for s in states], g[f"{tag}/states"]), tag


def test_default_walk_timeline(golden):
    # n=420 = ST 100, 8 x (DS 3 + SS 24|25), DS 3, ST 100
    g = golden("ref_generators")
    s = g["default_json/states"]
    assert len(s) == 420 and (s[:100] == 0).all() and (s[-100:] == 0).all()
    runs = np.diff(np.flatnonzero(np.diff(s) != 0))
    assert set(runs[1::2]) <= {24, 25} and set(runs[0::2]) == {3}


def test_config_rules():
    cfg = MPCConfig()
    assert cfg.dt == 1.5 / 150 and cfg.dsp_duration == 0.01 and cfg.gamma == 1.0
    assert MPCConfig(horizon=300).dt == 1.5 / 300
    assert MPCConfig(dt=0.02).dt == 0.02
    cfg = MPCConfig.from_json(os.path.join(REPO, "configs", "default.json"))
    assert (cfg.dsp_duration, cfg.standing_duration, cfg.gamma, cfg.strict) == (0.03, 1.0, 0.0, True)
    assert MPCConfig.from_dict({"dt": 0.5, "horizon": 100}).dt == 1.5 / 100      # horizon wins
    with pytest.raises(TypeError):
        MPCConfig.from_dict({"no_such_key": 1})
    left, right = np.array(cfg.left_foot_polytope), np.array(cfg.right_foot_polytope)
    assert left.shape == (11, 2) and np.array_equal(right, left * [1, -1])
    with pytest.raises(ValueError):
        broken = MPCConfig()
        broken.dt = None
        CoPGenerator(broken)


def test_lipm_model_surface():
    cfg = MPCConfig(h=0.8)
    model = LIPMModel(cfg)
    A, B, C = orc.model_matrices(cfg.dt, cfg.h, cfg.g)
    assert np.array_equal(model.A, A) and np.array_equal(model.B[:, 0], B) and np.array_equal(model.C[0], C)
    assert model.B.shape == (3, 1) and model.C.shape == (1, 3) and model.get_state_dimension() == 3
    x, u = np.array([[0.1], [0.2], [0.3]]), np.array([[2.0]])
    assert np.allclose(model.step(x, u), A @ x + B[:, None] * 2.0)
    assert np.allclose(model.get_zmp(x), 0.1 - 0.3 * cfg.h / cfg.g)


@pytest.mark.parametrize("N,h", [(10, 0.75), (150, 0.75), (64, 0.6)])
def test_preview_matrices_bit_identical_to_loop(N, h):
    dt = 1.5 / N
    Px, Pu = zmp_preview_matrices(N, dt, h, 9.81)
    Px_ref, Pu_ref = orc.preview_matrices(N, dt, h, 9.81)
    assert np.array_equal(Px, Px_ref) and np.array_equal(Pu, Pu_ref)
    Pvs, Pvu = velocity_preview_matrices(N, dt)
    A, B, _ = orc.model_matrices(dt, h, 9.81)
    vel = np.array([0.0, 1.0, 0.0])
    Apow = np.eye(3)
    for i in range(N):
        assert np.allclose(np.diag(Pvu, -i), vel @ Apow @ B, rtol=1e-12)
        Apow = Apow @ A
        assert np.allclose(Pvs[i], vel @ Apow, rtol=1e-12)

"""Host-side logic that needs no GPU: C-ABI surface, controller shell, Herdt helper tables."""

import ctypes
import os
import re

import numpy as np
import pytest

from mpc_bipedal.config import MPCConfig
from mpc_bipedal.controllers import ZMPController
from mpc_bipedal.generators import State
from mpc_bipedal.generators.speed_generation import SpeedTrajectoryGenerator

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE_TO_STATE = {0: State.STANDING, 1: State.DOUBLE_SUPPORT, 2: State.SINGLE_SUPPORT}


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    from mpc_bipedal import engine

    entry.build_cuda()
    header = open(os.path.join(REPO, "include", "lipm_mpc.h")).read()
    declared = set(re.findall(r"\b((?:lipm|herdt)_[a-z_]+)\s*\(", header))
    assert declared == set(engine.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(engine.LIBRARY_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    # no compute call without a GPU, but the pure-host entry points must work
    lib.lipm_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.lipm_version()
    assert lib.lipm_max_strict_horizon() >= 150


def test_plan_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mpc_bipedal import engine
    with pytest.raises(RuntimeError, match="CUDA device"):
        engine.LipmPlan(150, 0.75, 9.81, None, 1.0, 1e-6)
    ctrl = ZMPController(MPCConfig())
    with pytest.raises(RuntimeError):
        ctrl.predict_wieber_axis(np.zeros((3, 1)), 150, np.ones((150, 1)), -np.ones((150, 1)))


def test_controller_shell_matches_reference_surface():
    cfg = MPCConfig(h=0.8, horizon=100)
    ctrl = ZMPController(cfg)
    T = cfg.dt
    assert np.array_equal(ctrl.A, [[1, T, T ** 2 / 2], [0, 1, T], [0, 0, 1]])
    assert ctrl.B.shape == (3, 1) and ctrl.C.shape == (3,) and ctrl.C[2] == -cfg.h / cfg.g
    zero = np.zeros((3, 1))
    with pytest.raises(ValueError, match="z_max and z_min are required"):
        ctrl.generate_com_trajectory(zero, zero)
    cfg.method = "herdt"
    with pytest.raises(ValueError, match="v_ref and state_ref are required"):
        ctrl.generate_com_trajectory(zero, zero)
    cfg.method = "nope"
    with pytest.raises(ValueError, match="Unknown method"):
        ctrl.generate_com_trajectory(zero, zero)


def test_find_nb_steps_matches_reference(golden):
    g = golden("ref_herdt_helpers")
    ctrl = ZMPController(MPCConfig())
    tags = sorted({k.split("/")[0] for k in g.files if k.endswith("/nb_steps")})
    assert len(tags) == 5
    for tag in tags:
        codes = g[f"{tag}/states"]
        horizon = len(g[f"{tag}/nb_steps"]) - len(codes)
        padded = np.concatenate([codes, np.repeat(codes[-1:], horizon)])
        got = ctrl.find_nb_steps(np.array([CODE_TO_STATE[int(c)] for c in padded], dtype=object))
        assert np.array_equal(np.array(got, np.int32), g[f"{tag}/nb_steps"]), tag


def test_find_nb_steps_edge_timelines():
    ctrl = ZMPController(MPCConfig())
    S, D, X = State.STANDING, State.DOUBLE_SUPPORT, State.SINGLE_SUPPORT
    assert ctrl.find_nb_steps([]) == []
    assert ctrl.find_nb_steps([S, S, S]) == [(3, 3), (2, 3), (1, 3)]
    assert ctrl.find_nb_steps([X, X, D, X]) == [(2, 4), (1, 3), (2, 2), (1, 2)]
    # expectations below were produced by the reference's own find_nb_steps on these timelines
    assert ctrl.find_nb_steps([D, D, X, S]) == [(1, 1), (3, 4), (2, 4), (1, 3)]
    assert ctrl.find_nb_steps([S, D, X, X, D, D, X, S, S]) == [
        (1, 1), (3, 3), (2, 3), (1, 3), (1, 1), (4, 5), (3, 5), (2, 4), (1, 4)]
    assert ctrl.find_nb_steps([X, S, D, S, X, D]) == [(2, 6), (2, 6), (3, 3), (3, 3), (1, 3), (1, 1)]


def test_polytope_halfspaces_match_reference(golden):
    g = golden("ref_herdt_helpers")
    cfg = MPCConfig()
    ctrl = ZMPController(cfg)
    for side in ("left", "right"):
        A, b = ctrl._polytope_halfspace(np.array(getattr(cfg, f"{side}_foot_polytope")))
        assert A.shape == (9, 2)                                  # SURVEY: 9 facets per side
        assert np.allclose(A, g[f"polytope/{side}/A"]) and np.allclose(b, g[f"polytope/{side}/b"])
    with pytest.raises(ValueError):
        ctrl._polytope_halfspace(np.zeros((2, 2)))


def test_classic_speed_reference(golden):
    g = golden("ref_herdt_helpers")
    cfg = MPCConfig.from_json(os.path.join(REPO, "configs", "default.json"))
    vx, vy, states = SpeedTrajectoryGenerator(cfg).generate_speed_and_state(save_footsteps=False)
    assert np.array_equal(vx, g["default_json/vx"]) and np.array_equal(vy, g["default_json/vy"])
    assert len(states) == len(vx) == 420
    cfg.speed_generation = "bogus"
    with pytest.raises(ValueError, match="Unknown speed_generation"):
        SpeedTrajectoryGenerator(cfg).generate_speed_and_state(save_footsteps=False)


# ----------------------------------------------------------------------------- parameter sweeps (config 5)
def test_sweep_table_groups_and_shards(default_config):
    from mpc_bipedal import sweeps

    table = sweeps.make_sweep(1000, seed=3)
    assert len(table) == 1000 and len(table.variants) == 15 and len(table.h_values) == 8
    key = table.h_index.astype(np.int64) * 100 + table.variant
    assert (np.diff(key) >= 0).all()                                  # sorted by (h, variant)
    assert sorted(table.scenario_id.tolist()) == list(range(1000))    # a permutation of the caller's order
    buckets = sweeps.bucket_ranges(table)
    assert buckets[0][1] == 0 and buckets[-1][2] == 1000 and len(buckets) == 8
    for (h_idx, a, b), nxt in zip(buckets, buckets[1:] + [None]):
        assert a < b and (table.h_index[a:b] == h_idx).all()
        if nxt is not None:
            assert nxt[1] == b and nxt[0] != h_idx
    for world in (1, 2, 3, 8):
        covered = []
        sizes = []
        for rank in range(world):
            shares = sweeps.rank_indices(table, rank, world)
            assert [h for h, _ in shares] == [h for h, _, _ in buckets]          # every rank runs every CoM height
            for (h_idx, idx), (_, a, b) in zip(shares, buckets):
                assert ((idx >= a) & (idx < b)).all() and (np.diff(idx) == world).all()
                assert abs(len(idx) - (b - a) / world) < 1                       # ... and the same share of each
                covered.extend(idx.tolist())
            sizes.append(sum(len(idx) for _, idx in shares))
        assert sorted(covered) == list(range(1000))                              # a partition of the table
        assert max(sizes) - min(sizes) <= len(buckets)
    assert sweeps.series_capacity(default_config, table.variants) >= 600        # longest series (step 0.2): 566 samples

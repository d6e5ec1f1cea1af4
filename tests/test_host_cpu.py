"""CPU-only tests of the host mirror: registry, types, no silent CPU fallback."""
import numpy as np
import pytest
import torch

from dgppo_b200.env import ENV, make_env
from dgppo_b200.trainer.data import Rollout
from dgppo_b200.trainer import distributed as D
from dgppo_b200.utils.graph import GraphsTuple


def test_make_env_matches_reference_factory():
    env = make_env("LidarSpread", num_agents=8, num_obs=8)
    assert (env.num_agents, env.params["n_obs"], env.params["n_rays"], env.params["top_k_rays"]) == (8, 8, 32, 8)
    assert (env.state_dim, env.node_dim, env.edge_dim, env.action_dim, env.n_cost) == (4, 7, 4, 2, 2)
    assert (env.dt, env.max_episode_steps, env.area_size) == (0.03, 128, 1.5)
    full = make_env("LidarSpread", num_agents=3, num_obs=3, full_observation=True)
    assert full.params["comm_radius"] == 15.0
    # unlike the reference (env/__init__.py:38-46) the class-level PARAMS is not mutated
    assert ENV["LidarSpread"].PARAMS["n_obs"] == 3 and ENV["LidarSpread"].PARAMS["comm_radius"] == 0.5
    bic = make_env("LidarBicycleTarget", num_agents=16, num_obs=3)
    assert (bic.state_dim, bic.node_dim) == (5, 8)
    mpe = make_env("MPESpread", num_agents=8)
    assert mpe.area_size == 1.5 and mpe.params["obs_radius"] == 0.05
    with pytest.raises(AssertionError):
        make_env("VMASWheel", num_agents=2)


def test_state_and_action_limits():
    lo, hi = make_env("LidarSpread", 3, num_obs=3).state_lim()
    assert lo.tolist() == [0., 0., -0.5, -0.5] and hi.tolist() == [1.5, 1.5, 0.5, 0.5]
    lo, hi = make_env("MPESpread", 3).state_lim()
    assert lo.tolist() == [0., 0., -1., -1.] and hi.tolist() == [1.5, 1.5, 1., 1.]
    lo, hi = make_env("LidarBicycleTarget", 3, num_obs=3).state_lim()
    assert lo.tolist() == [0., 0., -1., -1., -0.5]
    lo, hi = make_env("MPESpread", 3).action_lim()
    assert lo.tolist() == [-1., -1.] and hi.tolist() == [1., 1.]


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    env = make_env("LidarSpread", num_agents=3, num_obs=3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        env.reset(0)
    from dgppo_b200.algo import make_algo
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        make_algo("dgppo", env=env, node_dim=7, edge_dim=4, state_dim=4, action_dim=2, n_agents=3)
    with pytest.raises(ValueError):
        make_algo("informarl", env=env)


def _graph(b=None):
    sh = (lambda *s: ((b,) if b else ()) + s)
    nt = torch.tensor([0, 0, 1, 1, 2, 2, 2, 2, -1], dtype=torch.int32)
    return GraphsTuple(
        n_node=torch.full(sh(), 9, dtype=torch.int32), n_edge=torch.full(sh(), 5, dtype=torch.int32),
        nodes=torch.arange(int(np.prod(sh(9, 7))), dtype=torch.float32).reshape(sh(9, 7)),
        edges=torch.zeros(sh(5, 4)), states=torch.arange(int(np.prod(sh(9, 4))), dtype=torch.float32).reshape(sh(9, 4)),
        receivers=torch.zeros(sh(5), dtype=torch.int32), senders=torch.zeros(sh(5), dtype=torch.int32),
        node_type=nt.expand(sh(9)).contiguous(), env_states=None)


def test_graphs_tuple_api():
    g = _graph()
    assert g.is_single and g.n_graphs == 1 and g.batch_shape == ()
    assert len(g) == 10 and g[2] is g.nodes                    # tuple of 10 fields in the reference's order
    assert g.type_states(0, 2).shape == (2, 4) and torch.equal(g.type_states(1, 2), g.states[2:4])
    assert torch.equal(g.type_nodes(2, 4), g.nodes[4:8])
    with pytest.raises(AssertionError):
        g.type_states(2, 3)
    g2 = g._replace(env_states="x")
    assert g2.env_states == "x" and g2.nodes is g.nodes and g.without_edge().edges is None
    gb = _graph(b=3)
    assert not gb.is_single and gb.n_graphs == 3 and gb.batch_shape == (3,)
    assert gb.type_states(2, 4).shape == (3, 4, 4)


def test_rollout_record_type():
    g = _graph(b=2)
    ro = Rollout(g, torch.zeros(2, 5, 3, 2), torch.zeros(2, 5, 1, 3, 1, 64), torch.zeros(2, 5),
                 torch.zeros(2, 5, 3, 2), torch.zeros(2, 5, dtype=torch.bool), None, g)
    assert ro._fields == ("graph", "actions", "rnn_states", "rewards", "costs", "dones", "log_pis", "next_graph")
    assert (ro.length, ro.time_horizon, ro.n_data) == (2, 5, 10)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 128, 4096):
        for w in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert (D.same_shuffle(64, 5) == D.same_shuffle(64, 5)).all()


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints one JSON line with
    the contract's keys; the oracle port runs on all host cores, one process per core."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "C1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "agent-steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == min(os.cpu_count() or 1, 64)
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["higher_is_better"] is True and "workload" in line["config"]

"""GPU: the compact rollout record.  A rollout into a compact record must equal the rollout into a
full graph record bit for bit (actions, log_pi, carries, rewards, costs), its lazily built graphs must equal the
stored ones, the pre-pass products computed straight from the state record must equal the graph-record ones, and the
record must stay under 3 KB per env-step at C3."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _algo(env_id, n, obs, T, compact, batch=512, seed=3):
    from dgppo_b200.algo import make_algo
    from dgppo_b200.env import make_env
    env = make_env(env_id, num_agents=n, num_obs=obs, max_step=T)
    algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                     action_dim=env.action_dim, n_agents=n, batch_size=batch, rnn_step=16, seed=seed,
                     compact_record=compact)
    return env, algo


@pytest.mark.parametrize("env_id,n,obs,b", [("LidarSpread", 8, 8, 64), ("MPESpread", 5, 3, 24), ("LidarBicycleTarget", 4, 3, 24),
                                            ("MPEFormation", 4, 2, 16), ("LidarSpread", 20, 4, 8),
                                            ("LidarSpread", 8, 8, 2048)])
def test_compact_rollout_equals_full_record(env_id, n, obs, b):
    T = 32
    keys = np.arange(b) + 17
    _, full = _algo(env_id, n, obs, T, False)
    _, comp = _algo(env_id, n, obs, T, True)
    rf = full.collect(full.params, keys)
    rc = comp.collect(comp.params, keys)
    for k in ("actions", "log_pis", "rewards", "costs", "rnn_states"):
        assert torch.equal(getattr(rf, k), getattr(rc, k)), k
    # lazily materialised graphs == the stored ones
    for k in ("nodes", "edges", "states", "receivers", "senders", "node_type", "n_node", "n_edge"):
        assert torch.equal(getattr(rc.graph, k), getattr(rf.graph, k)), k
        assert torch.equal(getattr(rc.next_graph, k), getattr(rf.next_graph, k)), "next " + k
    assert torch.equal(rc.graph.env_states.agent, rf.graph.env_states.agent)
    assert torch.equal(rc.graph.type_states(0, n), rf.graph.type_states(0, n))
    # pre-pass products straight from the state record
    pf, pc = full.prepass(rf, 0), comp.prepass(rc, 0)
    for k in ("bTp1_Vl", "bTp1ah_Vh", "bTah_Qh", "bT_Ql", "bTa_A", "bTp1ah_Vh_det", "bTah_Qh_det"):
        assert torch.equal(pf[k], pc[k]), k
    assert torch.equal(pf["det_rollout"].actions, pc["det_rollout"].actions)


def test_compact_record_size_and_update():
    env, algo = _algo("LidarSpread", 8, 8, 128, True, batch=2048)
    ro = algo.collect(algo.params, np.arange(32))
    rec = ro.graph.record
    assert rec.compact and rec.nodes is None
    assert rec.bytes_per_env_step() <= 3 * 1024, rec.bytes_per_env_step()        # 2.9 KB (full record: 10.7 KB)
    info = algo.update(ro, 0)                                                      # minibatch graphs built by K3 on demand
    assert all(np.isfinite(v) for v in info.values())
    assert rec._materialized is None                                               # nothing built the full graph record
    # a full-record algo on the same keys takes the same update step (same data, same minibatches)
    _, ref = _algo("LidarSpread", 8, 8, 128, False, batch=2048)
    info_ref = ref.update(ref.collect(ref.params, np.arange(32)), 0)
    for k in ("Vl/loss", "Vh/loss_Vh", "policy/loss"):
        np.testing.assert_allclose(info[k], info_ref[k], rtol=1e-5, atol=1e-7)

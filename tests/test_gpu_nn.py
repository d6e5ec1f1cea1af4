"""GPU parity of K4 (policy / Vh / Vl forward) against the oracle through the
C ABI.  Tolerance: north_star's rtol 1e-5 (fp32), with an absolute floor of
1e-5 x the output scale for values near zero; the fp64 oracle is the yardstick
showing the CUDA error is of the size of fp32 rounding noise."""
import numpy as np
import pytest

from dgppo_b200 import _lib
from dgppo_b200.algo import params as P
from oracle import env_np, nn_np
from tests import util
from tests.util import CONFIGS

pytestmark = pytest.mark.gpu
F = np.float32
RTOL, ATOL = 1e-5, 1e-5


@pytest.fixture(params=["v2", "v1"], autouse=True)
def kernel_generation(request, monkeypatch):
    """Every test runs on the weight-stationary v2 kernels (default when the
    shape fits) and on the fused v1 kernel (DGPPO_FORCE_V1=1, the large-n path)."""
    monkeypatch.setenv("DGPPO_FORCE_V1", "1" if request.param == "v1" else "0")
    return request.param


def _graph(cfg, b, seed):
    agent, goal, obstacles, mpe_obs = util.threshold_states(cfg, b, seed)
    return env_np.reset_graph(cfg, agent, goal, obstacles, mpe_obs)


def _close(got, ref32, ref64, what):
    scale = max(1.0, float(np.abs(ref64).max()))
    err = np.abs(got.astype(np.float64) - ref64)
    err32 = np.abs(ref32.astype(np.float64) - ref64)
    msg = f"{what}: max err vs fp64 {err.max():.3e} (oracle fp32's own {err32.max():.3e}), scale {scale:.3g}"
    print(msg)
    np.testing.assert_allclose(got, ref32, rtol=RTOL, atol=ATOL * scale, err_msg=msg)
    assert err.max() <= 20 * max(err32.max(), 1e-7 * scale) + ATOL * scale * 0.1, msg


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4", "C5", "target", "noobs_lidar", "one_agent",
                                  "n24", "mpe20", "target40", "dense20"])
@pytest.mark.parametrize("stochastic", [True, False])
def test_policy_forward(name, stochastic):
    cfg = CONFIGS[name]
    b = 5 if cfg.n >= 64 else (11 if cfg.n > 16 else 37)           # not a multiple of the tile size
    g = _graph(cfg, b, 21)
    tree = P.init_policy_params(cfg.node_dim, 4, 2, 2, seed=3, jitter=0.1, scale_final=1.0)
    nc = P.net_cfg(_lib.NET_POLICY, cfg.node_dim, 4, 2, 2)
    packed = P.pack_params(tree, nc)
    rng = np.random.default_rng(4)
    rnn = rng.standard_normal((b, cfg.n, 64)).astype(F) * 0.5
    eps = rng.standard_normal((b, cfg.n, 2)).astype(F) if stochastic else None
    a32, lp32, h32, _ = nn_np.policy_forward(tree, g, rnn, cfg.n, eps, 2, np.float32)
    a64, lp64, h64, _ = nn_np.policy_forward(tree, g, rnn, cfg.n, eps, 2, np.float64)
    a, lp, h = util.k_policy(cfg, nc, packed, g, rnn, eps)
    _close(h, h32, h64, f"rnn {name}")
    _close(a, a32, a64, f"action {name}")
    if stochastic:
        _close(lp, lp32, lp64, f"log_pi {name}")


def test_policy_log_prob_tails():
    """Actions that saturate |a| >= 0.999 take the log-cdf branches
    (distribution.py:25-35)."""
    cfg = CONFIGS["C1"]
    b = 16
    g = _graph(cfg, b, 5)
    tree = P.init_policy_params(cfg.node_dim, 4, 2, 2, seed=9, jitter=0.1, scale_final=1.0)
    tree["params"]["OutputDenseMean"]["kernel"] = tree["params"]["OutputDenseMean"]["kernel"] * 40.0
    nc = P.net_cfg(_lib.NET_POLICY, cfg.node_dim, 4, 2, 2)
    packed = P.pack_params(tree, nc)
    rng = np.random.default_rng(1)
    rnn = rng.standard_normal((b, cfg.n, 64)).astype(F)
    eps = rng.standard_normal((b, cfg.n, 2)).astype(F)
    a32, lp32, _, _ = nn_np.policy_forward(tree, g, rnn, cfg.n, eps, 2, np.float32)
    assert (np.abs(a32) >= 0.999).any(), "test premise: some actions saturate"
    a, lp, _ = util.k_policy(cfg, nc, packed, g, rnn, eps)
    np.testing.assert_allclose(a, a32, rtol=1e-5, atol=1e-5)
    sat = (np.abs(a32) >= 0.999).any(-1) | (np.abs(a) >= 0.999).any(-1)
    np.testing.assert_allclose(lp[~sat], lp32[~sat], rtol=1e-4, atol=1e-4)
    # on saturated entries both must take the same branch unless a sits on the threshold
    agree = np.isclose(lp[sat], lp32[sat], rtol=1e-3, atol=1e-3)
    assert agree.mean() > 0.9


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4", "C5", "n24", "mpe20"])
def test_vh_forward(name):
    cfg = CONFIGS[name]
    b = 5 if cfg.n >= 64 else (11 if cfg.n > 16 else 37)
    g = _graph(cfg, b, 31)
    tree = P.init_value_params(cfg.node_dim, 4, 2, 1, seed=5, jitter=0.1)
    nc = P.net_cfg(_lib.NET_VH, cfg.node_dim, 4, 1, 2)
    packed = P.pack_params(tree, nc)
    rnn = np.random.default_rng(2).standard_normal((b, cfg.n, 64)).astype(F) * 0.5
    v32 = nn_np.vh_forward(tree, g, rnn, cfg.n, 1, np.float32)
    v64 = nn_np.vh_forward(tree, g, rnn, cfg.n, 1, np.float64)
    v, _ = util.k_value(cfg, nc, packed, g, rnn)
    _close(v, v32, v64, f"Vh {name}")


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4", "C5", "n24", "mpe20", "target40"])
def test_vl_forward(name):
    cfg = CONFIGS[name]
    b = 5 if cfg.n >= 64 else (11 if cfg.n > 16 else 37)
    g = _graph(cfg, b, 41)
    tree = P.init_value_params(cfg.node_dim, 4, 1, 2, seed=6, jitter=0.1)
    nc = P.net_cfg(_lib.NET_VL, cfg.node_dim, 4, 2, 1)
    packed = P.pack_params(tree, nc)
    rnn = np.random.default_rng(3).standard_normal((b, 64)).astype(F) * 0.5
    v32, h32 = nn_np.vl_forward(tree, g, rnn, cfg.n, 2, np.float32)
    v64, h64 = nn_np.vl_forward(tree, g, rnn, cfg.n, 2, np.float64)
    v, h = util.k_value(cfg, nc, packed, g, rnn)
    _close(v, v32, v64, f"Vl {name}")
    _close(h, h32, h64, f"Vl carry {name}")


@pytest.mark.parametrize("head", ["wr4", "wr8", "wide"])
def test_head_kernel_variants(head, monkeypatch, kernel_generation):
    """The 64-row head kernels (DGPPO_HEAD=wr4|wr8) stay selectable for A/B runs and as the plan when
    the 128-row kernel's shared memory does not fit: same results as the default."""
    if kernel_generation == "v1":
        pytest.skip("DGPPO_HEAD only selects among the v2 head kernels")
    monkeypatch.setenv("DGPPO_HEAD", head)
    cfg = CONFIGS["C3"]
    b = 37
    g = _graph(cfg, b, 77)
    tree = P.init_policy_params(cfg.node_dim, 4, 2, 2, seed=13, jitter=0.1, scale_final=1.0)
    nc = P.net_cfg(_lib.NET_POLICY, cfg.node_dim, 4, 2, 2)
    packed = P.pack_params(tree, nc)
    rng = np.random.default_rng(5)
    rnn = rng.standard_normal((b, cfg.n, 64)).astype(F) * 0.5
    eps = rng.standard_normal((b, cfg.n, 2)).astype(F)
    a32, lp32, h32, _ = nn_np.policy_forward(tree, g, rnn, cfg.n, eps, 2, np.float32)
    a64, lp64, h64, _ = nn_np.policy_forward(tree, g, rnn, cfg.n, eps, 2, np.float64)
    a, lp, h = util.k_policy(cfg, nc, packed, g, rnn, eps)
    _close(h, h32, h64, f"rnn {head}")
    _close(a, a32, a64, f"action {head}")
    _close(lp, lp32, lp64, f"log_pi {head}")


@pytest.mark.parametrize("name", ["C1", "C3", "mpe20"])
def test_vl_scan_matches_slot_by_slot_and_oracle(name):
    """dgppo_vl_scan (GNN layers of all slots in one launch, GRU head slot by slot) against the oracle's
    recurrent replay (informarl.py:281-293) and against dgppo_gnn_value called slot by slot (bit-identical:
    the same kernels see the same rows).  Runs on the split v2 kernels and on the v1 fallback loop."""
    import ctypes as C
    import torch
    from tests.util import dev, p, stream
    cfg = CONFIGS[name]
    b, S = 7, 5
    graphs = [_graph(cfg, b, 100 + t) for t in range(S)]
    tree = P.init_value_params(cfg.node_dim, 4, 1, 2, seed=6, jitter=0.1)
    nc = P.net_cfg(_lib.NET_VL, cfg.node_dim, 4, 2, 1)
    packed = P.pack_params(tree, nc)
    h0 = (np.random.default_rng(3).standard_normal(64) * 0.3).astype(F)
    stack = {k: np.stack([g[k] for g in graphs], axis=1) for k in ("nodes", "edges", "receivers", "senders")}
    nodes, edges = dev(stack["nodes"]), dev(stack["edges"])
    recv, send = dev(stack["receivers"], torch.int32), dev(stack["senders"], torch.int32)
    carry = torch.zeros((b, S + 1, 64), device="cuda")
    carry[:, 0] = dev(h0)
    val = torch.empty((b, S), device="cuda")
    pk, cc = dev(packed), util.c_cfg(cfg)
    _lib.check(_lib.lib().dgppo_vl_scan(stream(), C.byref(cc), C.byref(nc), p(pk), p(nodes), p(edges), p(recv), p(send),
                                         S, p(carry), S + 1, p(val), S, S, b), "dgppo_vl_scan")
    torch.cuda.synchronize()
    val_h, carry_h = val.cpu().numpy(), carry.cpu().numpy()
    h32 = np.broadcast_to(h0, (b, 64)).copy()
    h_slot = h32.copy()
    for t in range(S):
        v32, h32 = nn_np.vl_forward(tree, graphs[t], h32, cfg.n, 2, np.float32)
        np.testing.assert_allclose(val_h[:, t], v32, rtol=2e-5, atol=2e-5, err_msg=f"Vl slot {t}")
        np.testing.assert_allclose(carry_h[:, t + 1], h32, rtol=2e-5, atol=2e-5, err_msg=f"carry slot {t}")
        v_slot, h_slot = util.k_value(cfg, nc, packed, graphs[t], h_slot)
        util.assert_bits_equal(val_h[:, t], v_slot.astype(F), f"scan vs slot-by-slot value {t}")
        util.assert_bits_equal(carry_h[:, t + 1], h_slot, f"scan vs slot-by-slot carry {t}")

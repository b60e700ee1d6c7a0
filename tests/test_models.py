"""The ray-free networks (hhmarl_2d_b200/models.py) against golden forward passes of the reference's own
model classes (tests/golden/gen_golden_models.py), for rollout (T=1) and training-shaped (T=5) batches."""
import os

import numpy as np
import pytest
import torch

from hhmarl_2d_b200 import models as M

G = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "models_forward.npz")))


@pytest.mark.parametrize("name,n_out", [("Fight1", 26), ("Fight2", 24), ("Esc1", 26), ("Esc2", 24)])
def test_forward_matches_reference_models(name, n_out):
    model = M.fill_from_seed(getattr(M, name)(), 100 + n_out + len(name))
    for tag, T in (("t1", 1), ("t5", 5)):
        obs = {k: torch.from_numpy(G[f"{name}_{tag}_{k}"]) for k in ("obs_1_own", "obs_2", "act_1_own", "act_2")}
        B = obs["obs_1_own"].shape[0]
        with torch.no_grad():
            logits, _ = model({"obs": obs}, None, torch.tensor([T] * (B // T)))
            val = model.value_function()
            flat = torch.cat([obs["act_1_own"], obs["act_2"], obs["obs_1_own"], obs["obs_2"]], dim=1)
            l2, v2 = model.forward_flat(flat, torch.tensor([T] * (B // T)))
        np.testing.assert_allclose(logits.numpy(), G[f"{name}_{tag}_logits"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(val.numpy(), G[f"{name}_{tag}_value"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(l2.numpy(), logits.numpy(), rtol=0, atol=0)
        assert logits.shape == (B, n_out) and val.shape == (B,)


def test_shared_layer_and_parameter_names():
    p1, p2 = M.build_policy_pair("fight")
    assert p1.shared_layer is p2.shared_layer                       # SHARED_LAYER singleton, A.6.16
    names = {n for n, _ in p1.named_parameters()}
    for n in ("inp1._model.0.weight", "att_act.in_proj_weight", "att_val.out_proj.bias", "shared_layer._model.0.weight",
              "v3._model.0.bias", "val_out._model.0.weight"):
        assert n in names
    a = M.deterministic_actions(torch.randn(5, 26), 1)
    assert a.shape == (5, 4) and (a[:, 0] < 13).all() and (a[:, 3] < 2).all()


def test_commander_gru_matches_reference_model():
    model = M.fill_from_seed(M.CommanderGru(), 777)
    for tag, T in (("t1", 1), ("t4", 4)):
        obs = {k: torch.from_numpy(G[f"Cmd_{tag}_{k}"]) for k in ("obs_1_own", "obs_2", "obs_3", "act_1_own", "act_2", "act_3")}
        B = obs["obs_1_own"].shape[0]
        state = [torch.from_numpy(G[f"Cmd_{tag}_h0"]), torch.from_numpy(G[f"Cmd_{tag}_h1"])]
        with torch.no_grad():
            logits, ns = model({"obs": obs}, state, torch.tensor([T] * (B // T)))
            val = model.value_function()
        np.testing.assert_allclose(logits.numpy(), G[f"Cmd_{tag}_logits"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(val.numpy(), G[f"Cmd_{tag}_value"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ns[0].numpy(), G[f"Cmd_{tag}_nh0"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(ns[1].numpy(), G[f"Cmd_{tag}_nh1"], rtol=1e-5, atol=1e-6)
        assert logits.shape == (B, 3)


@pytest.mark.parametrize("mode", ["fight", "escape"])
def test_packed_pair_forward_equals_per_model_forward(mode):
    from hhmarl_2d_b200.fused_forward import PackedPolicyPair
    torch.manual_seed(1)
    m1, m2 = M.build_policy_pair(mode)
    M.fill_from_seed(m1, 5); M.fill_from_seed(m2, 6)
    packed = PackedPolicyPair(m1, m2)
    B = 64
    f1, f2 = torch.rand(B, m1.central_dim), torch.rand(B, m2.central_dim)
    with torch.no_grad():
        l1, v1 = m1.forward_flat(f1)
        l2, v2 = m2.forward_flat(f2)
        p1, q1, p2, q2 = packed.forward(f1, f2)
    for a, b in ((l1, p1), (v1, q1), (l2, p2), (v2, q2)):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=2e-5, atol=2e-6)
    # re-packing follows weight updates
    with torch.no_grad():
        m1.act_out._model[0].weight.add_(0.01)
    packed.refresh()
    with torch.no_grad():
        np.testing.assert_allclose(m1.forward_flat(f1)[0].numpy(), packed.forward(f1, f2)[0].numpy(), rtol=2e-5, atol=2e-6)


def test_fragment_order_packing_of_the_fused_forward_weights():
    """FusedPolicyPair._to_fragments: [K, N] row-major -> [K/8][N/8][lane = 4 g + t][(b0, b1)] with
    b0 = w[8 ks + t, 8 nt + g], b1 = w[8 ks + t + 4, 8 nt + g] (the mma.m16n8k8 B operand of every lane)."""
    import torch
    from hhmarl_2d_b200.fused_forward import FusedPolicyPair
    K, N = 24, 40
    w = torch.arange(K * N, dtype=torch.float32).reshape(K, N)
    dst = torch.zeros(K, N)
    FusedPolicyPair._to_fragments(dst, w)
    flat = dst.view(-1)
    for ks in range(K // 8):
        for nt in range(N // 8):
            for g in range(8):
                for t in range(4):
                    for q in range(2):
                        idx = ((ks * (N // 8) + nt) * 32 + g * 4 + t) * 2 + q
                        assert flat[idx].item() == w[ks * 8 + t + 4 * q, nt * 8 + g].item()

"""FusedAdam (tt_adam_step, one launch for all parameter tensors) against torch.optim.Adam, the optimizer of the
reference's training script (train/train.py:179).  fp32 elementwise arithmetic in a different association order than
torch's foreach kernels: parameters agree to a few ulp per step (rtol 2e-6 after 5 steps)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(*s, generator=g).cuda().requires_grad_(True) for s in shapes]


@pytest.mark.parametrize("weight_decay", [0.0, 0.01])
def test_fused_adam_matches_torch_adam(weight_decay):
    import two_tower_models_b200 as tt

    shapes = [(1000, 128), (256, 128), (256,), (7,), (33, 5), (1,), (4099,)]  # vectorised, ragged and tiny tensors
    ours, ref = _params(shapes, 1), _params(shapes, 1)
    o1 = tt.FusedAdam(ours, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    o2 = torch.optim.Adam(ref, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=weight_decay)
    g = torch.Generator().manual_seed(2)
    for step in range(5):
        for a, b in zip(ours, ref):
            gr = torch.randn(a.shape, generator=g).cuda() * (0.0 if step == 3 else 1.0)  # an all-zero gradient step too
            a.grad, b.grad = gr.clone(), gr.clone()
        o1.step()
        o2.step()
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), float((a - b).abs().max())
        assert torch.allclose(o1.state[a]["exp_avg_sq"], o2.state[b]["exp_avg_sq"], rtol=5e-6, atol=1e-10)


def test_fused_adam_in_cuda_graph_and_on_the_model():
    """The step counter lives on the device, so forward + backward + optimizer replay as one CUDA graph and follow
    the same trajectory as eager torch.optim.Adam on a copy of the model."""
    import two_tower_models_b200 as tt
    from two_tower_models_b200.graph import GraphedTrainStep

    d, F, B, hash_size = 64, 64, 512, 1000
    torch.manual_seed(0)
    m1 = tt.TwoTowerBaseRetrieval(10, hash_size, d, F, hash_size, d, F, [1.0], tt.BaselineMIPSModule(64, d)).cuda()
    m2 = tt.TwoTowerBaseRetrieval(10, hash_size, d, F, hash_size, d, F, [1.0], tt.BaselineMIPSModule(64, d)).cuda()
    m2.load_state_dict(m1.state_dict())
    gen = torch.Generator().manual_seed(1)
    batch = dict(
        user_id=torch.randint(0, hash_size, (B,), generator=gen), user_features=torch.randn(B, F, generator=gen),
        user_history=torch.randint(0, hash_size, (B, 4), generator=gen), item_id=torch.randint(0, hash_size, (B,), generator=gen),
        item_features=torch.randn(B, F, generator=gen), position=torch.randint(0, 100, (B,), generator=gen),
        labels=torch.randint(0, 2, (B, 1), generator=gen).float())
    batch = {k: v.cuda() for k, v in batch.items()}
    order = ["user_id", "user_features", "user_history", "item_id", "item_features", "position", "labels"]
    opt2 = torch.optim.Adam(m2.parameters(), lr=1e-3)
    start = {k: v.detach().clone() for k, v in m1.state_dict().items()}
    opt1 = tt.FusedAdam(m1.parameters(), lr=1e-3)
    gstep = GraphedTrainStep(m1, batch, warmup=2, optimizer=opt1)  # 2 eager warm-up updates; capturing executes nothing
    n_replays = 3
    for _ in range(n_replays):
        gstep(batch)
    for _ in range(2 + n_replays):
        opt2.zero_grad(set_to_none=True)
        m2._packed.invalidate()
        m2.train_forward(*[batch[k] for k in order]).backward()
        opt2.step()
    torch.cuda.synchronize()
    # gradients that are analytically zero (sum_j dS_ij = 0 reaches these two biases unmasked) are rounding noise, and
    # Adam's normalisation turns noise into +-lr steps: their trajectories are not comparable
    noise_only = {"item_tower_arch.bias", "item_features_arch.2.bias"}
    for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        if k in noise_only:
            continue
        upd_a, upd_b = (a - start[k]).double(), (b - start[k]).double()
        assert float(upd_b.norm()) > 0, k
        # same gradients from the same kernels up to the order of the split-K atomics; Adam normalises every element
        # to a step of about lr, so elements whose gradient is near zero may step the other way: compare the updates
        # in the Frobenius norm
        rel = float((upd_a - upd_b).norm() / upd_b.norm())
        assert rel < 0.1, (k, rel)


def test_fused_adam_eager_loop_sees_fresh_weights():
    """Plain eager loop (train_forward, backward, FusedAdam.step) WITHOUT any manual cache invalidation: the optimizer
    writes the parameters through raw pointers, so it has to bump their version for the bf16 operand caches
    (ops.PackedWeights keys on `_version`).  The loss must move and follow torch.optim.Adam on a copy of the model."""
    import two_tower_models_b200 as tt

    d, F, B, hash_size = 64, 64, 512, 1000
    torch.manual_seed(0)
    m1 = tt.TwoTowerBaseRetrieval(10, hash_size, d, F, hash_size, d, F, [1.0], tt.BaselineMIPSModule(64, d)).cuda()
    m2 = tt.TwoTowerBaseRetrieval(10, hash_size, d, F, hash_size, d, F, [1.0], tt.BaselineMIPSModule(64, d)).cuda()
    m2.load_state_dict(m1.state_dict())
    gen = torch.Generator().manual_seed(1)
    batch = [torch.randint(0, hash_size, (B,), generator=gen), torch.randn(B, F, generator=gen),
             torch.randint(0, hash_size, (B, 4), generator=gen), torch.randint(0, hash_size, (B,), generator=gen),
             torch.randn(B, F, generator=gen), torch.randint(0, 100, (B,), generator=gen),
             torch.randint(0, 2, (B, 1), generator=gen).float()]
    batch = [t.cuda() for t in batch]
    o1, o2 = tt.FusedAdam(m1.parameters(), lr=3e-3), torch.optim.Adam(m2.parameters(), lr=3e-3)
    l1, l2 = [], []
    for _ in range(6):
        v0 = m1.user_tower_arch.weight._version
        o1.zero_grad(set_to_none=True)
        loss = m1.train_forward(*batch)
        loss.backward()
        o1.step()
        assert m1.user_tower_arch.weight._version > v0
        l1.append(float(loss))
        o2.zero_grad(set_to_none=True)
        m2._packed.invalidate()  # the reference trajectory re-casts by hand; m1 must not need this
        loss2 = m2.train_forward(*batch)
        loss2.backward()
        o2.step()
        l2.append(float(loss2))
    assert l1[-1] < l1[0] - 1e-3, l1  # the same batch six times: the loss has to fall
    for a, b in zip(l1, l2):
        assert abs(a - b) <= 2e-3 * abs(b), (l1, l2)

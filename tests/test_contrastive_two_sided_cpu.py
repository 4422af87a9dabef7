"""Host logic of the opt-in "two_sided" contrastive backend (b200mm.contrastive._ContrastiveTwoSidedFn) over the emulated kernels:
same loss value and gradients as the oracle and as the default backend, for symmetric InfoNCE (with the log-temperature gradient) and
MIL-NCE, with padded batches. (world-2 gloo coverage: tests/test_distributed_cpu.py runs the sharded-loss worker under both backends.)"""
import pytest
import torch

from oracle import restated
from tests import emulated_ops

BF = torch.bfloat16


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))


@pytest.fixture
def two_sided():
    import b200mm.contrastive as C

    prev = C.get_backend()
    C.set_backend("two_sided")
    yield C
    C.set_backend(prev)


@pytest.mark.parametrize("B,E", [(6, 32), (37, 64), (64, 16)])
def test_clip_loss_and_gradients(two_sided, B, E):
    C = two_sided
    g = torch.Generator().manual_seed(B)
    a = torch.nn.functional.normalize(torch.randn(B, E, generator=g), dim=-1).to(BF)
    b = torch.nn.functional.normalize(torch.randn(B, E, generator=g), dim=-1).to(BF)
    ls = torch.tensor(2.3, requires_grad=True)
    with emulated_ops.patched():
        a1, b1 = a.clone().requires_grad_(), b.clone().requires_grad_()
        loss = C.clip_contrastive_loss(a1, b1, ls)
        (loss * 1.7).backward()   # a non-unit upstream gradient reaches the kernels as a device scalar
        C.set_backend("gathered_grad")
        a0, b0 = a.clone().requires_grad_(), b.clone().requires_grad_()
        ls0 = torch.tensor(2.3, requires_grad=True)
        (C.clip_contrastive_loss(a0, b0, ls0) * 1.7).backward()
    af, bf_, lsf = a.float().requires_grad_(), b.float().requires_grad_(), torch.tensor(2.3, requires_grad=True)
    ref = restated.symmetric_info_nce(lsf.exp() * af @ bf_.t())
    (ref * 1.7).backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert _rel(a1.grad, af.grad) < 6e-3 and _rel(b1.grad, bf_.grad) < 6e-3
    assert abs(float(ls.grad) - float(lsf.grad)) < 2e-3 * max(1.0, abs(float(lsf.grad)))
    # at least as close to the oracle as the default backend (the positive's term is added in fp32 instead of being rounded into G)
    assert _rel(a1.grad, af.grad) < 1.5 * _rel(a0.grad, af.grad) + 1e-3
    assert abs(float(ls.grad) - float(ls0.grad)) < 2e-3 * max(1.0, abs(float(ls0.grad)))


@pytest.mark.parametrize("B,E", [(4, 32), (37, 64)])
def test_mil_nce_loss_and_gradients(two_sided, B, E):
    C = two_sided
    g = torch.Generator().manual_seed(100 + B)
    v = torch.nn.functional.normalize(torch.randn(B, E, generator=g), dim=-1).to(BF)
    t = torch.nn.functional.normalize(torch.randn(B, E, generator=g), dim=-1).to(BF)
    with emulated_ops.patched():
        v1, t1 = v.clone().requires_grad_(), t.clone().requires_grad_()
        loss = C.mil_nce_loss(v1, t1)
        loss.backward()
    vf, tf = v.float().requires_grad_(), t.float().requires_grad_()
    ref = restated.mil_nce_n1(restated.l1_simi_matrix(tf, vf, 1).view(B, B))
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert _rel(v1.grad, vf.grad) < 6e-3 and _rel(t1.grad, tf.grad) < 6e-3


def test_backend_selection_is_validated():
    import b200mm.contrastive as C

    with pytest.raises(ValueError):
        C.set_backend("nccl")
    assert C.get_backend() in ("gathered_grad", "two_sided")

"""GPU parity of the fused loss / validation-metric kernels (pytest -m gpu) through the C ABI, against the oracle's
restatement of `loss` / `training_step` (oracle/restate.py: focal_loss, training_loss; pinned on the unmodified
reference by tests/test_oracle_golden.py) evaluated in fp64 with torch autograd.

Tolerances: loss 2e-6 relative, logit gradients 1e-5 of the largest entry (fp32 exp / log against fp64), accuracies
exact, probabilities 1e-6.  AUROC: the reference delegates to torchmetrics (absent here, parity unpinned); it is held
to scikit-learn's one-vs-rest macro roc_auc_score instead, 1e-6.
"""
import pytest
import torch

from dune_transformercvn_b200 import loss as tloss
from dune_transformercvn_b200.config import PathOptions
from oracle import restate

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch.device("cuda:0")


def _case(b, l, e, p, seed, pad=0.3):
    g = torch.Generator().manual_seed(seed)
    ev = torch.randn(b, e, generator=g) * 2
    pr_lb = torch.randn(l, b, p, generator=g) * 2        # the network produces (slot, event) order and returns a view
    ev_t = torch.randint(0, e, (b,), generator=g)
    pr_t = torch.randint(0, p, (b, l), generator=g)
    pr_t[torch.rand(b, l, generator=g) < pad] = -1
    pr_t[:, 0].clamp_(min=0)
    return ev, pr_lb, ev_t, pr_t


@pytest.mark.parametrize("gamma", [0.0, 1.0, 2.0, 0.5])
@pytest.mark.parametrize("shape", [(2, 3, 4, 8), (16, 10, 4, 8), (256, 20, 4, 8), (5, 1, 3, 5)])
def test_fused_loss_matches_oracle(dev, gamma, shape):
    b, l, e, p = shape
    ev, pr_lb, ev_t, pr_t = _case(b, l, e, p, seed=b * 7 + l)
    opts = PathOptions.tutorial()
    opts.loss_gamma = gamma
    # oracle: fp64 autograd
    o_ev = ev.double().requires_grad_(True)
    o_pr = pr_lb.double().transpose(0, 1).requires_grad_(True)
    want = restate.training_loss(o_ev, o_pr, ev_t, pr_t, opts)
    want.backward()
    sel = pr_t >= 0
    want_ev = restate.focal_loss(o_ev.detach(), ev_t, gamma)
    want_pr = restate.focal_loss(o_pr.detach()[sel], pr_t[sel], gamma)
    # ours: the transposed view, as network.forward returns it
    d_ev = ev.to(dev).requires_grad_(True)
    d_lb = pr_lb.to(dev).requires_grad_(True)
    got, stats = tloss.training_loss(d_ev, d_lb.transpose(0, 1), ev_t.to(dev), pr_t.to(dev), opts)
    (got * 3.0).backward()     # upstream factor goes through tcvn_loss_backward
    s = stats.cpu().double()
    assert abs(float(got.detach()) - float(want.detach())) <= 2e-6 * abs(float(want.detach()))
    assert abs(float(s[1]) - float(want_ev)) <= 2e-6 * abs(float(want_ev))
    assert abs(float(s[2]) - float(want_pr)) <= 2e-6 * abs(float(want_pr))
    assert float(s[5]) == b and float(s[6]) == int(sel.sum())
    ge, gp = o_ev.grad * 3.0, o_pr.grad * 3.0
    assert float((d_ev.grad.cpu().double() - ge).abs().max()) <= 1e-5 * float(ge.abs().max())
    got_gp = d_lb.grad.cpu().double().transpose(0, 1)
    assert float((got_gp - gp).abs().max()) <= 1e-5 * float(gp.abs().max())
    assert float(got_gp[~sel].abs().max() if (~sel).any() else 0.0) == 0.0
    # accuracies as training_step logs them
    acc_e = float((ev.argmax(1) == ev_t).float().mean())
    acc_p = float((pr_lb.transpose(0, 1)[sel].argmax(1) == pr_t[sel]).float().mean())
    assert abs(float(s[3]) - acc_e) < 1e-6 and abs(float(s[4]) - acc_p) < 1e-6


def test_fused_loss_is_bit_reproducible_and_rejects_cpu(dev):
    ev, pr_lb, ev_t, pr_t = _case(64, 10, 4, 8, seed=3)
    opts = PathOptions.tutorial()
    args = (ev.to(dev), pr_lb.to(dev).transpose(0, 1), ev_t.to(dev), pr_t.to(dev), opts)
    a = tloss.training_loss(*args)[1].cpu()
    for _ in range(3):
        assert torch.equal(a, tloss.training_loss(*args)[1].cpu())
    from dune_transformercvn_b200.lib import TcvnError
    with pytest.raises(TcvnError):
        tloss.training_loss(ev, pr_lb.transpose(0, 1), ev_t, pr_t, opts)


def test_device_metrics_match_torch_and_sklearn(dev):
    from sklearn.metrics import roc_auc_score
    m = tloss.DeviceMetrics()
    evs, prs, ets, pts = [], [], [], []
    for step, (b, l) in enumerate([(32, 6), (32, 9), (17, 4)]):
        ev, pr_lb, ev_t, pr_t = _case(b, l, 4, 8, seed=100 + step)
        # make the scores informative so that the AUC is not 0.5 +- noise
        ev[torch.arange(b), ev_t] += 1.0
        m.update(ev.to(dev), pr_lb.to(dev).transpose(0, 1), ev_t.to(dev), pr_t.to(dev))
        sel = pr_t >= 0
        evs.append(torch.softmax(ev, -1)); ets.append(ev_t)
        prs.append(torch.softmax(pr_lb.transpose(0, 1)[sel], -1)); pts.append(pr_t[sel])
    got = m.compute()
    ev_p, ev_t, pr_p, pr_t = torch.cat(evs), torch.cat(ets), torch.cat(prs), torch.cat(pts)
    assert abs(got["event_epoch_accuracy"] - float((ev_p.argmax(1) == ev_t).float().mean())) < 1e-7
    assert abs(got["prong_epoch_accuracy"] - float((pr_p.argmax(1) == pr_t).float().mean())) < 1e-7
    # probabilities the kernel stored
    got_ev_p = torch.cat([p for p, _ in m.ev]).cpu()
    assert float((got_ev_p - ev_p).abs().max()) < 1e-6
    want_e = roc_auc_score(ev_t.numpy(), ev_p.double().numpy(), multi_class="ovr", average="macro")
    want_p = roc_auc_score(pr_t.numpy(), (pr_p.double() / pr_p.double().sum(1, keepdim=True)).numpy(), multi_class="ovr",
                           average="macro")
    assert abs(got["event_epoch_AUC"] - want_e) < 1e-6
    assert abs(got["prong_epoch_AUC"] - want_p) < 1e-5
    m.reset()
    assert int(m.counters.sum()) == 0 and not m.ev


def test_fused_training_step_logs_what_the_reference_logs(dev):
    """loss.fused_training_step bound on a stand-in trainer (Lightning is not installed): same logged names and values
    as training_step (neutrino_full_base_trainer.py:162-192), returned loss differentiable."""
    import types
    ev, pr_lb, ev_t, pr_t = _case(8, 5, 4, 8, seed=9)
    d_ev = ev.to(dev).requires_grad_(True)
    logged = {}
    me = types.SimpleNamespace(gamma=1.0, event_loss_scale=0.9, prong_loss_scale=0.1,
                               shared_step=lambda batch: (ev_t.to(dev), pr_t.to(dev), d_ev, pr_lb.to(dev).transpose(0, 1)),
                               log=lambda k, v: logged.__setitem__(k, float(v)))
    total = tloss.fused_training_step(me, None, 0)
    total.backward()
    opts = PathOptions.tutorial()
    want = restate.training_loss(ev.double(), pr_lb.double().transpose(0, 1), ev_t, pr_t, opts)
    assert set(logged) == {"prong_loss", "event_loss", "train_loss", "train_event_accuracy", "train_prong_accuracy"}
    assert abs(logged["train_loss"] - float(want)) < 2e-6 * abs(float(want))
    assert abs(0.9 * logged["event_loss"] + 0.1 * logged["prong_loss"] - logged["train_loss"]) < 1e-6
    assert d_ev.grad is not None and float(d_ev.grad.abs().sum()) > 0


def test_fused_loss_matches_the_reference_golden(dev, golden_dir):
    """tcvn_loss_forward against tests/golden/loss.pt: losses, accuracies and logit gradients the UNMODIFIED reference
    `loss` / `training_step` arithmetic produced in fp64 (oracle/make_golden_loss.py)."""
    import os
    import types
    cases = torch.load(os.path.join(golden_dir, "loss.pt"))
    for name, c in cases.items():
        ev = c["event_logits"].to(dev).requires_grad_(True)
        pr = c["prong_logits"].to(dev).requires_grad_(True)
        opts = types.SimpleNamespace(loss_gamma=c["gamma"], event_prong_loss_proportion=c["event_scale"])
        total, stats = tloss.training_loss(ev, pr, c["event_targets"].to(dev), c["prong_targets"].to(dev), opts)
        total.backward()
        s = stats.cpu().double()
        for got, want in ((float(total.detach()), c["total"]), (float(s[1]), c["event_loss"]), (float(s[2]), c["prong_loss"])):
            assert abs(got - want) < 2e-6 * abs(want), name
        assert abs(float(s[3]) - c["event_accuracy"]) < 1e-6 and abs(float(s[4]) - c["prong_accuracy"]) < 1e-6, name
        ge, gp = c["d_event_logits"], c["d_prong_logits"]
        assert float((ev.grad.cpu().double() - ge).abs().max()) < 1e-5 * float(ge.abs().max()), name
        assert float((pr.grad.cpu().double() - gp).abs().max()) < 1e-5 * float(gp.abs().max()), name

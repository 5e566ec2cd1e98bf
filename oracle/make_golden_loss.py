"""TEST INFRASTRUCTURE - golden vectors for the loss row: runs the UNMODIFIED reference
`NeutrinoFullBaseTrainer.loss` and the masking / mixing / accuracy arithmetic of `training_step`
(transformercvn/network/trainers/neutrino_full_base_trainer.py:148-192) in fp64 on seeded logits, and freezes inputs, the
three losses, the two accuracies and d total / d logits (autograd through the reference code) in tests/golden/loss.pt.
Build container only (needs /root/reference)."""
import importlib
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_import  # noqa: E402


def main():
    reference_import.load()
    base = importlib.import_module("transformercvn.network.trainers.neutrino_full_base_trainer")
    loss_fn = base.NeutrinoFullBaseTrainer.loss
    cases = {}
    for name, seed, b, l, e, p, gamma, scale in (("tutorial_gamma1", 1, 16, 10, 4, 8, 1.0, 0.9), ("cross_entropy", 2, 8, 6, 4, 8, 0.0, 0.9),
                                                 ("gamma2", 3, 5, 20, 4, 8, 2.0, 0.5), ("gamma_half", 4, 3, 1, 10, 8, 0.5, 0.7)):
        g = torch.Generator().manual_seed(seed)
        ev = (torch.randn(b, e, generator=g) * 2).double().requires_grad_(True)
        pr = (torch.randn(b, l, p, generator=g) * 2).double().requires_grad_(True)
        ev_t = torch.randint(0, e, (b,), generator=g)
        pr_t = torch.randint(0, p, (b, l), generator=g)
        pr_t[torch.rand(b, l, generator=g) < 0.3] = -1
        pr_t[:, 0].clamp_(min=0)
        me = types.SimpleNamespace(gamma=gamma)
        # training_step :162-192, verbatim sequence of operations on the reference's own loss
        event_loss = loss_fn(me, ev, ev_t)
        prong_mask = pr_t >= 0
        masked_logits = torch.masked_select(pr, prong_mask.unsqueeze(-1)).reshape(-1, pr.shape[-1])
        masked_targets = torch.masked_select(pr_t.long(), prong_mask)
        prong_loss = loss_fn(me, masked_logits, masked_targets)
        total = scale * event_loss + (1.0 - scale) * prong_loss
        total.backward()
        cases[name] = {"event_logits": ev.detach().float(), "prong_logits": pr.detach().float(), "event_targets": ev_t,
                       "prong_targets": pr_t, "gamma": gamma, "event_scale": scale, "total": float(total),
                       "event_loss": float(event_loss), "prong_loss": float(prong_loss),
                       "event_accuracy": float((ev.argmax(1) == ev_t).float().mean()),
                       "prong_accuracy": float((masked_logits.argmax(1) == masked_targets).float().mean()),
                       "d_event_logits": ev.grad.clone(), "d_prong_logits": pr.grad.clone()}
        print(name, float(total), float(event_loss), float(prong_loss))
    torch.save(cases, os.path.join(ROOT, "tests", "golden", "loss.pt"))


if __name__ == "__main__":
    main()

"""Loss of ``training_step`` and the metric state of ``validation_step`` on the device.

What it replaces in the reference (transformercvn/network/trainers/neutrino_full_base_trainer.py):
  * ``loss`` + the masked_select / log_softmax / softmax / argmax chain of ``training_step`` (:148-192, about 30 small
    ATen kernels and one host sync for the masked_select) -> :func:`training_loss`, ONE kernel forward
    (``tcvn_loss_forward`` also leaves d loss / d logits behind) and one in backward;
  * ``validation_step`` (:194-209) and the torchmetrics ``Accuracy`` state -> :class:`DeviceMetrics`
    (``tcvn_metrics_update``: probabilities + hit counters, no host sync until ``compute()``).
There is no CPU implementation: CPU tensors raise.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import lib as _lib

# layout of the 8 floats tcvn_loss_forward writes
LOSS_FIELDS = ("train_loss", "event_loss", "prong_loss", "train_event_accuracy", "train_prong_accuracy",
               "event_rows", "prong_rows")


def _prep(ev_logits, pr_logits, ev_targets, pr_targets):
    for t, what in ((ev_logits, "event_logits"), (pr_logits, "prong_logits"), (ev_targets, "event_targets"),
                    (pr_targets, "prong_targets")):
        _lib.require_cuda(t, what)
    if ev_logits.dim() != 2 or pr_logits.dim() != 3 or pr_logits.shape[0] != ev_logits.shape[0]:
        raise _lib.TcvnError(f"loss: event logits {tuple(ev_logits.shape)} / prong logits {tuple(pr_logits.shape)}: "
                             "expected (B, E) and (B, L, P)")
    if tuple(pr_targets.shape) != tuple(pr_logits.shape[:2]) or tuple(ev_targets.shape) != (ev_logits.shape[0],):
        raise _lib.TcvnError(f"loss: targets {tuple(ev_targets.shape)} / {tuple(pr_targets.shape)} do not match the logits")
    ev = ev_logits.detach().float().contiguous()
    pr = pr_logits.detach().float()
    if pr.stride(2) != 1 and pr.numel() > 0:
        pr = pr.contiguous()
    return ev, pr, ev_targets.long().contiguous(), pr_targets.long().contiguous()


class _FocalLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ev_logits, pr_logits, ev_targets, pr_targets, gamma, ev_scale, pr_scale):
        L = _lib.load()
        ev, pr, et, pt = _prep(ev_logits, pr_logits, ev_targets, pr_targets)
        b, e = ev.shape
        _, l, p = pr.shape
        dev = ev.device
        out = torch.empty(8, dtype=torch.float32, device=dev)
        g_ev = torch.empty((b, e), dtype=torch.float32, device=dev)
        g_pr = torch.empty((b, l, p), dtype=torch.float32, device=dev)
        _lib.check(L.tcvn_loss_forward(_lib.ptr(ev), _lib.ptr(et), b, e, _lib.ptr(pr), _lib.ptr(pt), l, p, pr.stride(0),
                                       pr.stride(1), float(gamma), float(ev_scale), float(pr_scale), _lib.ptr(out),
                                       _lib.ptr(g_ev), _lib.ptr(g_pr), _lib.stream_ptr(dev)), "tcvn_loss_forward")
        ctx.save_for_backward(g_ev, g_pr)
        ctx.dtypes = (ev_logits.dtype, pr_logits.dtype)
        return out

    @staticmethod
    def backward(ctx, d_out):
        L = _lib.load()
        g_ev, g_pr = ctx.saved_tensors
        d_ev, d_pr = torch.empty_like(g_ev), torch.empty_like(g_pr)
        up = d_out.contiguous().float()   # only element 0 (the total loss) is differentiable
        _lib.check(L.tcvn_loss_backward(_lib.ptr(up), _lib.ptr(g_ev), g_ev.numel(), _lib.ptr(g_pr), g_pr.numel(),
                                        _lib.ptr(d_ev), _lib.ptr(d_pr), _lib.stream_ptr(g_ev.device)), "tcvn_loss_backward")
        return d_ev.to(ctx.dtypes[0]), d_pr.to(ctx.dtypes[1]), None, None, None, None, None


def focal_loss_mix(ev_logits, pr_logits, ev_targets, pr_targets, gamma: float, event_scale: float,
                   prong_scale: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """``event_scale * loss(event) + prong_scale * loss(prong rows with target >= 0)`` and the stats vector."""
    out = _FocalLossFn.apply(ev_logits, pr_logits, ev_targets, pr_targets, float(gamma), float(event_scale), float(prong_scale))
    return out[0], out.detach()


def fused_training_step(self, batch, batch_idx):
    """Drop-in body for ``NeutrinoFullBaseTrainer.training_step`` (neutrino_full_base_trainer.py:162-192), bound by
    ``install(fused_loss=True)``: same logged names, same returned total loss, one kernel instead of the
    masked_select / log_softmax / softmax / argmax chain (and no host sync)."""
    event_targets, prong_targets, event_logits, prong_logits = self.shared_step(batch)
    total, stats = focal_loss_mix(event_logits, prong_logits, event_targets, prong_targets, self.gamma,
                                  self.event_loss_scale, self.prong_loss_scale)
    self.log("prong_loss", stats[2])
    self.log("event_loss", stats[1])
    self.log("train_loss", stats[0])
    self.log("train_event_accuracy", stats[3])
    self.log("train_prong_accuracy", stats[4])
    return total


def training_loss(ev_logits, pr_logits, ev_targets, pr_targets, options) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(total_loss, stats)``: total = a * event_loss + (1 - a) * prong_loss with a = ``event_prong_loss_proportion``
    (neutrino_full_base_trainer.py:67-68,177), differentiable w.r.t. both logits; ``stats`` is the detached device
    vector named by :data:`LOSS_FIELDS` (what ``training_step`` logs), read it without forcing a sync per step."""
    a = float(options.event_prong_loss_proportion)
    return focal_loss_mix(ev_logits, pr_logits, ev_targets, pr_targets, options.loss_gamma, a, 1.0 - a)


def _auroc_macro(prob: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """One-vs-rest ROC AUC averaged over the classes that occur (exact: Mann-Whitney U with tie-averaged ranks, the
    area under the un-binned ROC curve torchmetrics' AUROC(task="multiclass", thresholds=None) integrates)."""
    n, c = prob.shape
    aucs = []
    for k in range(c):
        pos = target == k
        n_pos = int(pos.sum())
        if n_pos == 0 or n_pos == n:
            continue
        s = prob[:, k].double()
        order = torch.argsort(s)
        ss = s[order]
        # average ranks over ties
        uniq, inv, cnt = torch.unique_consecutive(ss, return_inverse=True, return_counts=True)
        end = torch.cumsum(cnt, 0).double()
        avg_rank = end - (cnt.double() - 1) / 2
        ranks = torch.empty(n, dtype=torch.float64, device=prob.device)
        ranks[order] = avg_rank[inv]
        u = ranks[pos].sum() - n_pos * (n_pos + 1) / 2
        aucs.append(u / (n_pos * (n - n_pos)))
    if not aucs:
        return torch.tensor(float("nan"), device=prob.device)
    return torch.stack(aucs).mean().float()


class DeviceMetrics:
    """Accuracy / AUROC state of ``validation_step`` kept on the device: ``update`` is one kernel launch, nothing is
    read back until ``compute`` (``validation_epoch_end``, neutrino_full_base_trainer.py:211-230)."""

    def __init__(self):
        self.counters = None
        self.ev: List[Tuple[torch.Tensor, torch.Tensor]] = []
        self.pr: List[Tuple[torch.Tensor, torch.Tensor]] = []

    def reset(self) -> None:
        if self.counters is not None:
            self.counters.zero_()
        self.ev.clear()
        self.pr.clear()

    def update(self, ev_logits, pr_logits, ev_targets, pr_targets) -> None:
        L = _lib.load()
        ev, pr, et, pt = _prep(ev_logits, pr_logits, ev_targets, pr_targets)
        b, e = ev.shape
        _, l, p = pr.shape
        dev = ev.device
        if self.counters is None or self.counters.device != dev:
            self.counters = torch.zeros(4, dtype=torch.int64, device=dev)
        ev_prob = torch.empty((b, e), dtype=torch.float32, device=dev)
        pr_prob = torch.empty((b, l, p), dtype=torch.float32, device=dev)
        _lib.check(L.tcvn_metrics_update(_lib.ptr(ev), _lib.ptr(et), b, e, _lib.ptr(pr), _lib.ptr(pt), l, p, pr.stride(0),
                                         pr.stride(1), _lib.ptr(self.counters), _lib.ptr(ev_prob), _lib.ptr(pr_prob),
                                         _lib.stream_ptr(dev)), "tcvn_metrics_update")
        self.ev.append((ev_prob, et))
        self.pr.append((pr_prob.view(b * l, p), pt.view(b * l)))

    def compute(self) -> Dict[str, float]:
        if self.counters is None:
            raise _lib.TcvnError("DeviceMetrics.compute() before any update()")
        c = self.counters.tolist()   # the one host sync of the validation epoch
        ev_p = torch.cat([p for p, _ in self.ev]); ev_t = torch.cat([t for _, t in self.ev])
        pr_p = torch.cat([p for p, _ in self.pr]); pr_t = torch.cat([t for _, t in self.pr])
        sel = pr_t >= 0
        ev_auc = float(_auroc_macro(ev_p, ev_t))
        pr_auc = float(_auroc_macro(pr_p[sel], pr_t[sel]))
        ev_acc = c[0] / max(c[1], 1)
        pr_acc = c[2] / max(c[3], 1)
        return {"event_epoch_accuracy": ev_acc, "prong_epoch_accuracy": pr_acc, "val_epoch_accuracy": (ev_acc + pr_acc) / 2,
                "event_epoch_AUC": ev_auc, "prong_epoch_AUC": pr_auc, "val_epoch_AUC": (ev_auc + pr_auc) / 2}

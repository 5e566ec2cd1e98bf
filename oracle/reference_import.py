"""TEST INFRASTRUCTURE — imports the UNMODIFIED reference from /root/reference (build container only).

The reference's dense path only touches MinkowskiEngine / h5py / pytorch_lightning /
torchmetrics at import time; none is installed here, so four empty stand-in
modules are registered before the import.  No reference source is copied.  The GPU box has no
/root/reference: nothing under tests/ -m gpu, smoke() or bench.py may call this module.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TCVN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "transformercvn"))


def _stub(name: str, **attrs):
    if name in sys.modules:
        return
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m


def load():
    """Returns a namespace with the reference symbols on the hot path."""
    import torch
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _stub("MinkowskiEngine", SparseTensor=type("SparseTensor", (), {}))
    _stub("h5py", File=type("File", (), {}))
    _stub("pytorch_lightning", LightningModule=torch.nn.Module)
    _stub("torchmetrics", Accuracy=object, AUROC=object)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from transformercvn.options import Options
    from transformercvn.network.networks.neutrino_full_dense_network import NeutrinoDenseNetwork
    from transformercvn.network.trainers import neutrino_full_dense_trainer as dense_trainer
    from transformercvn.network.layers.dense_net import DenseNet
    from transformercvn.dataset.minkowski_dataset import MinkowskiCollection
    ns = types.SimpleNamespace()
    ns.Options = Options
    ns.NeutrinoDenseNetwork = NeutrinoDenseNetwork
    ns.DenseNet = DenseNet
    ns.sparse_to_dense = dense_trainer.sparse_to_dense
    ns.dense_trainer = dense_trainer
    ns.MinkowskiCollection = MinkowskiCollection
    ns.tutorial_options = lambda: Options.load(
        os.path.join(REFERENCE_ROOT, "option_files", "fdhd_beam_2018prod_aiml_tutorial_2025_04_21.json"))
    return ns

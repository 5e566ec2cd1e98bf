"""Swaps the two names the reference's dense trainer looks up for the CUDA-backed drop-ins.

The reference binds them at import / call time by name:
  * ``NeutrinoDenseNetwork`` is imported into ``neutrino_full_dense_trainer`` (that file, line 10) and
    instantiated by ``create_network`` (:28-44);
  * ``sparse_to_dense`` is a module global of the same file (:15), looked up by ``preprocess_pixels`` (:66)
    and imported by both notebooks.
``install()`` rebinds both on the already-importable reference package; ``uninstall()`` restores them.
"""
from __future__ import annotations

import importlib
from typing import Dict

_saved: Dict[str, object] = {}
_TRAINER = "transformercvn.network.trainers.neutrino_full_dense_trainer"
_NETWORK = "transformercvn.network.networks.neutrino_full_dense_network"


def install() -> None:
    from .ingest import sparse_to_dense
    from .network import NeutrinoDenseNetwork
    trainer = importlib.import_module(_TRAINER)
    network = importlib.import_module(_NETWORK)
    if not _saved:
        _saved["trainer.NeutrinoDenseNetwork"] = trainer.NeutrinoDenseNetwork
        _saved["trainer.sparse_to_dense"] = trainer.sparse_to_dense
        _saved["network.NeutrinoDenseNetwork"] = network.NeutrinoDenseNetwork
    trainer.NeutrinoDenseNetwork = NeutrinoDenseNetwork
    trainer.sparse_to_dense = sparse_to_dense
    network.NeutrinoDenseNetwork = NeutrinoDenseNetwork


def uninstall() -> None:
    if not _saved:
        return
    trainer = importlib.import_module(_TRAINER)
    network = importlib.import_module(_NETWORK)
    trainer.NeutrinoDenseNetwork = _saved["trainer.NeutrinoDenseNetwork"]
    trainer.sparse_to_dense = _saved["trainer.sparse_to_dense"]
    network.NeutrinoDenseNetwork = _saved["network.NeutrinoDenseNetwork"]
    _saved.clear()

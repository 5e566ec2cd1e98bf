"""Swaps the names the reference's trainers look up for the CUDA-backed drop-ins.

The reference binds them at import / call time by name:
  * ``NeutrinoDenseNetwork`` is imported into ``neutrino_full_dense_trainer`` (that file, line 10) and
    instantiated by ``create_network`` (:28-44);
  * ``sparse_to_dense`` is a module global of the same file (:15), looked up by ``preprocess_pixels`` (:66)
    and imported by both notebooks;
  * ``NeutrinoSDXLNetwork`` is imported into ``neutrino_full_sdxl_trainer`` (:4) from
    ``networks/neutrino_full_sdxl_network`` - whose own import of ``diffusers`` fails where that package is absent; the
    drop-in does not need it, so ``install()`` then provides that module itself;
  * ``training_step`` is a method of ``NeutrinoFullBaseTrainer`` (neutrino_full_base_trainer.py:162); with
    ``fused_loss=True`` it is replaced by the one-kernel loss (loss.fused_training_step).
``install()`` rebinds them on the already-importable reference package; ``uninstall()`` restores them.
"""
from __future__ import annotations

import importlib
import sys
import types
from typing import Dict

_saved: Dict[str, object] = {}
_TRAINER = "transformercvn.network.trainers.neutrino_full_dense_trainer"
_NETWORK = "transformercvn.network.networks.neutrino_full_dense_network"
_SDXL_TRAINER = "transformercvn.network.trainers.neutrino_full_sdxl_trainer"
_SDXL_NETWORK = "transformercvn.network.networks.neutrino_full_sdxl_network"
_BASE_TRAINER = "transformercvn.network.trainers.neutrino_full_base_trainer"


def install(sdxl: bool = True, fused_loss: bool = False) -> None:
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
    if sdxl:
        from .sdxl import NeutrinoSDXLNetwork
        try:
            sdxl_network = importlib.import_module(_SDXL_NETWORK)
            _saved.setdefault("sdxl_network.NeutrinoSDXLNetwork", sdxl_network.NeutrinoSDXLNetwork)
        except ImportError:      # diffusers absent: the reference module cannot be imported, ours does not need it
            sdxl_network = types.ModuleType(_SDXL_NETWORK)
            sys.modules[_SDXL_NETWORK] = sdxl_network
            _saved.setdefault("sdxl_network.module_created", True)
        sdxl_network.NeutrinoSDXLNetwork = NeutrinoSDXLNetwork
        sdxl_trainer = importlib.import_module(_SDXL_TRAINER)
        _saved.setdefault("sdxl_trainer.NeutrinoSDXLNetwork", sdxl_trainer.NeutrinoSDXLNetwork)
        sdxl_trainer.NeutrinoSDXLNetwork = NeutrinoSDXLNetwork
    if fused_loss:
        from .loss import fused_training_step
        base = importlib.import_module(_BASE_TRAINER)
        _saved.setdefault("base.training_step", base.NeutrinoFullBaseTrainer.training_step)
        base.NeutrinoFullBaseTrainer.training_step = fused_training_step


def uninstall() -> None:
    if not _saved:
        return
    trainer = importlib.import_module(_TRAINER)
    network = importlib.import_module(_NETWORK)
    trainer.NeutrinoDenseNetwork = _saved["trainer.NeutrinoDenseNetwork"]
    trainer.sparse_to_dense = _saved["trainer.sparse_to_dense"]
    network.NeutrinoDenseNetwork = _saved["network.NeutrinoDenseNetwork"]
    if "sdxl_trainer.NeutrinoSDXLNetwork" in _saved:
        importlib.import_module(_SDXL_TRAINER).NeutrinoSDXLNetwork = _saved["sdxl_trainer.NeutrinoSDXLNetwork"]
    if "sdxl_network.NeutrinoSDXLNetwork" in _saved:
        importlib.import_module(_SDXL_NETWORK).NeutrinoSDXLNetwork = _saved["sdxl_network.NeutrinoSDXLNetwork"]
    if _saved.get("sdxl_network.module_created"):
        sys.modules.pop(_SDXL_NETWORK, None)
        sys.modules.pop(_SDXL_TRAINER, None)
    if "base.training_step" in _saved:
        importlib.import_module(_BASE_TRAINER).NeutrinoFullBaseTrainer.training_step = _saved["base.training_step"]
    _saved.clear()

"""install() swaps the names the reference trainer resolves (CPU; needs /root/reference, skipped elsewhere)."""
import pytest

from oracle import reference_import


@pytest.mark.skipif(not reference_import.available(), reason="reference tree not present (GPU box)")
def test_install_rebinds_reference_names(tutorial_options):
    ref = reference_import.load()
    from dune_transformercvn_b200 import install
    from dune_transformercvn_b200.network import NeutrinoDenseNetwork
    from dune_transformercvn_b200.ingest import sparse_to_dense
    orig = ref.dense_trainer.NeutrinoDenseNetwork
    install.install()
    try:
        assert ref.dense_trainer.NeutrinoDenseNetwork is NeutrinoDenseNetwork
        assert ref.dense_trainer.sparse_to_dense is sparse_to_dense
        # the reference's own create_network builds OUR class from ITS Options object
        net = ref.dense_trainer.NeutrinoFullDenseTrainer.create_network(None, ref.tutorial_options(), 1, 1, 3, 8, 4)
        assert isinstance(net, NeutrinoDenseNetwork)
        # and a reference state_dict loads strictly into it
        torch_ref = orig(ref.tutorial_options(), 1, 1, 3, 8, 4)
        res = net.load_state_dict(torch_ref.state_dict(), strict=True)
        assert not res.missing_keys and not res.unexpected_keys
    finally:
        install.uninstall()
    assert ref.dense_trainer.NeutrinoDenseNetwork is orig


@pytest.mark.skipif(not reference_import.available(), reason="reference tree not present (GPU box)")
def test_install_covers_the_sdxl_trainer_and_the_fused_training_step(tutorial_options):
    """The reference's --sdxl trainer cannot even be imported here (diffusers is absent); after install() its own
    create_network builds OUR class.  fused_loss=True swaps training_step and uninstall() restores it."""
    import importlib
    ref = reference_import.load()
    from dune_transformercvn_b200 import install
    from dune_transformercvn_b200.loss import fused_training_step
    from dune_transformercvn_b200.sdxl import NeutrinoSDXLNetwork
    base = importlib.import_module("transformercvn.network.trainers.neutrino_full_base_trainer")
    orig_step = base.NeutrinoFullBaseTrainer.training_step
    install.install(sdxl=True, fused_loss=True)
    try:
        sdxl_trainer = importlib.import_module("transformercvn.network.trainers.neutrino_full_sdxl_trainer")
        net = sdxl_trainer.NeutrinoFullSDXLTrainer.create_network(None, ref.tutorial_options(), 1, 1, 3, 8, 4)
        assert isinstance(net, NeutrinoSDXLNetwork)
        assert base.NeutrinoFullBaseTrainer.training_step is fused_training_step
    finally:
        install.uninstall()
    assert base.NeutrinoFullBaseTrainer.training_step is orig_step

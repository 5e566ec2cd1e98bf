"""Host-side logic of the training path on CPU: arena layout, optimizer grouping, and the data-parallel gradient
exchange over a world_size-2 gloo group (the N>1 path of bench.py / training.GradientExchange)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dune_transformercvn_b200 import lib as tl
from dune_transformercvn_b200 import training
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.network import NeutrinoDenseNetwork


@pytest.fixture(scope="module")
def net():
    return NeutrinoDenseNetwork(PathOptions.tutorial(), 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES)


def test_arena_layout_follows_the_state_dict(net):
    a = net.train_engine.arena
    floats = [(k, v) for k, v in net.state_dict().items() if v.is_floating_point()]
    assert [s.name for s in a.specs] == [k for k, _ in floats]
    assert a.total == sum(v.numel() for _, v in floats)
    # sub-arenas handed to the C walks are contiguous and ordered as the C side expects
    order = ["position", "prong", "event", "combined", "encoder", "event_decoder", "prong_decoder"]
    los = [a.seg[t][0] for t in order]
    assert los == sorted(los)
    assert a.seg["prong"][1] == a.seg["event"][0] and a.seg["event"][1] == a.seg["combined"][0]
    assert a.seg["prong_decoder"][1] == a.total
    import ctypes as C
    L = tl.load()
    for tag, width in (("prong", 256), ("event", 288)):
        d = net.engine.cnn_desc(width)
        assert L.tcvn_cnn_arena_floats(C.byref(d)) == a.seg[tag][1] - a.seg[tag][0]
        for prec in (tl.TCVN_FP32, tl.TCVN_BF16):
            assert L.tcvn_cnn_train_workspace_bytes(C.byref(d), prec, 8) > 8 * 64 * 200 * 140 * 4
    sd = net.engine.seq_desc()
    assert L.tcvn_seq_train_workspace_bytes(C.byref(sd), 4, 10, 22) > 0
    assert L.tcvn_seq_train_workspace_bytes(C.byref(sd), 4, 40, 22) == 0      # more slots than the kernels hold


def test_parameters_without_gradient_match_the_reference(net):
    a = net.train_engine.arena
    no_grad = sorted(n for n, _ in net.named_parameters() if not a.has_grad(n))
    assert len(no_grad) == 13                        # SURVEY.md §8 a20 (probe on the reference)
    assert sum(p.numel() for n, p in net.named_parameters() if not a.has_grad(n)) == 856


def test_training_on_cpu_fails_loudly(net):
    with pytest.raises(tl.TcvnError):
        net.train_engine.arena.bind()
    opt = training.TcvnAdamW(training.reference_param_groups(net, 1e-2), lr=1e-3)
    with pytest.raises(tl.TcvnError):
        opt.step()
    groups = training.reference_param_groups(net, 1e-2)
    assert len(groups[0]["params"]) == 464 and len(groups[1]["params"]) == 314


class _FakeSpec:
    def __init__(self, name, numel, is_param):
        self.name, self.numel, self.is_param = name, numel, is_param


class _FakeArena:
    """The part of FlatArena the exchange touches: parameters and BatchNorm buffers interleaved in one flat vector."""

    def __init__(self, rank):
        self.specs = [_FakeSpec("w0", 4, True), _FakeSpec("bn.running_mean", 3, False), _FakeSpec("w1", 2, True),
                      _FakeSpec("bn.running_var", 3, False)]
        self.offset, cur = {}, 0
        for s in self.specs:
            self.offset[s.name] = cur
            cur += s.numel
        self.flat = torch.arange(cur, dtype=torch.float32) + 100.0 * rank


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # what TrainEngine does on its own once torch.distributed is up with > 1 rank ("auto")
        eng = training.TrainEngine.__new__(training.TrainEngine)
        eng.exchange = "auto"
        ex = eng._exchange()
        assert isinstance(ex, training.GradientExchange) and ex.world == world and eng._exchange() is ex
        g = torch.arange(10, dtype=torch.float32) * (rank + 1)
        ex.reduce(g[6:])          # slices are exchanged as soon as they are final, in backward order
        ex.reduce(g[:6])
        ex.wait()
        assert not ex.pending
        arena = _FakeArena(rank)
        ex.broadcast_buffers(arena)      # DDP's broadcast_buffers: rank 0's running statistics, parameters untouched
        torch.save((g, arena.flat), out + f".{rank}")
    finally:
        dist.destroy_process_group()


def test_gradient_exchange_world2_gloo(tmp_path):
    port = 29500 + os.getpid() % 2000
    out = str(tmp_path / "g")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    want = torch.arange(10, dtype=torch.float32) * 1.5     # mean of x1 and x2
    base = torch.arange(12, dtype=torch.float32)
    buf = torch.zeros(12, dtype=torch.bool)
    buf[4:7] = True
    buf[9:12] = True
    for r in range(2):
        g, flat = torch.load(out + f".{r}")
        assert torch.allclose(g, want)
        assert torch.equal(flat[buf], base[buf])                      # buffers: rank 0's values everywhere
        assert torch.equal(flat[~buf], base[~buf] + 100.0 * r)        # parameters: left alone


def test_no_exchange_without_a_process_group():
    eng = training.TrainEngine.__new__(training.TrainEngine)
    eng.exchange = "auto"
    assert eng._exchange() is None and eng.exchange == "auto"

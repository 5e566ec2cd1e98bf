"""Throughput of the --sdxl variant's fp32 path (BASELINE configs[3]) on a bounded sample: events/s, images/s, TFLOP/s."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dune_transformercvn_b200 import synth
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.sdxl import NeutrinoSDXLNetwork
events = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
net = NeutrinoSDXLNetwork(PathOptions.tutorial(), 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES).to(dev).eval()
batch = synth.make_batch(events, seed=1234, max_prongs=10).to(dev)
images = batch.num_events + batch.num_prongs
with torch.no_grad():
    net.forward_sparse(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        net.forward_sparse(batch)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
print(f"sdxl fp32: {events} events / {images} images per step, {ms:.1f} ms -> {events / ms * 1e3:.1f} events/s, "
      f"{images / ms * 1e3:.1f} images/s, {images * 57.2e9 / ms / 1e9:.1f} TFLOP/s (57.2 GFLOP/image, SURVEY 8d)")

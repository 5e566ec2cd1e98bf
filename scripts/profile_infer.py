"""One steady-state inference step of BASELINE configs[1] (256 events, bf16, hit lists consumed by the stem) bracketed by
cudaProfilerStart/Stop, for ncu --profile-from-start off --metrics gpu__time_duration.sum (launch list) or --set full -k."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.network import NeutrinoDenseNetwork

events = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
net = NeutrinoDenseNetwork(PathOptions.tutorial(), 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16").to(dev).eval()
batch = bench.make_inputs(events, 1234).to(dev)
with torch.no_grad():
    for _ in range(3):
        net.forward_sparse(batch)
    net.freeze_packed(True)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    net.forward_sparse(batch)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("images", batch.num_events + batch.num_prongs)

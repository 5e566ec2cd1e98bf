"""One eval forward of the --sdxl variant (bf16 tcgen05 walk) over a few events, bracketed by cudaProfilerStart/Stop:
ncu --profile-from-start off --metrics gpu__time_duration.sum (launch list) or --set full -k regex:umma_gemm_kernel."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from dune_transformercvn_b200.config import NUM_EVENT_CLASSES, NUM_PRONG_CLASSES, PathOptions
from dune_transformercvn_b200.sdxl import NeutrinoSDXLNetwork

events = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda:0")
net = NeutrinoSDXLNetwork(PathOptions.tutorial(), 1, 1, 3, NUM_PRONG_CLASSES, NUM_EVENT_CLASSES, precision="bf16").to(dev).eval()
batch = bench.make_inputs(events, 1234).to(dev)
with torch.no_grad():
    for _ in range(2):
        net.forward_sparse(batch)
    net.freeze_packed(True)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    net.forward_sparse(batch)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("images", batch.num_events + batch.num_prongs)

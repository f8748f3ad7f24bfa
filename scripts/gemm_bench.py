"""Times the tcgen05 NT GEMM alone (M x 256 x 256, fp32 out) with CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from posterior_matching_b200 import _lib
S = torch.cuda.current_stream().cuda_stream
for M in (16384, 131072, 524288):
    A = torch.randn(M, 256, device="cuda").to(torch.bfloat16)
    Bt = (torch.randn(256, 256, device="cuda") / 16).to(torch.bfloat16)
    bias = torch.zeros(256, device="cuda")
    y = torch.empty(M, 256, device="cuda")
    f = lambda: _lib.check(_lib.lib.pmvae_tc_gemm_nt(A.data_ptr(), 256, Bt.data_ptr(), 256, bias.data_ptr(), M, 256, 256, y.data_ptr(), S), "nt")
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"PMVAE_TC_DEBUG={os.environ.get('PMVAE_TC_DEBUG','0')} M={M}: {ms*1e3:.1f} us  {2*M*65536/ms/1e9:.1f} TFLOP/s  "
          f"bytes {(M*512+M*1024)/ms/1e6:.0f} GB/s", flush=True)

"""Forward time of the cat_res64 UNet (B=64) under a few PyTorch-level settings (GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bndm_b200.unet import get_model, count_forward_flops

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64


def bench(model, x, t, graph=True, n=10):
    with torch.no_grad():
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                y = model(x, t, return_dict=False)[0]
        torch.cuda.current_stream().wait_stream(s)
        if graph:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                y = model(x, t, return_dict=False)[0]
            run = g.replay
        else:
            run = lambda: model(x, t, return_dict=False)[0]
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            run()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, y


torch.manual_seed(0)
base = get_model(3, 6, 64).to(dev).eval()
fl = count_forward_flops(base, 64, 64) * B
x = torch.randn(B, 3, 64, 64, device=dev)
t = torch.full((B,), 0.5, device=dev)
ms, y0 = bench(base, x, t)
print(f"fp32 NCHW (conv TF32, matmul fp32)  graph : {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s")
ms, _ = bench(base, x, t, graph=False)
print(f"fp32 NCHW eager                           : {ms:8.3f} ms")
from bndm_b200.fused_unet import fuse_unet
fused = fuse_unet(base)
ms, y = bench(fused, x, t)
print(f"FUSED channels-last + K5 (graph)          : {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s maxdiff {(y - y0).abs().max().item():.2e}")
ms, y = bench(fused, x, t, graph=False)
print(f"FUSED eager                               : {ms:8.3f} ms")
torch.backends.cudnn.benchmark = True
ms, y = bench(fused, x, t)
print(f"FUSED cudnn.benchmark (graph)             : {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s maxdiff {(y - y0).abs().max().item():.2e}")
ms, y = bench(base, x, t)
print(f"fp32 NCHW cudnn.benchmark                 : {ms:8.3f} ms  maxdiff {(y - y0).abs().max().item():.2e}")
m2 = get_model(3, 6, 64).to(dev).eval(); m2.load_state_dict(base.state_dict()); m2 = m2.to(memory_format=torch.channels_last)
ms, y = bench(m2, x.contiguous(memory_format=torch.channels_last), t)
print(f"fp32 channels_last cudnn.benchmark        : {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s maxdiff {(y - y0).abs().max().item():.2e}")
torch.backends.cuda.matmul.allow_tf32 = True
ms, y = bench(m2, x.contiguous(memory_format=torch.channels_last), t)
print(f"  + matmul TF32                           : {ms:8.3f} ms  maxdiff {(y - y0).abs().max().item():.2e}")
torch.backends.cuda.matmul.allow_tf32 = False
m3 = get_model(3, 6, 64).to(dev).eval(); m3.load_state_dict(base.state_dict()); m3 = m3.to(torch.bfloat16)
ms, y = bench(m3, x, t)
print(f"bf16 NCHW                                 : {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s maxdiff {(y - y0).abs().max().item():.2e}")
m3 = m3.to(memory_format=torch.channels_last)
ms, y = bench(m3, x.contiguous(memory_format=torch.channels_last), t)
print(f"bf16 channels_last                        : {ms:8.3f} ms  {fl / ms / 1e9:7.1f} TFLOP/s maxdiff {(y - y0).abs().max().item():.2e}")
print("ref out rms", y0.pow(2).mean().sqrt().item())

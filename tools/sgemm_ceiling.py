"""cuBLAS SGEMM (fp32 SIMT, TF32 off) throughput: the practical FP32 ceiling on this chip."""
import torch
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
torch.set_float32_matmul_precision("highest")
for n in (4096, 8192):
    a = torch.randn(n, n, device='cuda'); b = torch.randn(n, n, device='cuda')
    for _ in range(3): c = a @ b
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): c = a @ b
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("cuBLAS sgemm n=%d: %.2f TFLOP/s" % (n, 2 * n ** 3 / ms / 1e9))

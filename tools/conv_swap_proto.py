"""Validate and time tools/conv_swap_proto.cu (weights as the A operand, TMA-fed 128-byte-swizzled pixel tiles,
N = 256) against torch on the 5x5 layers with C_in % 64 == 0 and C_out % 128 == 0, and print the current
conv_pipe_kernel time for the same layer beside it.

Run on a B200:  python tools/conv_swap_proto.py [batch]"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "critic-vae_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import torch
import torch.nn.functional as F
from cvae_native import binding as L

HERE = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(HERE, "libconv_swap_proto.so")
if not os.path.exists(so):
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-shared", "-Xcompiler", "-fPIC", "-o", so,
                    os.path.join(HERE, "conv_swap_proto.cu")], check=True)
lib = ctypes.CDLL(so)
lib.conv_swap_run.restype = ctypes.c_int
lib.conv_swap_run.argtypes = [ctypes.c_int] * 6 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                              ctypes.POINTER(ctypes.c_float), ctypes.c_int]

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev, bf = "cuda", torch.bfloat16
# name, H, C_in, C_out  (5x5, pad 2)
LAYERS = [("E2f", 16, 64, 128), ("E3f", 8, 128, 256), ("D0f", 4, 256, 128), ("D0g", 4, 128, 256), ("E3g", 8, 256, 128)]
g = torch.Generator(device=dev).manual_seed(0)
for name, H, cin, cout in LAYERS:
    x = torch.randn(B, H, H, cin, device=dev, generator=g).to(bf)                 # NHWC
    w = (torch.randn(cout, cin, 5, 5, device=dev, generator=g) * 0.05)
    bias = torch.randn(cout, device=dev, generator=g) * 0.1
    ksteps = 25 * cin // 16
    wpack = torch.zeros(cout * ksteps * 16, dtype=bf, device=dev)
    job = L.PackJob(kind=L.PACK_FWD5, n=cout, ksteps=ksteps, k_channels=cin, cout=cout, cin=cin, src=w.data_ptr(), src2=None,
                    dst=wpack.data_ptr())
    L.check(L.lib.cvae_pack_weights((L.PackJob * 1)(job), 1, L.stream_ptr()))
    torch.cuda.synchronize()
    out = torch.zeros(B, H, H, cout, dtype=bf, device=dev)
    ms = ctypes.c_float(0)
    rc = lib.conv_swap_run(B, H, H, 5, cin, cout, x.data_ptr(), wpack.data_ptr(), bias.data_ptr(), 1, out.data_ptr(), ctypes.byref(ms), 10)
    ref = F.relu(F.conv2d(x.permute(0, 3, 1, 2).float(), w.to(bf).float(), bias, padding=2)).permute(0, 2, 3, 1)
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    flops = 2.0 * B * H * H * cout * cin * 25
    ok = rc == 0 and err <= 2e-2 * scale
    print(f"{'PASS' if ok else 'FAIL'} {name}: rc={rc} max err {err:.3e} (max |ref| {scale:.2f})  {ms.value * 1e3:7.1f} us  "
          f"{flops / max(ms.value, 1e-9) * 1e-9:7.1f} TF/s", flush=True)

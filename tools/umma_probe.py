"""GPU bring-up probe: checks every tcgen05 layout assumption the conv kernels make.

Run on a B200:  python tools/umma_probe.py   (writes gpurun_out/umma_probe.log)

Each case builds raw shared-memory images, states descriptor fields, and compares the UMMA result
with numpy under hypothesis H1 (sbo = M/N-direction core-matrix stride, lbo = K-direction) and H2
(the two fields swapped).
"""
import ctypes
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "libumma_probe.so"))
lib.probe_run.restype = ctypes.c_int
lib.probe_run.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32] + \
    [ctypes.c_uint32] * 12 + [ctypes.c_void_p]

K_MAJOR, MN_MAJOR = 0, 1


def idesc(n, a_major, b_major, m=128):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_major << 15) | (b_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def bf16_image(nbytes, rng):
    """Random bf16 values in [-1,1) as a uint16 image plus its float32 view."""
    vals = torch.from_numpy(rng.uniform(-1, 1, nbytes // 2).astype(np.float32)).to(torch.bfloat16)
    return vals.view(torch.int16).numpy().view(np.uint16).copy(), vals.float().numpy()


def logical(img_f32, rows, ksteps, major, off, lbo, sbo, kstep):
    """Matrix [rows][16*ksteps] the hardware should read from the image under hypothesis H1."""
    r = np.arange(rows)[:, None]
    k = np.arange(16 * ksteps)[None, :]
    s, kk = k // 16, k % 16
    if major == K_MAJOR:
        addr = off + s * kstep + (kk // 8) * lbo + (kk % 8) * 2 + (r // 8) * sbo + (r % 8) * 16
    else:
        addr = off + s * kstep + (kk // 8) * lbo + (kk % 8) * 16 + (r // 8) * sbo + (r % 8) * 2
    assert addr.max() + 2 <= img_f32.size * 2, (addr.max(), img_f32.size * 2)
    return img_f32[addr // 2]


def run_case(name, n, ksteps, a, b, use_bulk=0, log=print):
    """a / b: dict(bytes, major, off, lbo, sbo, kstep)."""
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    a_u16, a_f = bf16_image(a["bytes"], rng)
    b_u16, b_f = bf16_image(b["bytes"], rng)
    a_dev = torch.from_numpy(a_u16.view(np.int16)).cuda()
    b_dev = torch.from_numpy(b_u16.view(np.int16)).cuda()
    out = torch.zeros(128, n, dtype=torch.float32, device="cuda")
    rc = lib.probe_run(a_dev.data_ptr(), a["bytes"], b_dev.data_ptr(), b["bytes"], a["off"], b["off"],
                       a["lbo"], a["sbo"], b["lbo"], b["sbo"], a["kstep"], b["kstep"], ksteps,
                       idesc(n, a["major"], b["major"]), n, use_bulk, out.data_ptr())
    got = out.cpu().numpy().astype(np.float64)
    res = {}
    for hyp, swap in (("H1", False), ("H2", True)):
        try:
            la = logical(a_f, 128, ksteps, a["major"], a["off"], a["sbo"] if swap else a["lbo"],
                         a["lbo"] if swap else a["sbo"], a["kstep"])
            lb = logical(b_f, n, ksteps, b["major"], b["off"], b["sbo"] if swap else b["lbo"],
                         b["lbo"] if swap else b["sbo"], b["kstep"])
            exp = la.astype(np.float64) @ lb.astype(np.float64).T
            res[hyp] = float(np.abs(got - exp).max())
        except AssertionError:
            res[hyp] = float("nan")
    ok = rc == 0 and res["H1"] < 1e-3
    log(f"{'PASS' if ok else 'FAIL'} {name:<34} rc={rc} n={n} ksteps={ksteps} "
        f"maxerr H1={res['H1']:.3e} H2={res['H2']:.3e} |got|max={np.abs(got).max():.3f}")
    return ok


def main():
    os.makedirs("gpurun_out", exist_ok=True)
    logf = open("gpurun_out/umma_probe.log", "w")

    def log(s):
        print(s)
        logf.write(s + "\n")
        logf.flush()

    log(f"device: {torch.cuda.get_device_name(0)} cc={torch.cuda.get_device_capability(0)}")
    results = []

    def canon_k(rows, ksteps):
        # core matrices ordered [row block][k chunk]; compact
        kch = 2 * ksteps
        return dict(bytes=rows * kch * 16, major=K_MAJOR, off=0, lbo=128, sbo=kch * 128, kstep=256)

    # 1. canonical K-major x K-major at several N
    for n in (16, 32, 64, 128, 256):
        results.append(run_case(f"kmajor_canonical_n{n}", n, 4, canon_k(128, 4), canon_k(n, 4), log=log))
    # 2. same with the B image brought in by cp.async.bulk
    results.append(run_case("kmajor_canonical_bulkB", 64, 4, canon_k(128, 4), canon_k(64, 4), use_bulk=1, log=log))

    # 3. halo planes: A = [q][vpix][8ch] planes, plane stride P = 16 (mod 128), shifted starts
    VP, Q = 220, 8
    P = VP * 16 + 16
    for shift in (0, 1, 3, 8, 11, 70):
        a = dict(bytes=Q * P, major=K_MAJOR, off=shift * 16, lbo=P, sbo=128, kstep=2 * P)
        results.append(run_case(f"halo_kmajor_shift{shift}", 64, 4, a, canon_k(64, 4), log=log))

    # 4. MN-major x MN-major (wgrad): A = dY planes (128 ch), B = X planes (N ch) shifted, K = pixels
    VP = 100
    P = VP * 16 + 16
    for n, shift in ((32, 0), (64, 5), (128, 13), (256, 2)):
        a = dict(bytes=16 * P, major=MN_MAJOR, off=0, lbo=128, sbo=P, kstep=256)
        b = dict(bytes=(n // 8) * P, major=MN_MAJOR, off=shift * 16, lbo=128, sbo=P, kstep=256)
        results.append(run_case(f"wgrad_mnmajor_n{n}_shift{shift}", n, 4, a, b, log=log))

    # 5. shift trick: MN-major A whose 16 M-blocks are 16 one-pixel shifts of one plane (sbo = 16)
    a = dict(bytes=(64 + 32) * 16, major=MN_MAJOR, off=0, lbo=128, sbo=16, kstep=256)
    b = dict(bytes=4 * P, major=MN_MAJOR, off=3 * 16, lbo=128, sbo=P, kstep=256)
    results.append(run_case("wgrad_shift_trick_A", 32, 4, a, b, log=log))
    # same for the B side (N-blocks are shifts)
    a = dict(bytes=16 * P, major=MN_MAJOR, off=0, lbo=128, sbo=P, kstep=256)
    b = dict(bytes=(64 + 32) * 16, major=MN_MAJOR, off=0, lbo=128, sbo=16, kstep=256)
    results.append(run_case("wgrad_shift_trick_B", 48, 4, a, b, log=log))

    # 6. mixed: A K-major halo planes, B MN-major
    VP, Q = 220, 8
    P2 = VP * 16 + 16
    a = dict(bytes=Q * P2, major=K_MAJOR, off=5 * 16, lbo=P2, sbo=128, kstep=2 * P2)
    b = dict(bytes=8 * 1040, major=MN_MAJOR, off=0, lbo=128, sbo=1040, kstep=256)
    results.append(run_case("mixed_kmajorA_mnmajorB", 64, 4, a, b, log=log))

    # 7. many K steps (accumulate flag over a long chain)
    a = canon_k(128, 16)
    b = canon_k(128, 16)
    results.append(run_case("kmajor_long_k256", 128, 16, a, b, log=log))

    log(f"SUMMARY {sum(results)}/{len(results)} passed")
    return 0 if all(results) else 1


if __name__ == "__main__":
    sys.exit(main())

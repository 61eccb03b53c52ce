"""Swizzled-layout probe (round-2 groundwork): can a 128-byte-swizzled operand tile that TMA wrote with the
absolute-address swizzle be read by tcgen05.mma through a descriptor whose start address is shifted by an arbitrary
number of 128-byte rows (the halo / filter-tap trick), for K-major A (conv forward) and MN-major A/B (weight gradient)?
Each case prints the max error under several hypotheses about which bytes the hardware reads.

Run on a B200:  python tools/umma_probe_swz.py   (writes gpurun_out/umma_probe_swz.log)"""
import ctypes, os, sys
import numpy as np, torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "libumma_probe.so"))
lib.probe_run_swz.restype = ctypes.c_int
lib.probe_run_swz.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32] + [ctypes.c_uint32] * 12 + \
    [ctypes.c_void_p] + [ctypes.c_uint32] * 4
K_MAJOR, MN_MAJOR = 0, 1


def idesc(n, a_major, b_major, m=128):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_major << 15) | (b_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def bf16_image(nbytes, rng):
    vals = torch.from_numpy(rng.uniform(-1, 1, nbytes // 2).astype(np.float32)).to(torch.bfloat16)
    return vals.view(torch.int16).numpy().view(np.uint16).copy(), vals.float().numpy()


def addr_matrix(rows, ksteps, op, hyp):
    """Byte address (inside the operand image) of logical element [row r of M/N][k] under hypothesis `hyp`.
    hyp: dict(swap=bool lbo/sbo swapped, xor='none'|'abs'|'rel')."""
    r = np.arange(rows)[:, None]
    k = np.arange(16 * ksteps)[None, :]
    s, kk = k // 16, k % 16
    lbo, sbo = (op["sbo"], op["lbo"]) if hyp["swap"] else (op["lbo"], op["sbo"])
    start = op["off"] + s * op["kstep"]
    if op["major"] == K_MAJOR:
        if op["swz"]:     # 128-byte rows: 64 K elements per row, 8-row atoms; a K = 16 step is 32 bytes inside the row
            rel = (r // 8) * sbo + (r % 8) * 128 + kk * 2
        else:
            rel = (kk // 8) * lbo + (kk % 8) * 2 + (r // 8) * sbo + (r % 8) * 16
    else:
        if op["swz"]:     # 128-byte rows hold 64 M/N elements, rows are K; 8-row atoms 1024 B apart (sbo), 64-element column blocks lbo apart
            rel = (kk // 8) * sbo + (kk % 8) * 128 + (r // 64) * lbo + (r % 64) * 2
        else:
            rel = (kk // 8) * lbo + (kk % 8) * 16 + (r // 8) * sbo + (r % 8) * 2
    lin = start + rel
    if op["swz"] == 64:   # 64-byte swizzle (untested hypothesis, round 2): 16-byte chunk index (2 bits) ^= address bits 7-8
        if op["major"] == K_MAJOR:
            rel = (r // 8) * sbo + (r % 8) * 64 + kk * 2
        else:
            rel = (kk // 8) * sbo + (kk % 8) * 64 + (r // 32) * lbo + (r % 32) * 2
        lin = start + rel
        if hyp["xor"] == "abs":
            lin = lin ^ (((lin >> 7) & 3) << 4)
        elif hyp["xor"] == "rel":
            lin = start + (rel ^ (((rel >> 7) & 3) << 4))
        return lin
    if op["swz"] and hyp["xor"] == "abs":
        lin = lin ^ (((lin >> 7) & 7) << 4)
    elif op["swz"] and hyp["xor"] == "rel":
        lin = start + (rel ^ (((rel >> 7) & 7) << 4))
    return lin


def run_case(name, n, ksteps, a, b, log):
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    a_u16, a_f = bf16_image(a["bytes"], rng)
    b_u16, b_f = bf16_image(b["bytes"], rng)
    a_dev = torch.from_numpy(a_u16.view(np.int16)).cuda()
    b_dev = torch.from_numpy(b_u16.view(np.int16)).cuda()
    out = torch.zeros(128, n, dtype=torch.float32, device="cuda")
    rc = lib.probe_run_swz(a_dev.data_ptr(), a["bytes"], b_dev.data_ptr(), b["bytes"], a["off"], b["off"], a["lbo"], a["sbo"], b["lbo"],
                           b["sbo"], a["kstep"], b["kstep"], ksteps, idesc(n, a["major"], b["major"]), n, 0, out.data_ptr(),
                           {0: 0, 1: 2, 128: 2, 64: 4}[a["swz"]], {0: 0, 1: 2, 128: 2, 64: 4}[b["swz"]], a.get("boff", 0), b.get("boff", 0))
    got = out.cpu().numpy().astype(np.float64)
    res = []
    for swap_a in (False, True):
        for xa in (("abs", "rel") if a["swz"] else ("none",)):
            for swap_b in (False, True):
                for xb in (("abs", "rel") if b["swz"] else ("none",)):
                    try:
                        aa = addr_matrix(128, ksteps, a, dict(swap=swap_a, xor=xa))
                        bb = addr_matrix(n, ksteps, b, dict(swap=swap_b, xor=xb))
                        if aa.max() + 2 > a["bytes"] or bb.max() + 2 > b["bytes"] or aa.min() < 0 or bb.min() < 0:
                            continue
                        exp = a_f[aa // 2].astype(np.float64) @ b_f[bb // 2].astype(np.float64).T
                        res.append((float(np.abs(got - exp).max()), f"A[{'swap' if swap_a else 'std'},{xa}] B[{'swap' if swap_b else 'std'},{xb}]"))
                    except Exception as e:   # noqa
                        pass
    res.sort()
    best = res[0] if res else (float("nan"), "none")
    ok = rc == 0 and best[0] < 1e-3
    log(f"{'PASS' if ok else 'FAIL'} {name:<44} rc={rc} best {best[1]} err={best[0]:.3e}" +
        ("" if len(res) < 2 else f" | next {res[1][1]} err={res[1][0]:.3e}"))
    return ok, best[1]


def main():
    os.makedirs("gpurun_out", exist_ok=True)
    logf = open("gpurun_out/umma_probe_swz.log", "w")

    def log(s):
        print(s); logf.write(s + "\n"); logf.flush()

    log(f"device: {torch.cuda.get_device_name(0)}")
    canon_k = lambda rows, ksteps: dict(bytes=rows * 2 * ksteps * 16, major=K_MAJOR, off=0, lbo=128, sbo=2 * ksteps * 128, kstep=256, swz=0)
    ROWS = 128 + 32
    # 1. K-major 128B-swizzled A, rows = pixels (128 B = 64 channels), start shifted by `shift` rows; B canonical no-swizzle
    for shift in (0, 8, 16, 1, 3, 5, 11):
        for boff in sorted({0, shift % 8}):
            a = dict(bytes=ROWS * 128, major=K_MAJOR, off=shift * 128, lbo=16, sbo=1024, kstep=32, swz=1, boff=boff)
            run_case(f"kmajor_sw128_A_shift{shift}_boff{boff}", 64, 4, a, canon_k(64, 4), log)
    # 2. MN-major 128B-swizzled A (weight gradient dY^T): rows = K (pixels), 128 B = 64 channels, two column blocks for M = 128
    KR = 64 + 32
    blk = KR * 128
    for shift in (0, 8, 1, 3, 13):
        for boff in sorted({0, shift % 8}):
            a = dict(bytes=2 * blk, major=MN_MAJOR, off=shift * 128, lbo=blk, sbo=1024, kstep=2048, swz=1, boff=boff)
            bq = dict(bytes=(64 // 8) * (KR * 16 + 16), major=MN_MAJOR, off=0, lbo=128, sbo=KR * 16 + 16, kstep=256, swz=0)
            run_case(f"mnmajor_sw128_A_shift{shift}_boff{boff}", 64, 4, a, bq, log)
    # 3. both MN-major swizzled, B shifted (the filter tap), N = 64 and 128
    for n in (64, 128):
        for shift in (0, 8, 3, 13):
            for boff in sorted({0, shift % 8}):
                a = dict(bytes=2 * blk, major=MN_MAJOR, off=0, lbo=blk, sbo=1024, kstep=2048, swz=1, boff=0)
                b = dict(bytes=(n // 64) * blk, major=MN_MAJOR, off=shift * 128, lbo=blk, sbo=1024, kstep=2048, swz=1, boff=boff)
                run_case(f"mnmajor_sw128_AB_n{n}_Bshift{shift}_boff{boff}", n, 4, a, b, log)
    # 4. (round-2 groundwork, not yet run) 64-byte swizzle for the 32-channel operands: MN-major, 64-byte rows = 32
    #    channels, 8-row atoms of 512 B (SBO), four column blocks for M = 128, a K = 16 step is 1024 B; B shifted by rows
    KR = 64 + 32
    blk64 = KR * 64
    for shift in (0, 8, 3, 13):
        a = dict(bytes=4 * blk64, major=MN_MAJOR, off=0, lbo=blk64, sbo=512, kstep=1024, swz=64, boff=0)
        b = dict(bytes=2 * blk64, major=MN_MAJOR, off=shift * 64, lbo=blk64, sbo=512, kstep=1024, swz=64, boff=0)
        run_case(f"mnmajor_sw64_AB_n64_Bshift{shift}", 64, 4, a, b, log)
    # 5. (round-2 groundwork, not yet run) conv forward with the weights as A: A = canonical no-swizzle K-major
    #    (the packed weight blocks), B = K-major 128B-swizzled pixel tile, N = 256 pixels, start shifted by `shift` rows
    for n in (128, 256):
        for shift in (0, 3, 11):
            b = dict(bytes=(n + 32) * 128, major=K_MAJOR, off=shift * 128, lbo=16, sbo=1024, kstep=32, swz=1, boff=0)
            run_case(f"weightsA_kmajor_sw128_B_n{n}_shift{shift}", n, 4, canon_k(128, 4), b, log)
    # 6. the same with 8-row atoms that are NOT contiguous (SBO = 1280: one 8-pixel image row per atom, rows PW = 10 pixels apart)
    for shift in (0, 1, 11):
        b = dict(bytes=32 * 1280 + 64 * 128, major=K_MAJOR, off=shift * 128, lbo=16, sbo=1280, kstep=32, swz=1, boff=0)
        run_case(f"weightsA_kmajor_sw128_B_n256_sbo1280_shift{shift}", 256, 4, canon_k(128, 4), b, log)
    # 7. 64-byte-swizzled K-major pixel tile (32-channel sources: 64-byte rows, 8-row atoms of 512 B), K = 32 = two steps of 32 B
    for n in (128, 256):
        for shift in (0, 3, 11):
            b = dict(bytes=(n + 32) * 64, major=K_MAJOR, off=shift * 64, lbo=16, sbo=512, kstep=32, swz=64, boff=0)
            run_case(f"weightsA_kmajor_sw64_B_n{n}_shift{shift}", n, 2, canon_k(128, 2), b, log)
    return 0


if __name__ == "__main__":
    sys.exit(main())

// Halo-plane loaders shared by the conv GEMM and weight-gradient kernels.
//
// A "plane" is [virtual pixel][8 channels] bf16 (16 bytes per pixel), the no-swizzle UMMA core-matrix
// layout with pixels as rows.  Virtual pixel v of a (B, H, W, pad) grid decodes as
//   vrow = v / PW, vcol = v % PW (PW = W + pad), image n = vrow / IH, r = vrow % IH (IH = H + pad);
// it is a real pixel (n, r - pad, vcol) iff vcol < W, r >= pad, n < B and v >= 0 -- everything else
// is the shared zero padding of the convolution.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace cvae {

struct PlaneSrc {
    int B, H, W, pad, PW, IH;
    int planes;   // 8-channel planes produced per pixel
    int src_c;    // channel count of the raw source tensor (NHWC / S2D modes)
    int ones;     // NCHW3 mode: write 1.0 into channel 3 of every pixel slot (bias-gradient column)
    const void* src;
    const void* src2;
};

// Division by a run-time constant via multiply-high (exact for 0 <= n < 2^31).
struct FastDiv {
    uint32_t d, mul, shift;
};
static inline FastDiv make_fastdiv(int d) {
    FastDiv f{(uint32_t)d, 0u, 0u};
    if (d > 1) {
        int lg = 0;
        while ((1LL << lg) < d) ++lg;
        const int p = 31 + lg;
        f.mul = (uint32_t)(((1ULL << p) + (uint64_t)d - 1) / (uint64_t)d);
        f.shift = (uint32_t)(p - 32);
    }
    return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t n, const FastDiv& f) {
    return f.d == 1 ? n : (__umulhi(n, f.mul) >> f.shift);
}

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ void cp_async_commit() {
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Asynchronous plane fill (cp.async, 16 bytes per copy, zero-fill for padding) for the bf16 loaders.
// Planes [q_first, q_first + nplanes) of the source are written to planes 0.. of the buffer.
// Lanes are grouped so that `g` consecutive lanes copy consecutive 16-byte chunks of one pixel
// (coalesced global reads); every lane decodes its pixel once per iteration.
template <int LOADER>
__device__ __forceinline__ void fill_planes_async(const PlaneSrc& a, const FastDiv& dPW, const FastDiv& dIH, uint8_t* planes,
                                                  int plane_stride, int v_first, int count, int q_first, int nplanes,
                                                  int tid, int nthreads) {
    static_assert(LOADER == CVAE_LOAD_NHWC || LOADER == CVAE_LOAD_S2D, "cp.async fill needs a bf16 source");
    int g = 1;
    while (g < 8 && g * 2 <= nplanes) g <<= 1;          // lanes per pixel: 1, 2, 4 or 8
    const int sub = tid & (g - 1);
    const int workers = nthreads / g;                    // pixel walkers
    const int wid = tid / g;
    // every walker owns a contiguous run of pixel slots and decodes (image, row, column) once, then steps
    const int per = (count + workers - 1) / workers;
    int j = wid * per;
    const int j_end = min(j + per, count);
    if (j >= j_end) return;
    const uint32_t base = smem_u32(planes);
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(a.src);
    const int cq = a.src_c >> 3;                         // S2D: 16-byte chunks per raw pixel
    int v = v_first + j;
    int n, r, vcol;
    if (v < 0) {                                         // before the first image: pure padding
        n = -1; r = 0; vcol = 0;
    } else {
        const uint32_t vrow = fast_div((uint32_t)v, dPW);
        vcol = (int)((uint32_t)v - vrow * (uint32_t)a.PW);
        n = (int)fast_div(vrow, dIH);
        r = (int)(vrow - (uint32_t)n * (uint32_t)a.IH);
    }
    for (; j < j_end; ++j, ++v) {
        if (v == 0) { n = 0; r = 0; vcol = 0; }
        const bool valid = (v >= 0) && (vcol < a.W) && (r >= a.pad) && (n < a.B);
        const int h = r - a.pad, w = vcol;
        const uint32_t dst = base + (uint32_t)j * 16u;
        const uint32_t nbytes = valid ? 16u : 0u;
        if constexpr (LOADER == CVAE_LOAD_NHWC) {
            const __nv_bfloat16* s = valid ? src + (((size_t)n * a.H + h) * a.W + w) * a.src_c + (size_t)q_first * 8 : src;
            for (int q = sub; q < nplanes; q += g)
                cp_async16(dst + (uint32_t)q * (uint32_t)plane_stride, valid ? s + q * 8 : src, nbytes);
        } else {
            // source [B][2H][2W][C]; plane q <-> (phase ab = q / (C/8), channel chunk q % (C/8))
            const size_t pix00 = valid ? (((size_t)n * 2 * a.H + 2 * h) * (2 * a.W) + 2 * w) : 0;
            for (int q = sub; q < nplanes; q += g) {
                const int qq = q_first + q, ab = qq / cq, cc = qq - ab * cq;
                const size_t pix = pix00 + (size_t)(ab >> 1) * (2 * a.W) + (ab & 1);
                cp_async16(dst + (uint32_t)q * (uint32_t)plane_stride, valid ? src + pix * a.src_c + cc * 8 : src, nbytes);
            }
        }
        if (v >= 0 && ++vcol == a.PW) {
            vcol = 0;
            if (++r == a.IH) { r = 0; ++n; }
        }
    }
}

// Synchronous plane fill for the two fp32 NCHW sources (the threads convert to bf16 themselves); the bf16 sources go
// through fill_planes_async.
template <int LOADER>
__device__ __forceinline__ void fill_planes(const PlaneSrc& a, const FastDiv& dPW, const FastDiv& dIH, uint8_t* planes, int plane_stride,
                                            int v_first, int count, int tid, int nthreads) {
    static_assert(LOADER == CVAE_LOAD_NCHW3 || LOADER == CVAE_LOAD_S2D_NCHW3_DTANH, "bf16 sources use fill_planes_async");
    if constexpr (LOADER == CVAE_LOAD_NCHW3) {
        // Batches of kBatch pixels per thread: all 3 * kBatch loads are issued before the first value is used, so a
        // batch costs one global-memory latency instead of kBatch (the producer was the limiter of encoder conv 0:
        // ~960 cycles per pixel pair with the loads two deep, profiles/r02_e0f_producer.log).
        constexpr int kBatch = 8;
        const size_t cs = (size_t)a.H * a.W;
        const float c3 = a.ones ? 1.f : 0.f;
        for (int j0 = tid; j0 < count; j0 += kBatch * nthreads) {
            float f[kBatch][3];
            bool ok[kBatch];
#pragma unroll
            for (int q = 0; q < kBatch; ++q) {
                const int j = j0 + q * nthreads;
                const int v = v_first + j;
                ok[q] = false;
                f[q][0] = f[q][1] = f[q][2] = 0.f;
                if (j < count && v >= 0) {
                    const uint32_t vrow = fast_div((uint32_t)v, dPW);
                    const int vcol = (int)((uint32_t)v - vrow * (uint32_t)a.PW);
                    const int n = (int)fast_div(vrow, dIH);
                    const int r = (int)(vrow - (uint32_t)n * (uint32_t)a.IH);
                    if ((vcol < a.W) && (r >= a.pad) && (n < a.B)) {
                        const float* s = reinterpret_cast<const float*>(a.src) + ((size_t)n * 3 * a.H + (r - a.pad)) * a.W + vcol;
                        f[q][0] = __ldg(s);
                        f[q][1] = __ldg(s + cs);
                        f[q][2] = __ldg(s + 2 * cs);
                    }
                }
                ok[q] = j < count;
            }
#pragma unroll
            for (int q = 0; q < kBatch; ++q)
                if (ok[q])
                    *reinterpret_cast<uint4*>(planes + (size_t)(j0 + q * nthreads) * 16) =
                        make_uint4(pack_bf16x2(f[q][0], f[q][1]), pack_bf16x2(f[q][2], c3), 0u, 0u);
        }
        return;
    }
#pragma unroll 2
    for (int j = tid; j < count; j += nthreads) {
        const int v = v_first + j;
        int vrow = v / a.PW;
        int vcol = v - vrow * a.PW;
        int n = vrow / a.IH;
        int r = vrow - n * a.IH;
        const bool valid = (v >= 0) && (vcol < a.W) && (r >= a.pad) && (n < a.B);
        const int h = r - a.pad, w = vcol;
        uint8_t* dst = planes + (size_t)j * 16;
        if constexpr (LOADER == CVAE_LOAD_NCHW3) {
        } else {  // CVAE_LOAD_S2D_NCHW3_DTANH: 12 channels (a,b,c) + 4 zeros, two planes
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = 0.f;
            if (valid) {
                const int H2 = 2 * a.H, W2 = 2 * a.W;
                const float* g = reinterpret_cast<const float*>(a.src);
                const float* rc = reinterpret_cast<const float*>(a.src2);
                // the two horizontal phases of a row are adjacent floats (2w is even): one 8-byte load each
                float2 gv[6], rv[6];
#pragma unroll
                for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const size_t idx = (((size_t)n * 3 + c) * H2 + 2 * h + a2) * W2 + 2 * w;
                        gv[a2 * 3 + c] = __ldg(reinterpret_cast<const float2*>(g + idx));
                        rv[a2 * 3 + c] = __ldg(reinterpret_cast<const float2*>(rc + idx));
                    }
#pragma unroll
                for (int a2 = 0; a2 < 2; ++a2)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const float2 gg = gv[a2 * 3 + c], rr = rv[a2 * 3 + c];
                        f[(a2 * 2 + 0) * 3 + c] = gg.x * (1.f - rr.x * rr.x);
                        f[(a2 * 2 + 1) * 3 + c] = gg.y * (1.f - rr.y * rr.y);
                    }
            }
            uint4 p0, p1;
            p0.x = pack_bf16x2(f[0], f[1]);   p0.y = pack_bf16x2(f[2], f[3]);
            p0.z = pack_bf16x2(f[4], f[5]);   p0.w = pack_bf16x2(f[6], f[7]);
            p1.x = pack_bf16x2(f[8], f[9]);   p1.y = pack_bf16x2(f[10], f[11]);
            p1.z = pack_bf16x2(f[12], f[13]); p1.w = pack_bf16x2(f[14], f[15]);
            if (a.planes > 0) *reinterpret_cast<uint4*>(dst) = p0;
            if (a.planes > 1) *reinterpret_cast<uint4*>(dst + plane_stride) = p1;
        }
    }
}

// Fill `count` pixel slots of one plane with bf16 1.0 (the ones-operand of the bias-gradient trick).
__device__ __forceinline__ void fill_ones_plane(uint8_t* plane, int count, int tid, int nthreads) {
    const uint32_t one2 = 0x3F803F80u;
    for (int j = tid; j < count; j += nthreads)
        *reinterpret_cast<uint4*>(plane + (size_t)j * 16) = make_uint4(one2, one2, one2, one2);
}

}  // namespace cvae

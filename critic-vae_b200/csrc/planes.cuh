// Halo-plane loaders shared by the conv GEMM and weight-gradient kernels.
//
// A "plane" is [virtual pixel][8 channels] bf16 (16 bytes per pixel), the no-swizzle UMMA core-matrix
// layout with pixels as rows.  Virtual pixel v of a (B, H, W, pad) grid decodes as
//   vrow = v / PW, vcol = v % PW (PW = W + pad), image n = vrow / IH, r = vrow % IH (IH = H + pad);
// it is a real pixel (n, r - pad, vcol) iff vcol < W, r >= pad, n < B and v >= 0 -- everything else
// is the shared zero padding of the convolution.
#pragma once
#include "common.cuh"

namespace cvae {

struct PlaneSrc {
    int B, H, W, pad, PW, IH;
    int planes;   // 8-channel planes produced per pixel
    int src_c;    // channel count of the raw source tensor (NHWC / S2D modes)
    int ones;     // NCHW3 mode: write 1.0 into channel 3 of every pixel slot (bias-gradient column)
    const void* src;
    const void* src2;
};

template <int LOADER>
__device__ __forceinline__ void fill_planes(const PlaneSrc& a, uint8_t* planes, int plane_stride,
                                            int v_first, int count, int tid, int nthreads) {
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
        if constexpr (LOADER == CVAE_LOAD_NHWC) {
            const uint4* s = reinterpret_cast<const uint4*>(
                reinterpret_cast<const __nv_bfloat16*>(a.src) +
                ((size_t)(n * a.H + h) * a.W + w) * a.src_c);
            // issue up to 8 independent 16-byte loads before the first store (memory-level parallelism)
            for (int q0 = 0; q0 < a.planes; q0 += 8) {
                uint4 val[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    val[i] = (valid && q0 + i < a.planes) ? __ldg(s + q0 + i) : make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (q0 + i < a.planes) *reinterpret_cast<uint4*>(dst + (size_t)(q0 + i) * plane_stride) = val[i];
            }
        } else if constexpr (LOADER == CVAE_LOAD_S2D) {
            // source [B][2H][2W][C]; plane q <-> (phase ab = q / (C/8), channel chunk q % (C/8))
            const int cq = a.src_c >> 3;
            const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(a.src);
            for (int q0 = 0; q0 < a.planes; q0 += 8) {
                uint4 val[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int q = q0 + i, ab = q / cq, cc = q - ab * cq;
                    val[i] = make_uint4(0, 0, 0, 0);
                    if (valid && q < a.planes) {
                        const size_t pix = ((size_t)(n * 2 * a.H + 2 * h + (ab >> 1)) * (2 * a.W) + 2 * w + (ab & 1));
                        val[i] = __ldg(reinterpret_cast<const uint4*>(base + pix * a.src_c) + cc);
                    }
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (q0 + i < a.planes) *reinterpret_cast<uint4*>(dst + (size_t)(q0 + i) * plane_stride) = val[i];
            }
        } else if constexpr (LOADER == CVAE_LOAD_NCHW3) {
            uint4 val = make_uint4(0, 0, 0, 0);
            float c3 = a.ones ? 1.f : 0.f;
            if (valid) {
                const float* s = reinterpret_cast<const float*>(a.src) + ((size_t)n * 3 * a.H + h) * a.W + w;
                const size_t cs = (size_t)a.H * a.W;
                val.x = pack_bf16x2(__ldg(s), __ldg(s + cs));
                val.y = pack_bf16x2(__ldg(s + 2 * cs), c3);
            } else {
                val.y = pack_bf16x2(0.f, c3);
            }
            *reinterpret_cast<uint4*>(dst) = val;
        } else {  // CVAE_LOAD_S2D_NCHW3_DTANH: 12 channels (a,b,c) + 4 zeros, two planes
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = 0.f;
            if (valid) {
                const int H2 = 2 * a.H, W2 = 2 * a.W;
                const float* g = reinterpret_cast<const float*>(a.src);
                const float* rc = reinterpret_cast<const float*>(a.src2);
#pragma unroll
                for (int ab = 0; ab < 4; ++ab)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const size_t idx = (((size_t)n * 3 + c) * H2 + 2 * h + (ab >> 1)) * W2 + 2 * w + (ab & 1);
                        const float rv = __ldg(rc + idx);
                        f[ab * 3 + c] = __ldg(g + idx) * (1.f - rv * rv);
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

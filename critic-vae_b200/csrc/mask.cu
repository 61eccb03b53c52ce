// The -video mask pipeline (vae_utility.py:256-284,148-160,56-68) on the device.
//
//   diff_grey : |recon(0) - recon(pred)| (fp32, as numpy does) -> grey-scale in fp64, per-frame max.
//   mask_iou  : clamp to mean_max, * diff_factor, * 255, truncate to uint8 (all fp64, no FMA so the
//               bits match numpy), threshold, and the gt x value joint histogram from which tp/fn/fp
//               for EVERY threshold follow -- one pass instead of the reference's 13 re-runs
//               (vae.py:121-123).  One frame per CTA, the frame staged in shared memory.
// mean_max / diff_factor stay on the host (statistics.mean is exact-rational, vae_utility.py:106-110).
#include "common.cuh"

namespace cvae {

__global__ void diff_grey_kernel(int frames, const float* __restrict__ hi, const float* __restrict__ lo,
                                 double* __restrict__ diff, double* __restrict__ maxv) {
    grid_dependency_sync();
    __shared__ double red[8];
    const int f = blockIdx.x;
    const float* h = hi + (size_t)f * 3 * 4096;
    const float* l = lo + (size_t)f * 3 * 4096;
    double m = 0.0;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
        // np.subtract / abs run in fp32, np.dot promotes to fp64: ((r*wr + g*wg) + b*wb)
        const double r = (double)fabsf(__fsub_rn(l[i], h[i]));
        const double g = (double)fabsf(__fsub_rn(l[4096 + i], h[4096 + i]));
        const double b = (double)fabsf(__fsub_rn(l[8192 + i], h[8192 + i]));
        const double v = __dadd_rn(__dadd_rn(__dmul_rn(r, 0.2989), __dmul_rn(g, 0.5870)), __dmul_rn(b, 0.1140));
        diff[(size_t)f * 4096 + i] = v;
        m = fmax(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmax(m, red[i]);
        maxv[f] = m;
    }
}

__global__ void mask_iou_kernel(int frames, const double* __restrict__ diff, const uint8_t* __restrict__ gt,
                                double mean_max, double factor, int thr, uint8_t* __restrict__ diff_u8,
                                uint8_t* __restrict__ mask, unsigned long long* __restrict__ hist) {
    grid_dependency_sync();
    __shared__ unsigned int sh[512];            // [gt][value]
    __shared__ __align__(16) uint8_t q[4096];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (int f = blockIdx.x; f < frames; f += gridDim.x) {
        const double* d = diff + (size_t)f * 4096;
        const uint8_t* g = gt + (size_t)f * 4096;
        // 16 elements per thread: all eight 16-byte loads of the frame slice are issued before the first use
        double2 dv[8];
        uchar2 gv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i2 = k * 256 + threadIdx.x;   // pair index
            dv[k] = __ldg(reinterpret_cast<const double2*>(d) + i2);
            gv[k] = __ldg(reinterpret_cast<const uchar2*>(g) + i2);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i2 = k * 256 + threadIdx.x;
            double v0 = dv[k].x, v1 = dv[k].y;
            if (v0 > mean_max) v0 = mean_max;                     // prepare_diff, vae_utility.py:280
            if (v1 > mean_max) v1 = mean_max;
            v0 = __dmul_rn(__dmul_rn(v0, factor), 255.0);         // :281 then :155
            v1 = __dmul_rn(__dmul_rn(v1, factor), 255.0);
            const uint8_t u0 = (uint8_t)(int)v0, u1 = (uint8_t)(int)v1;   // astype(np.uint8): truncation
            reinterpret_cast<uchar2*>(q)[i2] = make_uchar2(u0, u1);
            atomicAdd(&sh[(gv[k].x ? 256 : 0) + u0], 1u);
            atomicAdd(&sh[(gv[k].y ? 256 : 0) + u1], 1u);
        }
        __syncthreads();
        // coalesced 16-byte stores of the staged frame
        for (int i = threadIdx.x; i < 256; i += blockDim.x) {
            const uint4 v = reinterpret_cast<const uint4*>(q)[i];
            if (diff_u8) reinterpret_cast<uint4*>(diff_u8 + (size_t)f * 4096)[i] = v;
            if (mask) {
                uint4 m;
                const uint32_t t = (uint32_t)thr;
                auto thr4 = [t](uint32_t w) {
                    uint32_t r = 0;
#pragma unroll
                    for (int b = 0; b < 4; ++b) r |= (((w >> (8 * b)) & 0xFFu) > t ? 1u : 0u) << (8 * b);
                    return r;
                };
                m.x = thr4(v.x); m.y = thr4(v.y); m.z = thr4(v.z); m.w = thr4(v.w);
                reinterpret_cast<uint4*>(mask + (size_t)f * 4096)[i] = m;
            }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 512; i += blockDim.x)
        if (sh[i]) atomicAdd(hist + i, (unsigned long long)sh[i]);
}

// counts[t] = (tp, fn, fp) for threshold thr[t]: T = value > thr (vae_utility.py:157,57-59)
__global__ void iou_counts_kernel(const unsigned long long* __restrict__ hist, int nthr, const int* __restrict__ thr,
                                  long long* __restrict__ counts) {
    grid_dependency_sync();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthr) return;
    long long tp = 0, fn = 0, fp = 0;
    for (int v = 0; v < 256; ++v) {
        const long long g1 = (long long)hist[256 + v], g0 = (long long)hist[v];
        if (v > thr[t]) { tp += g1; fp += g0; } else { fn += g1; }
    }
    counts[3 * t] = tp; counts[3 * t + 1] = fn; counts[3 * t + 2] = fp;
}

// tp / fn / fp of two boolean arrays of any length (vae_utility.py:57-59)
__global__ void iou_pair_kernel(long long n, const uint8_t* __restrict__ g, const uint8_t* __restrict__ t,
                                unsigned long long* __restrict__ counts) {
    grid_dependency_sync();
    unsigned int tp = 0, fn = 0, fp = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const bool a = g[i] != 0, b = t[i] != 0;
        tp += a && b; fn += a && !b; fp += !a && b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tp += __shfl_xor_sync(0xffffffffu, tp, o);
        fn += __shfl_xor_sync(0xffffffffu, fn, o);
        fp += __shfl_xor_sync(0xffffffffu, fp, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (tp) atomicAdd(counts, (unsigned long long)tp);
        if (fn) atomicAdd(counts + 1, (unsigned long long)fn);
        if (fp) atomicAdd(counts + 2, (unsigned long long)fp);
    }
}

}  // namespace cvae

using namespace cvae;

extern "C" int cvae_iou_counts(int64_t n, const uint8_t* gt, const uint8_t* mask, int64_t* counts3, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(n >= 0 && counts3 && (n == 0 || (gt && mask)), CVAE_EINVAL, "iou_counts: bad argument");
    CVAE_CUDA(cudaMemsetAsync(counts3, 0, 3 * sizeof(int64_t), stream));
    if (n == 0) return CVAE_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    cvae::launch(iou_pair_kernel, (int)blocks, 256, 0, stream, n, gt, mask, (unsigned long long*)counts3);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_diff_grey(int frames, const float* recon_hi, const float* recon_lo, double* diff,
                              double* max_values, void* stream) {
    CVAE_REQUIRE(frames >= 0 && (frames == 0 || (recon_hi && recon_lo && diff && max_values)), CVAE_EINVAL, "diff_grey: bad argument");
    if (frames == 0) return CVAE_OK;
    cvae::launch(diff_grey_kernel, frames, 256, 0, (cudaStream_t)stream, frames, recon_hi, recon_lo, diff, max_values);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_mask_iou(int frames, const double* diff, const uint8_t* gt, double mean_max, double diff_factor,
                             int thr, int nthr, const int* thr_list, uint8_t* diff_u8, uint8_t* mask,
                             uint64_t* hist512, int64_t* counts, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(frames >= 0 && hist512 && (frames == 0 || (diff && gt)), CVAE_EINVAL, "mask_iou: bad argument");
    CVAE_REQUIRE(nthr == 0 || (thr_list && counts), CVAE_EINVAL, "mask_iou: threshold list");
    CVAE_CUDA(cudaMemsetAsync(hist512, 0, sizeof(uint64_t) * 512, stream));
    if (frames > 0) {
        int grid = frames < sm_count() * 6 ? frames : sm_count() * 6;
        cvae::launch(mask_iou_kernel, grid, 256, 0, stream, frames, diff, gt, mean_max, diff_factor, thr, diff_u8, mask,
                                                  (unsigned long long*)hist512);
        CVAE_LAUNCH_CHECK();
    }
    if (nthr > 0) {
        cvae::launch(iou_counts_kernel, (nthr + 63) / 64, 64, 0, stream, (const unsigned long long*)hist512, nthr, thr_list, (long long*)counts);
        CVAE_LAUNCH_CHECK();
    }
    return CVAE_OK;
}

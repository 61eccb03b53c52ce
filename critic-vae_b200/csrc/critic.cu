// Frozen critic forward (critic_net.py:15-42, evaluate at :66-69): 4 x [conv3x3 pad 1 -> ReLU ->
// MaxPool2] -> conv4x4 -> ReLU -> Linear(32,32) -> ReLU -> Linear(32,1) -> Sigmoid (dropout is
// identity in eval mode).  3.4 MFLOP per frame with 8..32 channels: too narrow for UMMA tiles, so
// one CTA per frame keeps every intermediate map in shared memory and runs fp32 on the CUDA cores
// (fp32 also keeps the value bit-close to the reference; it conditions the decoder).
//
// weights: the 14 state_dict tensors concatenated in key order (fp32), 11,873 floats.
#include "common.cuh"

namespace cvae {

static constexpr int W0 = 0, B0 = W0 + 8 * 3 * 9, W1 = B0 + 8, B1 = W1 + 8 * 8 * 9, W2 = B1 + 8,
                     B2 = W2 + 8 * 8 * 9, W3 = B2 + 8, B3 = W3 + 16 * 8 * 9, W4 = B3 + 16,
                     B4 = W4 + 32 * 16 * 16, F1W = B4 + 32, F1B = F1W + 32 * 32, F2W = F1B + 32,
                     F2B = F2W + 32, kCriticFloats = F2B + 1;
static_assert(kCriticFloats == 11873, "critic parameter count");

// conv3x3(pad 1) + ReLU + MaxPool2 on a CIN x S x S map in shared memory -> COUT x S/2 x S/2.
// One task = one pooled pixel x OC output channels (4 conv positions x OC channels in registers).  OC shrinks with the
// map so that every layer has >= 128 tasks for the 512 threads: the serial FMA chain of a task, not the FMA rate, was what
// the small layers cost (8 channels per task left 64 and 32 threads busy in layers 3 and 4).  The accumulation order of
// each output (ci, ky, kx) does not depend on OC: results are bit-identical.
template <int CIN, int COUT, int S, int OC>
__device__ __forceinline__ void conv_relu_pool(const float* __restrict__ in, const float* __restrict__ w,
                                               const float* __restrict__ b, float* __restrict__ out) {
    constexpr int P = S / 2, G = COUT / OC;
    for (int t = threadIdx.x; t < P * P * G; t += blockDim.x) {
        const int px = t % P, py = (t / P) % P, g = t / (P * P);
        float acc[4][OC];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int o = 0; o < OC; ++o) acc[q][o] = 0.f;
        for (int ci = 0; ci < CIN; ++ci) {
            float patch[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int y = 2 * py - 1 + r, x = 2 * px - 1 + c;
                    patch[r][c] = (y >= 0 && y < S && x >= 0 && x < S) ? in[(ci * S + y) * S + x] : 0.f;
                }
#pragma unroll
            for (int o = 0; o < OC; ++o) {
                const float* wk = w + ((g * OC + o) * CIN + ci) * 9;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float wv = wk[ky * 3 + kx];
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            acc[q][o] = fmaf(wv, patch[(q >> 1) + ky][(q & 1) + kx], acc[q][o]);
                    }
            }
        }
#pragma unroll
        for (int o = 0; o < OC; ++o) {
            const float bb = b[g * OC + o];
            float m = fmaxf(fmaxf(acc[0][o], acc[1][o]), fmaxf(acc[2][o], acc[3][o])) + bb;
            out[((g * OC + o) * P + py) * P + px] = fmaxf(m, 0.f);   // relu(max(.)+b) == max(relu(.+b))
        }
    }
}

static constexpr int kCriticThreads = 512;
__global__ void __launch_bounds__(kCriticThreads, 1)
critic_fwd_kernel(int frames, const float* __restrict__ x, const float* __restrict__ weights, float* __restrict__ pred) {
    grid_dependency_sync();
    extern __shared__ float sm[];
    float* wsm = sm;                          // 11873 (+3 pad)
    float* in = wsm + 11876;                  // 3 x 64 x 64
    float* p1 = in + 3 * 4096;                // 8 x 32 x 32
    float* p2 = p1 + 8 * 1024;                // 8 x 16 x 16
    float* p3 = p2 + 8 * 256;                 // 8 x 8 x 8
    float* p4 = p3 + 8 * 64;                  // 16 x 4 x 4
    float* v5 = p4 + 16 * 16;                 // 32
    float* v6 = v5 + 32;                      // 32
    for (int i = threadIdx.x; i < kCriticFloats; i += blockDim.x) wsm[i] = __ldg(weights + i);
    for (int f = blockIdx.x; f < frames; f += gridDim.x) {
        __syncthreads();
        const float4* src = reinterpret_cast<const float4*>(x + (size_t)f * 3 * 4096);
        for (int i = threadIdx.x; i < 3 * 1024; i += blockDim.x) reinterpret_cast<float4*>(in)[i] = __ldg(src + i);
        __syncthreads();
        conv_relu_pool<3, 8, 64, 8>(in, wsm + W0, wsm + B0, p1);
        __syncthreads();
        conv_relu_pool<8, 8, 32, 4>(p1, wsm + W1, wsm + B1, p2);
        __syncthreads();
        conv_relu_pool<8, 8, 16, 2>(p2, wsm + W2, wsm + B2, p3);
        __syncthreads();
        conv_relu_pool<8, 16, 8, 2>(p3, wsm + W3, wsm + B3, p4);
        __syncthreads();
        if (threadIdx.x < 256) {   // conv 4x4 valid on 16 x 4 x 4 == dot over 256 values; 8 threads per output channel
            const int o = threadIdx.x >> 3, part = threadIdx.x & 7;
            float s = 0.f;
            for (int k = part; k < 256; k += 8) s = fmaf(wsm[W4 + o * 256 + k], p4[k], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (part == 0) v5[o] = fmaxf(s + wsm[B4 + o], 0.f);
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            float s = wsm[F1B + threadIdx.x];
            for (int k = 0; k < 32; ++k) s = fmaf(wsm[F1W + threadIdx.x * 32 + k], v5[k], s);
            v6[threadIdx.x] = fmaxf(s, 0.f);
            __syncwarp();
            float t = wsm[F2W + threadIdx.x] * v6[threadIdx.x];
            t = warp_sum(t);
            if (threadIdx.x == 0) pred[f] = 1.f / (1.f + expf(-(t + wsm[F2B])));
        }
    }
}

}  // namespace cvae

using namespace cvae;

extern "C" int cvae_critic_param_count(void) { return kCriticFloats; }

extern "C" int cvae_critic_fwd(int frames, const float* x, const float* weights, float* pred, void* stream) {
    CVAE_REQUIRE(frames >= 0 && (frames == 0 || (x && weights && pred)), CVAE_EINVAL, "critic_fwd: bad argument");
    if (frames == 0) return CVAE_OK;
    const size_t smem = sizeof(float) * (11876 + 3 * 4096 + 8 * 1024 + 8 * 256 + 8 * 64 + 256 + 64);
    CVAE_OPT_IN_SMEM(critic_fwd_kernel, smem);
    // one frame per CTA and round: the fewest CTAs that need the same number of rounds (256 frames: 128 CTAs x 2)
    const int sms = sm_count(), rounds = (frames + sms - 1) / sms;
    const int grid = (frames + rounds - 1) / rounds;
    cvae::launch(critic_fwd_kernel, grid, kCriticThreads, smem, (cudaStream_t)stream, frames, x, weights, pred);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

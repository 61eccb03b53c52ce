// Reconstruction + KL loss (vae_nets.py:53-62): 5-level MS-SSIM (vae_nets.py:150-247) forward and
// backward, KLD (vae_nets.py:57-58).
//
// One CTA per (frame, channel) plane keeps both image pyramids in shared memory and runs the 11x11
// window as two separable 11-tap passes (the reference window is an outer product, vae_nets.py:175-179)
// with up to 8 outputs per thread per pass so the FP32 pipe, not shared-memory bandwidth, is the limiter;
// all five pyramid levels go through each pass together.
// The MS-SSIM means are batch-global (vae_nets.py:207,212) and enter the loss non-linearly
// (vae_nets.py:243-246), hence two phases: forward accumulates the ten level sums with double
// atomics, loss_finalize turns them into the loss and the per-level chain-rule coefficients, and
// the backward kernel recomputes the maps and applies them.  powf of a negative cs mean gives NaN
// exactly like the reference's `mcs ** weights`.
#include "common.cuh"

namespace cvae {

struct Window { float g[11]; };

static constexpr int kLevels = 5;
// padded strides (S+1) keep both the row-wise and the column-wise pass bank-conflict free
__host__ __device__ constexpr int lvl_size(int l) { return 64 >> l; }
__host__ __device__ constexpr int lvl_off(int l) {
    int o = 0;
    for (int i = 0; i < l; ++i) o += lvl_size(i) * (lvl_size(i) + 1);
    return o;
}
static constexpr int kPyr = lvl_off(5);          // 5580 floats
static constexpr int kMap = 64 * 65;             // one padded 64x64 map
static constexpr int kMsThreads = 512;

// Every "map" below is a pyramid-shaped buffer (kPyr floats): level l of map m sits at m * kPyr + lvl_off(l).
//
// Task decomposition.  In every phase a thread runs ONE level-0 task (8 outputs; 64 rows x 8 segments = 512
// tasks = one per thread) and at most ONE task of the coarser levels, whose tasks are made finer so the whole
// tail of the pyramid is a single round as well (warp-aligned ranges):
//   threads   0..255  level 1, 4 outputs (32 x 8 tasks)      threads 384..447  level 3, 1 output (8 x 8)
//   threads 256..383  level 2, 2 outputs (16 x 8 tasks)      threads 448..463  level 4, 1 output (4 x 4)
// Running the five levels one after the other (the first version) left the small levels latency bound: a
// level costs one full dependent 11-tap chain per thread no matter how few threads have work.
__device__ __forceinline__ void small_task(int tid, int& level, int& out, int& t) {
    if (tid < 256) { level = 1; out = 4; t = tid; }
    else if (tid < 384) { level = 2; out = 2; t = tid - 256; }
    else if (tid < 448) { level = 3; out = 1; t = tid - 384; }
    else if (tid < 464) { level = 4; out = 1; t = tid - 448; }
    else { level = -1; out = 0; t = 0; }
}

// horizontal 11-tap pass of task t (row = t % S, OUT columns from (t / S) * OUT) producing NM maps from
// per-pixel inputs built by `make`.
template <int NM, int OUT, typename Make>
__device__ __forceinline__ void hpass(int S, int t, const Window& w, float* __restrict__ out, Make make) {
    const int st = S + 1;
    const int r = t % S, c0 = (t / S) * OUT;
    float acc[NM][OUT];
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
        for (int j = 0; j < OUT; ++j) acc[m][j] = 0.f;
#pragma unroll
    for (int i = 0; i < OUT + 10; ++i) {
        const int c = c0 - 5 + i;
        float v[NM];
        if (c >= 0 && c < S) make(r * st + c, v);
        else {
#pragma unroll
            for (int m = 0; m < NM; ++m) v[m] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < OUT; ++j) {
            const int k = i - j;  // tap index: input c = (c0+j) - 5 + k
            if (k >= 0 && k < 11) {
#pragma unroll
                for (int m = 0; m < NM; ++m) acc[m][j] = fmaf(w.g[k], v[m], acc[m][j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < OUT; ++j)
        if (c0 + j < S) {
#pragma unroll
            for (int m = 0; m < NM; ++m) out[m * kPyr + r * st + c0 + j] = acc[m][j];
        }
}

// vertical 11-tap pass of task t (column = t % S, OUT rows from (t / S) * OUT) over NM maps;
// `sink(idx, vals)` consumes the blurred values of pixel idx.
template <int NM, int OUT, typename Sink>
__device__ __forceinline__ void vpass(int S, int t, const Window& w, const float* __restrict__ in, Sink sink) {
    const int st = S + 1;
    const int c = t % S, r0 = (t / S) * OUT;
    float acc[NM][OUT];
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
        for (int j = 0; j < OUT; ++j) acc[m][j] = 0.f;
#pragma unroll
    for (int i = 0; i < OUT + 10; ++i) {
        const int r = r0 - 5 + i;
        if (r >= 0 && r < S) {
#pragma unroll
            for (int m = 0; m < NM; ++m) {
                const float v = in[m * kPyr + r * st + c];
#pragma unroll
                for (int j = 0; j < OUT; ++j) {
                    const int k = i - j;
                    if (k >= 0 && k < 11) acc[m][j] = fmaf(w.g[k], v, acc[m][j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < OUT; ++j)
        if (r0 + j < S) {
            float v[NM];
#pragma unroll
            for (int m = 0; m < NM; ++m) v[m] = acc[m][j];
            sink((r0 + j) * st + c, v);
        }
}

// run `body(level, S, OUT-tag, task)` for this thread's level-0 task and its coarse-level task
#define CVAE_MS_FOR_TASKS(...)                                                   \
    {                                                                            \
        { constexpr int OUT = 8; const int lvl = 0, S = 64, t = threadIdx.x; __VA_ARGS__ } \
        int lvl_, out_, t_;                                                      \
        small_task(threadIdx.x, lvl_, out_, t_);                                 \
        if (out_ == 4) { constexpr int OUT = 4; const int lvl = 1, S = 32, t = t_; __VA_ARGS__ }        \
        else if (out_ == 2) { constexpr int OUT = 2; const int lvl = 2, S = 16, t = t_; __VA_ARGS__ }   \
        else if (out_ == 1) { constexpr int OUT = 1; const int lvl = lvl_, S = 64 >> lvl_, t = t_; __VA_ARGS__ } \
    }

// The scalar part of the loss from the ten batch-global level sums (vae_nets.py:224-246): returns P = prod_l<4 (cs_l^w_l *
// ssim_4^w_4) (recon loss = 1 - P) and writes the chain-rule coefficients dL/d(map pixel): coef[0..3] for the cs maps of
// levels 0..3, coef[4] for the ssim map of level 4.  Shared by loss_finalize_kernel and by the backward kernel when it is
// given the sums instead of the coefficients (so that loss_finalize can run beside it).
__device__ __forceinline__ float ms_chain_coefficients(const double* __restrict__ sums, int B, float* coef) {
    const float wts[5] = {0.0448f, 0.2856f, 0.3001f, 0.2363f, 0.1333f};
    float cs[5], ss4, P = 1.f;
    for (int l = 0; l < 5; ++l) cs[l] = (float)(sums[l] / ((double)B * 3 * lvl_size(l) * lvl_size(l)));
    ss4 = (float)(sums[9] / ((double)B * 3 * 16));
    const float p4 = powf(ss4, wts[4]);
    for (int l = 0; l < 4; ++l) P *= powf(cs[l], wts[l]) * p4;   // torch.prod(pow1[:-1] * pow2[-1])
    // L = 1 - P;  dL/dcs_l = -P w_l / cs_l;  dL/dss4 = -P 4 w_4 / ss4;  per-pixel: / N_l
    for (int l = 0; l < 4; ++l) coef[l] = (-P * wts[l] / cs[l]) / ((float)B * 3.f * lvl_size(l) * lvl_size(l));
    coef[4] = (-P * 4.f * wts[4] / ss4) / ((float)B * 3.f * 16.f);
    return P;
}

// coef layout (floats): [0..3] dL/d(cs_map pixel) for levels 0..3, [4] dL/d(ssim_map pixel) level 4
template <bool BWD>
__global__ void __launch_bounds__(kMsThreads, 1)
msssim_kernel(int planes, const float* __restrict__ recon, const float* __restrict__ x, const Window w,
              double* __restrict__ sums, const float* __restrict__ coef, const float* __restrict__ grad_out,
              float* __restrict__ d_recon) {
    grid_dependency_sync();
    extern __shared__ float sm[];
    float* A = sm;                    // recon pyramid
    float* Bp = A + kPyr;             // target pyramid
    float* Hm = Bp + kPyr;            // 5 horizontally blurred maps (reused for the 3 gradient maps)
    float* Dm = Hm + 5 * kPyr;        // BWD: 3 derivative maps; map 0 is reused for the per-level gradient G
    __shared__ float wsum[2][2][kMsThreads / 32];   // [level-0 | coarse level][cs | ssim][warp]
    const float C1 = 0.0001f, C2 = 0.0009f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float gcoef[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (BWD) {
        if (coef != nullptr) {
#pragma unroll
            for (int l = 0; l < 5; ++l) gcoef[l] = coef[l];
        } else {
            ms_chain_coefficients(sums, planes / 3, gcoef);      // (every thread for itself: five powf, no barrier)
        }
    }

    double cta_tot = 0.0;      // forward, threads 0..9: this CTA's share of the ten level sums (one atomic per CTA, not per plane)
    for (int plane = blockIdx.x; plane < planes; plane += gridDim.x) {
        const float* ra = recon + (size_t)plane * 4096;
        const float* xb = x + (size_t)plane * 4096;
        __syncthreads();
        for (int i = threadIdx.x; i < 4096; i += kMsThreads) {
            const int r = i >> 6, c = i & 63;
            A[r * 65 + c] = __ldg(ra + i);
            Bp[r * 65 + c] = __ldg(xb + i);
        }
        __syncthreads();
        for (int l = 1; l < kLevels; ++l) {   // F.avg_pool2d(., (2,2)) pyramid, vae_nets.py:232-233
            const int S = lvl_size(l), st = S + 1, ps = 2 * S + 1;
            const float* pa = A + lvl_off(l - 1);
            const float* pb = Bp + lvl_off(l - 1);
            for (int i = threadIdx.x; i < S * S; i += kMsThreads) {
                const int r = i / S, c = i % S;
                const int q = 2 * r * ps + 2 * c;
                A[lvl_off(l) + r * st + c] = 0.25f * (pa[q] + pa[q + 1] + pa[q + ps] + pa[q + ps + 1]);
                Bp[lvl_off(l) + r * st + c] = 0.25f * (pb[q] + pb[q + 1] + pb[q + ps] + pb[q + ps + 1]);
            }
            __syncthreads();
        }

        // ---- phase 1: horizontal blur of (a, b, a*a, b*b, a*b), all levels ----
        CVAE_MS_FOR_TASKS({
            const float* a = A + lvl_off(lvl);
            const float* b = Bp + lvl_off(lvl);
            hpass<5, OUT>(S, t, w, Hm + lvl_off(lvl), [&](int idx, float* v) {
                const float av = a[idx], bv = b[idx];
                v[0] = av; v[1] = bv; v[2] = av * av; v[3] = bv * bv; v[4] = av * bv;
            });
        })
        __syncthreads();
        // ---- phase 2: vertical blur -> SSIM / CS maps (forward: sums; backward: derivative maps) ----
        float acc_cs[2] = {0.f, 0.f}, acc_ss[2] = {0.f, 0.f};
        CVAE_MS_FOR_TASKS({
            const float gc = BWD ? gcoef[lvl < 4 ? lvl : 4] : 0.f;
            float* dm = Dm + lvl_off(lvl);
            float cs_acc = 0.f, ss_acc = 0.f;
            vpass<5, OUT>(S, t, w, Hm + lvl_off(lvl), [&](int idx, const float* v) {
                const float mu1 = v[0], mu2 = v[1];
                const float s1 = v[2] - mu1 * mu1, s2 = v[3] - mu2 * mu2, s12 = v[4] - mu1 * mu2;
                const float v1 = 2.f * s12 + C2, v2 = s1 + s2 + C2;
                const float a1 = 2.f * mu1 * mu2 + C1, a2 = mu1 * mu1 + mu2 * mu2 + C1;
                const float cs = v1 / v2;
                cs_acc += cs;
                ss_acc += (a1 * v1) / (a2 * v2);
                if (BWD) {
                    // derivative of the level's scalar w.r.t. mu1, blur(a*a), blur(a*b) at this pixel
                    float dmu, daa, dab;
                    if (lvl < 4) {         // cs = v1 / v2
                        dab = 2.f / v2;
                        daa = -v1 / (v2 * v2);
                        dmu = (-2.f * mu2) / v2 + (v1 / (v2 * v2)) * (2.f * mu1);
                    } else {               // ssim = (a1 v1) / (a2 v2)
                        const float den = a2 * v2, Sv = (a1 * v1) / den;
                        const float dv1 = a1 / den, dv2 = -Sv / v2, da1 = v1 / den, da2 = -Sv / a2;
                        dab = 2.f * dv1;
                        daa = dv2;
                        dmu = dv1 * (-2.f * mu2) + dv2 * (-2.f * mu1) + da1 * (2.f * mu2) + da2 * (2.f * mu1);
                    }
                    dm[idx] = dmu * gc;
                    dm[kPyr + idx] = daa * gc;
                    dm[2 * kPyr + idx] = dab * gc;
                }
            });
            acc_cs[lvl > 0] = cs_acc;
            acc_ss[lvl > 0] = ss_acc;
        })
        if (!BWD) {
            // per-warp partial sums; a warp's coarse-level tasks all belong to one level (warp-aligned ranges)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float c = warp_sum(acc_cs[q]), s2 = warp_sum(acc_ss[q]);
                if (lane == 0) { wsum[q][0][warp] = c; wsum[q][1][warp] = s2; }
            }
            __syncthreads();
            if (threadIdx.x < 10) {
                const int which = threadIdx.x / 5, l = threadIdx.x % 5;
                // warps of level l in the coarse ranges: L1 0-7, L2 8-11, L3 12-13, L4 14
                const int w0 = l == 0 ? 0 : (l == 1 ? 0 : (l == 2 ? 8 : (l == 3 ? 12 : 14)));
                const int w1 = l == 0 ? kMsThreads / 32 : (l == 1 ? 8 : (l == 2 ? 12 : (l == 3 ? 14 : 15)));
                float tot = 0.f;
                for (int i = w0; i < w1; ++i) tot += wsum[l > 0][which][i];
                cta_tot += (double)tot;
            }
        } else {
            __syncthreads();
            // ---- phase 3 / 4: dL/da_l(q) = blur(dmu)(q) + 2 a(q) blur(daa)(q) + b(q) blur(dab)(q)  (symmetric window)
            CVAE_MS_FOR_TASKS({
                const float* dm = Dm + lvl_off(lvl);
                hpass<3, OUT>(S, t, w, Hm + lvl_off(lvl), [&](int idx, float* v) {
                    v[0] = dm[idx]; v[1] = dm[kPyr + idx]; v[2] = dm[2 * kPyr + idx];
                });
            })
            __syncthreads();
            CVAE_MS_FOR_TASKS({
                const float* a = A + lvl_off(lvl);
                const float* b = Bp + lvl_off(lvl);
                float* g = Dm + lvl_off(lvl);    // derivative map 0 is dead after phase 3
                vpass<3, OUT>(S, t, w, Hm + lvl_off(lvl), [&](int idx, const float* v) {
                    g[idx] = v[0] + 2.f * a[idx] * v[1] + b[idx] * v[2];
                });
            })
            __syncthreads();
            // chain through the average pools: each finer pixel inherits 1/4 of its parent's gradient
            const float go = grad_out ? __ldg(grad_out) : 1.f;
            float* dr = d_recon + (size_t)plane * 4096;
            for (int i = threadIdx.x; i < 4096; i += kMsThreads) {
                const int r = i >> 6, c = i & 63;
                float acc = 0.f, scale = 1.f;
#pragma unroll
                for (int l = 0; l < kLevels; ++l) {
                    const int S = lvl_size(l);
                    acc += scale * Dm[lvl_off(l) + (r >> l) * (S + 1) + (c >> l)];
                    scale *= 0.25f;
                }
                dr[i] = acc * go;
            }
        }
    }
    if (!BWD && threadIdx.x < 10) atomicAdd(sums + (threadIdx.x / 5) * 5 + threadIdx.x % 5, cta_tot);
}
#undef CVAE_MS_FOR_TASKS

// one block: KLD reduction + loss scalars + chain-rule coefficients
__global__ void loss_finalize_kernel(int B, const float* __restrict__ ml, const double* __restrict__ kld_partial, int n_partial,
                                     const double* __restrict__ sums, float kld_weight, float* __restrict__ losses,
                                     float* __restrict__ coef) {
    grid_dependency_sync();
    __shared__ double red[32];
    double acc = 0.0;
    if (kld_partial) {   // the latent kernel already reduced the KL term per 64 rows: add the partials in block order
        if (threadIdx.x == 0)
            for (int i = 0; i < n_partial; ++i) acc += kld_partial[i];
    } else {
        for (int i = threadIdx.x; i < B * 32; i += blockDim.x) {
            const int b = i >> 5, d = i & 31;
            const float mu = ml[b * 64 + d], lv = ml[b * 64 + 32 + d];
            acc += (double)(1.f + lv - mu * mu - expf(lv));
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
        const float kld = (float)(-0.5 * tot / B) * kld_weight;
        float c[5];
        const float P = ms_chain_coefficients(sums, B, c);
        const float recon_loss = 1.f - P;
        losses[0] = recon_loss + kld;
        losses[1] = recon_loss;
        losses[2] = kld;
        for (int l = 0; l < 5; ++l) coef[l] = c[l];
    }
}

// dKLD/dmu = w mu / B, dKLD/dlogvar = w 0.5 (exp(lv) - 1) / B, times the upstream gradient
__global__ void kld_bwd_kernel(int B, const float* __restrict__ ml, float kld_weight, const float* __restrict__ grad_out,
                               float* __restrict__ dmu, float* __restrict__ dlv) {
    grid_dependency_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 32) return;
    const int b = i >> 5, d = i & 31;
    const float go = (grad_out ? __ldg(grad_out) : 1.f) * kld_weight / (float)B;
    dmu[i] = go * ml[b * 64 + d];
    dlv[i] = go * 0.5f * (expf(ml[b * 64 + 32 + d]) - 1.f);
}

}  // namespace cvae

using namespace cvae;

static size_t ms_smem(bool bwd) {
    return sizeof(float) * (size_t)(2 * kPyr + 5 * kPyr + (bwd ? 3 * kPyr : 0));
}

// One persistent CTA per SM walks planes blockIdx.x, + grid, ...: with 768 planes on 148 SMs the sixth round would run on 28
// SMs only.  Launch the fewest CTAs that still need the same number of rounds (768 / 6 = 128): same duration, and the
// other SMs stay free for the kernels of the side streams.
static int ms_grid(int planes) {
    const int sms = sm_count();
    if (planes <= sms) return planes;
    const int rounds = (planes + sms - 1) / sms;
    return (planes + rounds - 1) / rounds;
}

// The two halves of cvae_loss_fwd as calls of their own, so that a training step can run the second one (the scalars the host
// reads: one block) on another stream beside cvae_loss_bwd_sums, which derives its coefficients from the sums itself.
extern "C" int cvae_loss_sums(int batch, const float* recon, const float* x, const float* window11, double* sums, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && recon && x && window11 && sums, CVAE_EINVAL, "loss_sums: bad argument");
    Window w;
    memcpy(w.g, window11, sizeof(w.g));
    CVAE_OPT_IN_SMEM(msssim_kernel<false>, ms_smem(false));
    CVAE_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 10, stream));
    const int planes = batch * 3;
    const int grid = ms_grid(planes);
    cvae::launch(msssim_kernel<false>, grid, kMsThreads, ms_smem(false), stream, planes, recon, x, w, sums, nullptr, nullptr, nullptr);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_loss_finalize(int batch, const float* mu_logvar, const double* kld_partial, const double* sums, float kld_weight,
                                  float* coef, float* losses, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && (mu_logvar || kld_partial) && sums && coef && losses, CVAE_EINVAL, "loss_finalize: bad argument");
    cvae::launch(loss_finalize_kernel, 1, kld_partial ? 32 : 1024, 0, stream, batch, mu_logvar, kld_partial, (batch + 63) / 64, sums, kld_weight, losses, coef);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_loss_fwd(int batch, const float* recon, const float* x, const float* mu_logvar, const double* kld_partial,
                             const float* window11, float kld_weight, double* sums, float* coef, float* losses,
                             void* stream_) {
    CVAE_REQUIRE(batch > 0 && recon && x && (mu_logvar || kld_partial) && window11 && sums && coef && losses, CVAE_EINVAL, "loss_fwd: bad argument");
    const int rc = cvae_loss_sums(batch, recon, x, window11, sums, stream_);
    if (rc != CVAE_OK) return rc;
    return cvae_loss_finalize(batch, mu_logvar, kld_partial, sums, kld_weight, coef, losses, stream_);
}

// d(recon loss)/d(recon) from the level sums cvae_loss_sums left (the coefficients cvae_loss_finalize would hand to
// cvae_loss_bwd are recomputed in the kernel: bit-identical); the KL term's gradient is the caller's (cvae_latent_bwd /
// cvae_bottleneck_bwd with kld_grad_scale).
extern "C" int cvae_loss_bwd_sums(int batch, const float* recon, const float* x, const float* window11, const double* sums,
                                  const float* grad_out, float* d_recon, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && recon && x && window11 && sums && d_recon, CVAE_EINVAL, "loss_bwd_sums: bad argument");
    Window w;
    memcpy(w.g, window11, sizeof(w.g));
    CVAE_OPT_IN_SMEM(msssim_kernel<true>, ms_smem(true));
    const int planes = batch * 3;
    const int grid = ms_grid(planes);
    cvae::launch(msssim_kernel<true>, grid, kMsThreads, ms_smem(true), stream, planes, recon, x, w, const_cast<double*>(sums), nullptr, grad_out, d_recon);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

extern "C" int cvae_loss_bwd(int batch, const float* recon, const float* x, const float* mu_logvar,
                             const float* window11, float kld_weight, const float* coef, const float* grad_out,
                             float* d_recon, float* d_mu, float* d_logvar, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    CVAE_REQUIRE(batch > 0 && recon && x && window11 && coef && d_recon && (d_mu == nullptr) == (d_logvar == nullptr) &&
                     (d_mu == nullptr || mu_logvar), CVAE_EINVAL, "loss_bwd: bad argument");
    Window w;
    memcpy(w.g, window11, sizeof(w.g));
    CVAE_OPT_IN_SMEM(msssim_kernel<true>, ms_smem(true));
    const int planes = batch * 3;
    const int grid = ms_grid(planes);
    cvae::launch(msssim_kernel<true>, grid, kMsThreads, ms_smem(true), stream, planes, recon, x, w, nullptr, coef, grad_out, d_recon);
    CVAE_LAUNCH_CHECK();
    if (d_mu) {   // NULL, NULL: the caller folds the KL term's backward into cvae_latent_bwd (kld_grad_scale)
        cvae::launch(kld_bwd_kernel, (batch * 32 + 255) / 256, 256, 0, stream, batch, mu_logvar, kld_weight, grad_out, d_mu, d_logvar);
        CVAE_LAUNCH_CHECK();
    }
    return CVAE_OK;
}

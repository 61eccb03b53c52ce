// Weight packing: fp32 master weights (reference layouts: OIHW conv, [out][in] linear) -> the bf16
// UMMA core-matrix blocks conv_gemm.cu streams (forward, data-gradient and up-sample-folded "phase"
// variants) and the transposed / NHWC-permuted fp32 matrices of the three Linear layers.
// One launch packs every layer; it runs at the head of each forward so Adam updates, load_state_dict
// or any user edit of a parameter are always picked up.
#include "common.cuh"
#include "wa_groups.cuh"

namespace cvae {

struct PackJobs {
    int count;
    long long total;
    cvae_pack_job job[CVAE_MAX_PACK_JOBS];
    long long start[CVAE_MAX_PACK_JOBS + 1];
};

// 5x5 taps (ky range lo..hi) folded onto low-res tap t (0..2) for output phase a:
// conv5x5(upsample2(x))[2i+a] reads x[(2i+a+ky-2)>>1]
__device__ __forceinline__ void phase_range(int a, int t, int& lo, int& hi) {
    if (a == 0) { lo = t == 0 ? 0 : (t == 1 ? 2 : 4); hi = t == 0 ? 1 : (t == 1 ? 3 : 4); }
    else        { lo = t == 0 ? 0 : (t == 1 ? 1 : 3); hi = t == 0 ? 0 : (t == 1 ? 2 : 4); }
}

// effective 3x3 weight of phase (a,b) for (co, ci, ty, tx): fp32 sum in ky-major order
__device__ __forceinline__ float phase_weight(const float* W, int cin, int co, int ci, int a, int b, int ty, int tx) {
    int y0, y1, x0, x1;
    phase_range(a, ty, y0, y1);
    phase_range(b, tx, x0, x1);
    float acc = 0.f;
    for (int ky = y0; ky <= y1; ++ky)
        for (int kx = x0; kx <= x1; ++kx) acc = acc + W[((size_t)co * cin + ci) * 25 + ky * 5 + kx];
    return acc;
}

// Value of element i = i0 + e of job j.  The bf16 layouts all have the 8 channels of one core-matrix row as their fastest
// index (e), so the kernel calls this in an unrolled loop over e and the compiler hoists everything that depends on i0 only.
__device__ __forceinline__ float pack_value(const cvae_pack_job& j, long long i0, int e, bool& is_bf16) {
    const float* W = (const float*)j.src;
    const long long i = i0 + e;
    is_bf16 = true;
    if (j.kind == CVAE_PACK_FC) {         // dst fp32 [4096 k'][64]: k' = p*256 + c  <->  k = c*16 + p
        is_bf16 = false;
        const int jj = (int)(i % 64), kp = (int)(i / 64);
        const int k = (kp % 256) * 16 + kp / 256;
        return jj < 32 ? W[(size_t)jj * 4096 + k] : ((const float*)j.src2)[(size_t)(jj - 32) * 4096 + k];
    }
    if (j.kind == CVAE_PACK_DECIN) {      // dst fp32 [34][4096 k']: rows 0..32 = W^T, row 33 = bias
        is_bf16 = false;
        const int kp = (int)(i % 4096), r = (int)(i / 4096);
        const int k = (kp % 256) * 16 + kp / 256;
        return r < 33 ? W[(size_t)k * 33 + r] : ((const float*)j.src2)[k];
    }
    // UMMA K-major blocks: [nb][kstep][n/8][kchunk 2][row 8][elem 8]
    const int N = j.n, NB = N < 128 ? N : 128;
    const int r = (int)((i0 >> 3) & 7), kc = (int)((i0 >> 6) & 1);
    const int ng = (int)((i0 >> 7) % (NB / 8));
    const long long rest = i0 / (16LL * NB);
    const int ks = (int)(rest % j.ksteps), nb = (int)(rest / j.ksteps);
    int n = nb * NB + ng * 8 + r;
    const int k16 = kc * 8 + e;

    if (j.kind == CVAE_PACK_PAIR8) {      // encoder conv 0: two taps of an 8-channel padded pixel per K step
        const int c = k16 & 7, half = k16 >> 3;
        if (c >= 3) return 0.f;
        int ky, kx;
        if (ks < 10) { ky = ks >> 1; kx = (ks & 1) * 2 + half; }
        else if (ks < 12) { ky = (ks - 10) * 2 + half; kx = 4; }
        else { if (half) return 0.f; ky = 4; kx = 4; }
        return W[((size_t)n * 3 + c) * 25 + ky * 5 + kx];
    }
    const int cpt = j.k_channels / 16;  // K steps per tap
    int tap, c;
    if (j.kind & (CVAE_PACK_KORDER_BLOCK64 | CVAE_PACK_KORDER_BLOCK32)) {
        // conv_wa.cu: K steps ordered (channel block, tap group, 16-channel step); with tap stacking GEMM row n of a
        // 128-row block is (channel n / J, jj = n % J) and carries tap dx = s - jj of its group, or zeros
        const int kb = (j.kind & CVAE_PACK_KORDER_BLOCK64) ? 64 : 32, spu = kb / 16;
        const int J = (j.kind & CVAE_PACK_STACK4) ? 4 : ((j.kind & CVAE_PACK_STACK2) ? 2 : 1);
        const int base = j.kind & 0xFF;
        const int ksize = (base == CVAE_PACK_FWD5 || base == CVAE_PACK_DGRAD5) ? 5 : 3;
        const int groups = wa_group_count(ksize, J);
        const int blk = ks / (groups * spu), rem = ks - blk * groups * spu;
        const int g = rem / spu;
        c = blk * kb + (rem - g * spu) * 16 + k16;
        int dy, s, lo;
        wa_group(ksize, J, g, dy, s, lo);
        const int nl = n % 128, jj = nl % J, dx = s - jj;
        if (dx < lo) return 0.f;
        n = (n / 128) * (128 / J) + nl / J;
        tap = (dy + ksize / 2) * ksize + dx + ksize / 2;
    } else {
        tap = ks / cpt;
        c = (ks % cpt) * 16 + k16;
    }
    switch (j.kind & 0xFF) {
        case CVAE_PACK_FWD5:     // n = co, c = ci
            return W[((size_t)n * j.cin + c) * 25 + tap];
        case CVAE_PACK_DGRAD5:   // n = ci, c = co, flipped taps
            return W[((size_t)c * j.cin + n) * 25 + (24 - tap)];
        case CVAE_PACK_PHASE_FWD: {  // n = (ab, co), c = ci, 3x3 taps
            if (n >= 4 * j.cout) return 0.f;
            const int ab = n / j.cout, co = n % j.cout;
            return phase_weight(W, j.cin, co, c, ab >> 1, ab & 1, tap / 3, tap % 3);
        }
        case CVAE_PACK_PHASE_DGRAD: {  // n = ci, c = (ab, co), flipped 3x3 taps
            if (c >= 4 * j.cout) return 0.f;
            const int ab = c / j.cout, co = c % j.cout, ft = 8 - tap;
            return phase_weight(W, j.cin, co, n, ab >> 1, ab & 1, ft / 3, ft % 3);
        }
    }
    return 0.f;
}

__global__ void pack_weights_kernel(const PackJobs jobs) {
    grid_dependency_sync();
    // one thread = 8 consecutive elements (every job's element count is a multiple of 8): one index decode, one 16-byte store
    for (long long idx8 = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx8 * 8 < jobs.total;
         idx8 += (long long)gridDim.x * blockDim.x) {
        const long long idx = idx8 * 8;
        int lo = 0, hi = jobs.count - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (jobs.start[mid] <= idx) lo = mid; else hi = mid - 1;
        }
        const cvae_pack_job& j = jobs.job[lo];
        const long long i0 = idx - jobs.start[lo];
        bool is_bf16 = true;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = pack_value(j, i0, e, is_bf16);
        if (is_bf16) {
            *reinterpret_cast<uint4*>((__nv_bfloat16*)j.dst + i0) =
                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        } else {
            float4* o = reinterpret_cast<float4*>((float*)j.dst + i0);
            o[0] = make_float4(v[0], v[1], v[2], v[3]);
            o[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
}

}  // namespace cvae

using namespace cvae;

extern "C" int64_t cvae_pack_elems(const cvae_pack_job* j) {
    if (!j) return -1;
    if (j->kind == CVAE_PACK_FC) return 4096LL * 64;
    if (j->kind == CVAE_PACK_DECIN) return 34LL * 4096;
    return (int64_t)j->n * j->ksteps * 16;
}

extern "C" int cvae_pack_weights(const cvae_pack_job* jobs, int count, void* stream) {
    CVAE_REQUIRE(jobs && count > 0 && count <= CVAE_MAX_PACK_JOBS, CVAE_EINVAL, "pack_weights: %d jobs", count);
    PackJobs pj{};
    pj.count = count;
    long long total = 0;
    for (int i = 0; i < count; ++i) {
        CVAE_REQUIRE(jobs[i].src && jobs[i].dst, CVAE_EINVAL, "pack_weights: job %d has a null tensor", i);
        if (jobs[i].kind != CVAE_PACK_FC && jobs[i].kind != CVAE_PACK_DECIN)
            CVAE_REQUIRE(jobs[i].n % 16 == 0 && jobs[i].ksteps > 0, CVAE_EINVAL, "pack_weights: job %d shape", i);
        if (jobs[i].kind & ~0xFF) {
            const int kb = (jobs[i].kind & CVAE_PACK_KORDER_BLOCK64) ? 64 : ((jobs[i].kind & CVAE_PACK_KORDER_BLOCK32) ? 32 : 0);
            CVAE_REQUIRE(kb && jobs[i].k_channels % kb == 0 && (jobs[i].kind & 0xFF) != CVAE_PACK_PAIR8 && (jobs[i].kind & 0xFF) <= CVAE_PACK_PHASE_DGRAD &&
                             jobs[i].n % 128 == 0, CVAE_EINVAL, "pack_weights: job %d cannot be packed block-major", i);
        }
        pj.job[i] = jobs[i];
        pj.start[i] = total;
        total += cvae_pack_elems(&jobs[i]);
    }
    pj.start[count] = total;
    pj.total = total;
    CVAE_REQUIRE(total % 8 == 0, CVAE_EINVAL, "pack_weights: element counts must be multiples of 8");
    const int threads = 256;
    long long blocks = (total / 8 + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    cvae::launch(pack_weights_kernel, (int)blocks, threads, 0, (cudaStream_t)stream, pj);
    CVAE_LAUNCH_CHECK();
    return CVAE_OK;
}

// Tap groups of the weights-as-A convolution (conv_wa.cu), shared by the kernel's launch code and the weight packer
// (pack.cu) so the two can never disagree.
//
// For a ksize x ksize filter and a stacking factor J (1, 2, 4) the taps dx in [-pad, pad] of every filter row are cut
// into chunks of J from the left; chunk k covers dx in [lo, hi], lo = -pad + k J, hi = min(lo + J - 1, pad).  The
// group's B operand (pixels) is shifted by s = hi, and GEMM row (channel, j) carries the weights of tap dx = s - j
// (all-zero when dx < lo).  Groups are ordered (filter row, chunk); J = 1 reproduces the plain tap order.
#pragma once

namespace cvae {

static constexpr int kWaMaxGroups = 32;

__host__ __device__ inline int wa_chunks(int ksize, int J) { return (ksize + J - 1) / J; }
__host__ __device__ inline int wa_group_count(int ksize, int J) { return ksize * wa_chunks(ksize, J); }
__host__ __device__ inline void wa_group(int ksize, int J, int g, int& dy, int& s, int& lo) {
    const int pad = ksize / 2, nc = wa_chunks(ksize, J);
    const int row = g / nc, k = g - row * nc;
    dy = row - pad;
    lo = -pad + k * J;
    const int hi = lo + J - 1;
    s = hi < pad ? hi : pad;
}

}  // namespace cvae

/* libcvae.so -- C ABI of the B200-native Critic-VAE hot path.
 *
 * The reference (lcicek/Critic-VAE) has no FFI: its hot path is PyTorch module calls.  This header
 * is the boundary a maintainer would bind instead (INTEGRATION.md shows the ctypes stub); every
 * entry point cites the reference lines it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every function returns 0 (CVAE_OK) or a negative CVAE_E* code; cvae_last_error() gives the
 *     thread-local message.  Nothing throws, nothing allocates or frees caller-visible memory.
 *   - all tensors are caller-owned raw DEVICE pointers on the current device; sizes are explicit.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it, never synchronise,
 *     and are CUDA-graph capturable.
 *   - activations between layers are NHWC bf16; the public tensors (frames, reconstructions, mu,
 *     logvar, parameters, gradients) are fp32 in the reference's layouts (NCHW, OIHW).
 */
#ifndef CVAE_H_
#define CVAE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CVAE_OK 0
#define CVAE_EINVAL (-1)   /* bad argument / unsupported shape */
#define CVAE_ECUDA (-2)    /* CUDA runtime error */
#define CVAE_EDEVICE (-3)  /* device-side pipeline fault (bounded wait expired) */

const char* cvae_last_error(void);
int cvae_version(void);
/* Reads and clears the device-side fault flag raised by bounded waits.  Synchronises `stream`. */
int cvae_check_device_fault(void* stream);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution on tcgen05 (forward, data-gradient and up-sample-folded variants).
 * Replaces nn.Conv2d / nn.Upsample / their autograd in vae_nets.py:69,74,79,84 (encoder) and
 * :117-133 (decoder).  GEMM rows are pixels of a `height x width` grid, columns are `n_total`
 * output channels, K runs over ksize*ksize taps x src_channels.
 * ---------------------------------------------------------------------------------------------- */
enum cvae_loader {
    CVAE_LOAD_NHWC = 0,          /* src: bf16 [B][H][W][C]                                          */
    CVAE_LOAD_NCHW3 = 1,         /* src: fp32 [B][3][H][W] frames, padded to 8 channels (E0)        */
    CVAE_LOAD_S2D = 2,           /* src: bf16 [B][2H][2W][C/4], space-to-depth view (decoder dgrad) */
    CVAE_LOAD_S2D_NCHW3_DTANH = 3 /* src: fp32 d_recon [B][3][2H][2W], src2: recon; * (1 - recon^2) */
};
enum cvae_epilogue {
    CVAE_EPI_STATS = 0,          /* raw bf16 NHWC + per-channel sum / sum-of-squares (pre-BatchNorm) */
    CVAE_EPI_BIAS_RELU = 1,      /* + bias, ReLU, bf16 NHWC                                          */
    CVAE_EPI_PHASE_BIAS_RELU = 2,/* n_total = 4*C: depth-to-space to [B][2H][2W][C], + bias, ReLU   */
    CVAE_EPI_PHASE_BIAS_TANH = 3,/* n_total = 16 (12 used): fp32 NCHW [B][3][2H][2W], + bias, tanh  */
    CVAE_EPI_MASK = 4,           /* * (act > 0), bf16 NHWC (ReLU backward)                           */
    CVAE_EPI_PLAIN = 5           /* bf16 NHWC                                                        */
};
enum cvae_ktab { CVAE_KTAB_GENERIC = 0, CVAE_KTAB_PAIR8 = 1 /* 8-channel source, two taps per K step */ };

typedef struct {
    int32_t batch, height, width;
    int32_t ksize;         /* 5 or 3 */
    int32_t src_channels;  /* A-operand channels after the loader transform, multiple of 8 */
    int32_t n_total;       /* GEMM N, multiple of 16 */
    int32_t loader, epilogue, ktab;
    int32_t tm;            /* 128-pixel tiles per CTA pass; 0 = automatic */
    const void* src;
    const void* src2;
    const void* wpack;     /* packed by cvae_pack_weights */
    const float* bias;
    const void* act;
    void* out;
    double* stats;         /* [2][n_total], accumulated with atomics (zero it first) */
} cvae_conv_desc;

int cvae_conv_gemm(const cvae_conv_desc* d, void* stream);
/* number of K=16 steps of a conv GEMM: the packed weight tensor is [n_total/nb][ksteps][nb][16] bf16,
 * nb = min(n_total, 128) */
int cvae_conv_ksteps(int ksize, int src_channels, int ktab);

/* ------------------------------------------------------------------------------------------------
 * Convolution weight / bias gradients on tcgen05 (split-K over pixels, fp32 partials, fold to OIHW).
 * Replaces autograd's dW, db of nn.Conv2d at vae_nets.py:69,74,79,84,117,121,125,129,133.
 * ---------------------------------------------------------------------------------------------- */
enum cvae_wgrad_kind {
    CVAE_WGRAD_5X5 = 0,          /* x: bf16 NHWC [B][H][W][cin], dy: bf16 NHWC [B][H][W][cout]               */
    CVAE_WGRAD_PHASE = 1,        /* up-sample-folded conv: x at [B][H][W][cin], dy at [B][2H][2W][cout]      */
    CVAE_WGRAD_SHIFT_FRAMES = 2, /* encoder conv 0: x = fp32 NCHW frames [B][3][H][W], dy bf16 NHWC          */
    CVAE_WGRAD_SHIFT_PHASE12 = 3 /* decoder conv 4: dy = fp32 NCHW d_recon [B][3][2H][2W], dy2 = recon       */
};
typedef struct {
    int32_t kind, batch, height, width;
    int32_t cout, cin;           /* reference weight shape [cout][cin][5][5] */
    int32_t splits;              /* split-K factor, 0 = automatic */
    const void* x;
    const void* dy;
    const void* dy2;
    void* dw;                    /* fp32 [cout][cin][5][5], overwritten */
    void* dbias;                 /* fp32 [cout], overwritten (may be NULL) */
    void* workspace;             /* >= cvae_conv_wgrad_workspace_bytes(d) */
} cvae_wgrad_desc;

int64_t cvae_conv_wgrad_workspace_bytes(const cvae_wgrad_desc* d);
int cvae_conv_wgrad(const cvae_wgrad_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CVAE_H_ */

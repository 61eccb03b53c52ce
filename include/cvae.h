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
/* Kernels launched by this library in this process so far (bench.py reports the per-step count). */
int64_t cvae_launch_count(void);
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
enum cvae_ktab {
    CVAE_KTAB_GENERIC = 0,
    CVAE_KTAB_PAIR8 = 1,   /* 8-channel source, two taps per K step */
    CVAE_KTAB_BLOCK64 = 2  /* weights packed block-major (CVAE_PACK_KORDER_BLOCK64 / _BLOCK32, optionally CVAE_PACK_STACKx): run the
                              weights-as-A kernel (conv_wa.cu).  src_channels % 32 == 0 (S2D: 4 x 64 or 4 x 32), n_total x stack in
                              {128, 256}, loader NHWC or S2D, any epilogue except PHASE_BIAS_TANH (stack > 1: STATS, MASK, PLAIN).
                              Same results as the GENERIC path. */
};

typedef struct {
    int32_t batch, height, width;
    int32_t ksize;         /* 5 or 3 */
    int32_t src_channels;  /* A-operand channels after the loader transform, multiple of 8 */
    int32_t n_total;       /* GEMM N, multiple of 16 */
    int32_t loader, epilogue, ktab;
    int32_t tm;            /* 128-pixel tiles per CTA pass; 0 = automatic */
    int32_t n_block;       /* GEMM columns per work item (16/32/64/128, divides min(n_total,128)); 0 = automatic */
    const void* src;
    const void* src2;
    const void* wpack;     /* packed by cvae_pack_weights */
    const float* bias;
    const void* act;
    void* out;
    double* stats;         /* [2][n_total], accumulated with atomics (zero it first) */
    /* CVAE_KTAB_BLOCK64 only (zero otherwise) */
    int32_t stack;         /* tap stacking factor the weights were packed with: 0/1, 2 (n_total = 64) or 4 (n_total = 32) */
    int32_t reserved;
    void* workspace;       /* split-K scratch of >= cvae_conv_gemm_workspace_bytes(d) bytes, ZERO-INITIALISED once by the caller
                              (the kernel leaves its arrival counters at zero); may be NULL when that query returns 0 */
    int64_t workspace_bytes;
} cvae_conv_desc;

int cvae_conv_gemm(const cvae_conv_desc* d, void* stream);
/* bytes of cvae_conv_desc.workspace this descriptor needs (0 for every GENERIC / PAIR8 call and for weights-as-A
 * calls that do not split K; pointers in the descriptor are not read); negative on an unsupported shape */
int64_t cvae_conv_gemm_workspace_bytes(const cvae_conv_desc* d);
/* tuning / test hook for the weights-as-A kernel: force cluster size (1, 2, 4), grid size, (block, tap group) units
 * per weight stage, a minimum number of tiles per CTA, the K split (1, 2, 4) and the weight fetch (1: 1-D bulk copies,
 * 2: 2-D tensor-map boxes); 0 = automatic.  Process-wide. */
void cvae_conv_wa_tune(int cluster, int grid, int units_per_stage, int tiles_per_cta, int ksplit, int weight_load);
/* profiling aid: device buffer of >= 8 * 148 uint64 cycle counters of the weights-as-A kernel's MMA thread (NULL = off) */
void cvae_conv_wa_debug_counters(void* device_buf);
/* profiling aid: device buffer of >= 8 * 148 uint64 cycle counters written by the pipelined kernel (NULL = off) */
void cvae_conv_debug_counters(void* device_buf);
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
    int32_t splits;              /* split-K factor, 0 = automatic (a hint: the TMA-fed variant always sizes its own) */
    const void* x;
    const void* dy;
    const void* dy2;
    void* dw;                    /* fp32 [cout][cin][5][5], overwritten */
    void* dbias;                 /* fp32 [cout], overwritten (may be NULL) */
    void* workspace;             /* >= cvae_conv_wgrad_workspace_bytes(d) */
    void* fold_stream;           /* optional second stream for the partial-sum fold (NULL: same stream); the fold is
                                    ordered after the GEMM with an event, the caller joins fold_stream itself */
    int64_t workspace_bytes;     /* size of `workspace`; the call fails with CVAE_EINVAL instead of writing past it
                                    (0 = unchecked) */
} cvae_wgrad_desc;

/* Bytes of split-K workspace cvae_conv_wgrad needs for this shape.  Only kind / batch / height / width / cout / cin /
 * splits are read; the figure covers the call with and without a bias gradient. */
int64_t cvae_conv_wgrad_workspace_bytes(const cvae_wgrad_desc* d);
/* profiling aid: device buffer of >= 8 * (CTAs of the launch) uint64 cycle counters (NULL = off) */
void cvae_wgrad_debug_counters(void* device_buf);
int cvae_conv_wgrad(const cvae_wgrad_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Weight packing (fp32 master parameters in the reference's layouts -> kernel operand layouts).
 * One launch for all layers; no reference counterpart (cuDNN/MKL-DNN pack internally).
 * ---------------------------------------------------------------------------------------------- */
#define CVAE_MAX_PACK_JOBS 24
enum cvae_pack_kind {
    CVAE_PACK_FWD5 = 0,        /* conv forward: n = cout, k_channels = cin, 25 taps                  */
    CVAE_PACK_DGRAD5 = 1,      /* conv data-gradient: n = cin, k_channels = cout, flipped taps       */
    CVAE_PACK_PAIR8 = 2,       /* encoder conv 0 (3 -> 32), 13 K steps of two taps                   */
    CVAE_PACK_PHASE_FWD = 3,   /* conv on 2x up-sampled input as 3x3 conv with n = 4*cout (padded)   */
    CVAE_PACK_PHASE_DGRAD = 4, /* its data-gradient: n = cin, k_channels = 4*cout (padded to 16)     */
    CVAE_PACK_FC = 5,          /* src = fc_mu.weight, src2 = fc_var.weight -> fp32 [4096 nhwc][64]    */
    CVAE_PACK_DECIN = 6        /* src = decoder_input.weight, src2 = bias -> fp32 [34][4096 nhwc]     */
};
/* OR-ed into `kind` of the conv forms for the weights-as-A kernel (CVAE_KTAB_BLOCK64):
 *   KORDER_BLOCK64 / _BLOCK32: K steps ordered (64- / 32-channel block, tap group, 16-channel group) instead of
 *     (tap, 16-channel group); k_channels a multiple of 64 / 32;
 *   STACK2 / STACK4: tap stacking -- GEMM row r of a 128-row block is (channel r / J, j = r % J) and carries tap
 *     dx = s - j of the group (csrc/wa_groups.cuh); `n` is then J x the layer's GEMM width (a multiple of 128). */
#define CVAE_PACK_KORDER_BLOCK64 0x100
#define CVAE_PACK_STACK2 0x200
#define CVAE_PACK_STACK4 0x400
#define CVAE_PACK_KORDER_BLOCK32 0x800
typedef struct {
    int32_t kind, n, ksteps, k_channels, cout, cin;
    const void* src;
    const void* src2;
    void* dst;
} cvae_pack_job;
int64_t cvae_pack_elems(const cvae_pack_job* job);
int cvae_pack_weights(const cvae_pack_job* jobs, int count, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm2d + MaxPool2d(2) + ReLU/Tanh (vae_nets.py:70-72,75-77,80-82,85-87).
 * scale_shift is fp32 [4][C] = scale | shift | mean | invstd, written by cvae_bn_finalize.
 * `stats` ([2][C] double: sum, sum of squares of the bias-free conv output) come from
 * CVAE_EPI_STATS.  Training mode also updates running_mean / running_var / num_batches_tracked
 * with nn.BatchNorm2d's defaults (momentum 0.1, unbiased running variance).  act: 0 ReLU, 1 Tanh.
 * ---------------------------------------------------------------------------------------------- */
int cvae_bn_finalize(int channels, int64_t count, int training, const double* stats, const float* gamma,
                     const float* beta, const float* conv_bias, float* running_mean, float* running_var,
                     int64_t* num_batches_tracked, float momentum, float eps, float* scale_shift, void* stream);
/* xhat_max (bf16 [B][H/2][W/2][C]: normalised value at the arg-max position) and argmax (uint16
 * [B][H/2][W/2][C/8]: 2 bits per channel) are what the backward needs per pooled element; pass both or
 * neither (NULL, NULL in evaluation). */
int cvae_bn_pool_act_fwd(int batch, int height, int width, int channels, int act, const void* conv_out,
                         const float* scale_shift, void* out, void* xhat_max, void* argmax, void* stream);
/* cvae_bn_finalize + cvae_bn_pool_act_fwd in one launch (count = batch * height * width; channels <= 256): every CTA
 * derives the scale / shift itself, block 0 publishes scale_shift and updates the running buffers. */
int cvae_bn_fwd(int batch, int height, int width, int channels, int act, int training, const void* conv_out,
                const double* stats, const float* gamma, const float* beta, const float* conv_bias, float* running_mean,
                float* running_var, int64_t* num_batches_tracked, float momentum, float eps, float* scale_shift, void* out,
                void* xhat_max, void* argmax, void* stream);
/* conv_out bf16 [B][H][W][C]; act_out, d_act, xhat_max bf16 [B][H/2][W/2][C]; argmax as above; sums: [2][C]
 * double scratch; d_conv bf16 [B][H][W][C]; dgamma, dbeta fp32 [C] (overwritten). */
int cvae_bn_pool_act_bwd(int batch, int height, int width, int channels, int act, const void* conv_out,
                         const void* act_out, const void* d_act, const void* xhat_max, const void* argmax,
                         const float* scale_shift, const float* gamma, double* sums, void* d_conv, float* dgamma,
                         float* dbeta, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Linear layers: fc_mu || fc_var (vae_nets.py:98-99,105-109) and decoder_input (:137,143-144).
 * act: bf16 NHWC-flattened bottleneck [B][4096]; mu_logvar fp32 [B][64] (mu | logvar);
 * wfc / wdec are the CVAE_PACK_FC / CVAE_PACK_DECIN outputs; gradients are written in the
 * reference's parameter layouts.  In the backward calls either half may be skipped by passing NULL
 * for its outputs (d_act / d_z_pred = data gradient; dw*, db* = parameter gradients), so the two
 * halves can run on different streams.
 * ---------------------------------------------------------------------------------------------- */
int cvae_fc_fwd(int batch, const void* act, const float* wfc, const float* bias_mu, const float* bias_var,
                float* mu_logvar, void* stream);
int cvae_fc_bwd(int batch, const float* d_mu_logvar, const void* act, const float* wfc, void* d_act,
                float* dw_mu, float* dw_var, float* db_mu, float* db_var, void* stream);
int cvae_decin_fwd(int batch, const float* z_pred, const float* wdec, void* out, void* stream);
int cvae_decin_bwd(int batch, const void* d_out, const float* z_pred, const float* wdec, float* d_z_pred,
                   float* dw, float* db, void* stream);

/* The three layers around the latent in ONE launch per direction (training path; vae_nets.py:108-111, :48-51, :143-144
 * and their data gradients).  Bit-identical to cvae_fc_fwd -> cvae_latent_fwd(sample = 1) -> cvae_decin_fwd, and to
 * cvae_decin_bwd (data) -> cvae_latent_bwd (no external mu / logvar gradients) -> cvae_fc_bwd (data).  The weight
 * gradients of the two layers stay with cvae_decin_bwd / cvae_fc_bwd (d_z_pred may be NULL). */
int cvae_bottleneck_fwd(int batch, const void* act, const float* wfc, const float* bias_mu, const float* bias_var,
                        const float* eps, const float* pred, const float* wdec, float* mu_logvar, float* z_pred,
                        void* dec_in, void* stream);
/* profiling aid: device buffer of 16 + 4 * CTAs int64; CTA 0 of the two kernels stamps clock64() at its phase boundaries,
 * every CTA %globaltimer at entry and exit (NULL = off) */
int cvae_bottleneck_debug(void* device_buf16);
/* profiling aid: 8-CTA clusters of the forward (0) / backward (1) kernel the device holds at once with `smem` dynamic
 * bytes per CTA (0: the kernel's own) */
int cvae_bottleneck_max_clusters(int which, int64_t smem);
int cvae_bottleneck_bwd(int batch, const void* d_dec_in, const float* wdec, const float* mu_logvar, const float* eps,
                        float kld_grad_scale, const float* wfc, float* d_z_pred, float* d_mu_logvar, void* d_act,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Latent: z = mu + eps * exp(0.5 logvar) (vae_nets.py:48-51; eps supplied by the host for parity,
 * sample = 0 decodes the mean as evaluate() does, :43-44) and the critic-value concat (:143):
 * z_pred fp32 [B][33] = z | pred.  The same pass reduces the KL term of vae_loss (vae_nets.py:57-58,
 * sum(1 + logvar - mu^2 - exp(logvar))) into kld_partial: one double per 64 rows
 * (cvae_latent_kld_partials(batch) of them; NULL = skip), which cvae_loss_fwd can take instead of
 * re-reading mu / logvar.  Backward adds the loss's direct gradients on mu / logvar (dmu_ext /
 * dlogvar_ext, may be NULL) and, when kld_grad_scale != 0, the KL term's own backward
 * (kld_grad_scale = kld_weight / batch x upstream gradient).  All tensors 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
int cvae_latent_kld_partials(int batch);
int cvae_latent_fwd(int batch, int sample, const float* mu_logvar, const float* eps, const float* pred,
                    float* z_pred, double* kld_partial, void* stream);
int cvae_latent_bwd(int batch, const float* mu_logvar, const float* eps, const float* d_z_pred,
                    const float* dmu_ext, const float* dlogvar_ext, float kld_grad_scale, float* d_mu_logvar,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * vae_loss (vae_nets.py:53-62): MS-SSIM (vae_nets.py:150-247, including the upstream window sign)
 * + KLD * kld_weight.  recon, x: fp32 NCHW [B][3][64][64]; window11: HOST pointer to the 11
 * normalised 1-D window weights; sums: [10] double scratch; coef: [8] float scratch carried to the
 * backward; losses: [3] = total, recon, KLD.  grad_out: device scalar (NULL = 1).
 * kld_partial: the per-64-row KL partial sums of cvae_latent_fwd for the same mu_logvar (NULL: the
 * KL term is reduced from mu_logvar here).  cvae_loss_bwd: d_mu / d_logvar may both be NULL when the
 * caller folds the KL backward into cvae_latent_bwd.
 * ---------------------------------------------------------------------------------------------- */
int cvae_loss_fwd(int batch, const float* recon, const float* x, const float* mu_logvar, const double* kld_partial,
                  const float* window11, float kld_weight, double* sums, float* coef, float* losses, void* stream);
int cvae_loss_bwd(int batch, const float* recon, const float* x, const float* mu_logvar, const float* window11,
                  float kld_weight, const float* coef, const float* grad_out, float* d_recon, float* d_mu,
                  float* d_logvar, void* stream);
/* cvae_loss_fwd in two calls (level sums, then the scalars + coefficients), and the reconstruction-loss gradient straight from
 * the sums: a training step runs cvae_loss_finalize (one block) on another stream beside cvae_loss_bwd_sums.  Same results. */
int cvae_loss_sums(int batch, const float* recon, const float* x, const float* window11, double* sums, void* stream);
int cvae_loss_finalize(int batch, const float* mu_logvar, const double* kld_partial, const double* sums, float kld_weight,
                       float* coef, float* losses, void* stream);
int cvae_loss_bwd_sums(int batch, const float* recon, const float* x, const float* window11, const double* sums,
                       const float* grad_out, float* d_recon, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Adam (torch.optim.Adam defaults, vae.py:36,58) on flat fp32 buffers; `step` is a device int64
 * counter (incremented by the call) so the launch can be replayed from a CUDA graph.
 * grad_scale multiplies the gradient first (1/world_size after a summing all-reduce).
 * ---------------------------------------------------------------------------------------------- */
int cvae_adam_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                   int64_t* step, float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);
/* The same update for a RANGE of the flat buffers (the four pointers offset alike, 16-byte aligned); `tick` != 0 advances the
 * device step counter afterwards.  One optimizer step may be applied range by range: tick only with the last one. */
int cvae_adam_update(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t* step,
                     float lr, float beta1, float beta2, float eps, float grad_scale, int tick, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Data-parallel training (the reference is single-GPU): NCCL all-reduce of the flat
 * gradient behind the C ABI, one communicator per process = per GPU, bound to libnccl.so.2 at run
 * time.  Rank 0 creates 128 id bytes, the host ships them to every rank (any transport), every rank
 * calls cvae_comm_init (a collective), then cvae_comm_allreduce_sum sums `count` floats in place,
 * asynchronously on `stream`; cvae_adam_step's grad_scale = 1 / world turns the sum into the mean.
 * ---------------------------------------------------------------------------------------------- */
int cvae_comm_unique_id(void* id128);
int cvae_comm_init(int rank, int world, const void* id128);
int cvae_comm_world(void);
int cvae_comm_allreduce_sum(float* buf, int64_t count, void* stream);
int cvae_comm_destroy(void);

/* uint8 HWC frames [N][64][64][3] -> fp32 NCHW [N][3][64][64] = astype(float32) / 255, the
 * preprocessing of vae_utility.py:324-343 (adjust_values + HWC->CHW) done on the device. */
int cvae_frames_u8_to_f32(int frames, const uint8_t* hwc_u8, float* nchw_f32, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Critic forward (critic_net.py:15-42,66-69).  x fp32 NCHW [N][3][64][64] in [0,1]; weights: the 14
 * state_dict tensors concatenated in key order; pred fp32 [N].
 * ---------------------------------------------------------------------------------------------- */
int cvae_critic_param_count(void);
int cvae_critic_fwd(int frames, const float* x, const float* weights, float* pred, void* stream);

/* ------------------------------------------------------------------------------------------------
 * -video mask pipeline.  cvae_diff_grey: vae_utility.py:270-275 (recon fp32 [N][3][64][64] ->
 * diff fp64 [N][64][64], per-frame max).  cvae_mask_iou: vae_utility.py:279-284,153-157 and the
 * integer counts of :57-59 for a list of thresholds at once: diff_u8 / mask (for `thr`) are
 * optional outputs, hist512 = [gt 0|1][value] uint64 scratch/output, counts int64 [nthr][3] =
 * tp, fn, fp.  mean_max / diff_factor are the host scalars of vae_utility.py:106-110.
 * ---------------------------------------------------------------------------------------------- */
int cvae_diff_grey(int frames, const float* recon_hi, const float* recon_lo, double* diff, double* max_values,
                   void* stream);
int cvae_mask_iou(int frames, const double* diff, const uint8_t* gt, double mean_max, double diff_factor, int thr,
                  int nthr, const int* thr_list, uint8_t* diff_u8, uint8_t* mask, uint64_t* hist512,
                  int64_t* counts, void* stream);
/* get_iou's integer part (vae_utility.py:57-59) for two boolean (uint8 0/1) arrays of n elements:
 * counts3 = tp, fn, fp.  The division and round(.,3) stay in Python. */
int cvae_iou_counts(int64_t n, const uint8_t* gt, const uint8_t* mask, int64_t* counts3, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CVAE_H_ */

#!/bin/bash
# compute-sanitizer over the kernel tests at their small shapes.  ONE tool per GPU call
# (B200_PROFILING.md: several tools in one call have left a GPU unusable): memcheck | racecheck | synccheck | initcheck.
#   gpurun -- bash tools/sanitize.sh memcheck        -> gpurun_out/sanitizer_memcheck.log (+ a one-screen summary)
# The selection covers every kernel family once: pixels-as-M conv (forward, data gradient, frames loader), weights-as-A
# conv (plain, cluster multicast, split-K, tap stacking, 64-byte-swizzled tiles), both weight-gradient kernels + folds,
# BatchNorm/pool, linear (cluster/DSMEM), latent, MS-SSIM, Adam, critic, mask pipeline.
TOOL=${1:-memcheck}
OUT=gpurun_out/sanitizer_$TOOL.log
mkdir -p gpurun_out
SEL='test_encoder_conv_with_stats[3-32-64-32-0] or test_encoder_conv0_from_nchw_frames[1] or test_dgrad_5x5[5-256-128-4] or test_decoder_upsample_folded_conv[3-64-32-8] or test_dgrad_last_conv_from_nchw_grad or test_wa_encoder_conv_with_stats[5-64-128-16-tune0] or test_wa_encoder_conv_with_stats[5-64-128-16-tune2] or test_wa_decoder_conv0_bias_relu[tune6] or test_wa_decoder_conv0_bias_relu[tune8] or test_wa_stacked_dgrad_5x5[3-64-128-16-tune0] or test_wa_stacked_dgrad_5x5[2-32-64-32-tune0] or test_wa_stacked_encoder_conv1_with_stats[3-tune0] or test_wa_stacked_dgrad_upsample_folded_with_relu_mask[2-64-32-8-tune0] or test_wa_dgrad_upsample_folded_with_relu_mask[3-tune0]'
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 --error-exitcode 86 \
    python -m pytest tests/test_conv_gemm.py tests/test_conv_wgrad.py tests/test_kernels_misc.py -m gpu -q -x -p no:cacheprovider \
    -k "$SEL or test_wgrad or test_bn_pool or test_linear or test_fused_latent_kld_kernels[5] or test_loss_forward or test_adam or test_critic or test_mask" \
    > $OUT 2>&1
RC=$?
echo "compute-sanitizer --tool $TOOL exit code $RC" | tee -a $OUT
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|error" $OUT | tail -8
grep -E "^=========" $OUT | grep -v "COMPUTE-SANITIZER" | head -40
exit 0

#!/bin/bash
# compute-sanitizer over one test per kernel family (ONE tool per gpurun call: B200_PROFILING.md).
#   tools/sanitize.sh memcheck|racecheck|synccheck|initcheck   -> gpurun_out/sanitizer_<tool>.log
# Families: CTA-pair NT GEMM + split-K TN GEMM + every elementwise backward kernel (stage test), cnet_fused (bit-equality
# test), pconv_coupling incl. band mode (32x32 maps), made_inverse push / pull, flow1d fwd / bwd, Split2d / loss / Adam
# (KD step golden).
set -u
TOOL=${1:-memcheck}
OUT=gpurun_out/sanitizer_${TOOL}.log
SEL='test_every_kernel_of_a_flowstep_is_exact_on_its_own_inputs and 12-16-40
 or test_cnet_fused_kernel_is_bit_identical_to_the_two_gemms and 8200-128
 or test_fused_conv3_coupling_matches_two_kernel_path_and_torch and (3-12-32-False or 7-24-8-False or 37-48-4-True)
 or test_resident_inverse_matches_paper_restatement
 or test_maf_forward_inverse_backward_vs_paper_restatement and 63-512-3-257
 or test_ragged_batch_gradients_match_oracle_both_directions and 63-16-2-1003
 or test_kd_training_step_golden'
timeout 1500 compute-sanitizer --tool "$TOOL" --error-exitcode 9 --print-limit 20 \
  python -m pytest tests -m gpu -x -q -k "$(echo $SEL)" > "$OUT" 2>&1
echo "exit code $?" >> "$OUT"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code" "$OUT" | tail -5

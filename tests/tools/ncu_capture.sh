#!/bin/bash
# `ncu --set full` captures of the kernels DESIGN.md section 6 discusses, each only after the same command has exited 0
# without ncu. Run on the GPU box:  bash tests/tools/ncu_capture.sh [tag]   -> gpurun_out/<tag>_*.ncu-rep + CSV summaries
# (copy the CSVs you want judged into profiles/). Optional second argument: which capture (conv16 | small | new | m3ae | all).
tag=${1:-r2}
which=${2:-all}
NCU="ncu --set full --clock-control none --import-source on -f"
run() {  # name, kernel regex, launch count, command...
  local name=$1 rx=$2 cnt=$3; shift 3
  timeout 300 "$@" > gpurun_out/${tag}_${name}_plain.log 2>&1 || { echo "$name: plain run failed"; return; }
  timeout 900 $NCU -k "regex:$rx" -c $cnt -o gpurun_out/${tag}_${name} "$@" > gpurun_out/${tag}_${name}_ncu.log 2>&1
  python tests/tools/ncu_summary.py gpurun_out/${tag}_${name}.ncu-rep gpurun_out/${tag}_${name}_ncu_full.csv "${EXTRA[@]}" \
    >> gpurun_out/${tag}_${name}_ncu.log 2>&1
  rm -f gpurun_out/${tag}_${name}.ncu-rep          # the merged-back directory is capped at 64 MiB: keep the summaries only
  echo "$name: $(wc -l < gpurun_out/${tag}_${name}_ncu_full.csv) rows"
}
EXTRA=()
want() { [ "$which" = all ] || [ "$which" = "$1" ]; }
want conv16 && run conv16 "conv16_persistent|conv_gemm_kernel|conv_strip16" 27 python tests/tools/profile_conv16.py
want small && run small "gs_project|head_rows|head_cols|head_reduce|fuse_eval" 12 python tests/tools/profile_top.py
want new && run new "head_gemm|head_softmax|head_split|head_db|frame_|spec_augment" 16 python tests/tools/profile_top.py
want m3ae && run m3ae "linear_gemm_persistent|attn_fwd_tc|attn_bwd|conv_gemm_kernel" 24 python tests/tools/profile_m3ae.py 32 257
true

#!/bin/bash
# One GPU-box session: [tests] [variants] [bench] [profile]; everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for what in "$@"; do
  case "$what" in
    tests)
      for f in tests/test_*_gpu.py; do
        name=$(basename "$f" .py)
        timeout 900 python -m pytest "$f" -m gpu -q -x --tb=short > "gpurun_out/$name.log" 2>&1
        echo "== $f -> exit $?"; tail -n 3 "gpurun_out/$name.log"
      done ;;
    tests_quick)
      timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_attention_gpu.py -m gpu -q --tb=short > gpurun_out/quick.log 2>&1
      echo "== quick tests -> exit $?"; grep -E "passed|failed|Error" gpurun_out/quick.log | tail -n 30 ;;
    variants)
      : > gpurun_out/variants.log
      for grp in attn xattn geglu gn ln upconv pdl; do
        timeout 300 python scripts/bench_variants.py $grp > gpurun_out/variants_$grp.log 2>&1
        echo "== variants $grp -> exit $?"; cp gpurun_out/variants.json gpurun_out/variants_$grp.json 2>/dev/null
        grep -E "us|Error|error" gpurun_out/variants_$grp.log | tail -n 30
      done ;;
    bench)
      timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
      echo "== bench -> exit $?"; tail -c 4000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err ;;
    bench_fast)
      timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
      echo "== bench -> exit $?"; tail -c 4000 gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err ;;
    variants_sel)
      for grp in $VARIANT_GROUPS; do
        timeout 300 python scripts/bench_variants.py $grp > gpurun_out/variants_$grp.log 2>&1
        echo "== variants $grp -> exit $?"; cp gpurun_out/variants.json gpurun_out/variants_$grp.json 2>/dev/null
        grep -E " us|Error|error" gpurun_out/variants_$grp.log | tail -n 80
      done ;;
    ncu_gn)
      ncu --set full --clock-control none -k regex:gn_ -c 12 -o gpurun_out/prof_gn -f python scripts/gn_only.py > gpurun_out/ncu_gn.log 2>&1
      echo "== ncu gn -> exit $?"
      ncu -i gpurun_out/prof_gn.ncu-rep --page raw --csv > gpurun_out/prof_gn_raw.csv 2>/dev/null ;;
    ncu_attn)
      for v in $ATT_VARIANTS; do
        ncu --set full --import-source on --clock-control none -k regex:attention2 -s 1 -c 1 -o gpurun_out/prof_attn_v$v -f python scripts/attn_only.py $v > gpurun_out/ncu_attn_$v.log 2>&1
        echo "== ncu attn v$v -> exit $?"
        ncu -i gpurun_out/prof_attn_v$v.ncu-rep --page source --csv > gpurun_out/prof_attn_v${v}_source.csv 2>/dev/null
        ncu -i gpurun_out/prof_attn_v$v.ncu-rep --page raw --csv > gpurun_out/prof_attn_v${v}_raw.csv 2>/dev/null
      done ;;
    ncu_gemm)
      ncu --set full --import-source on --clock-control none -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof_gemm1 -f python scripts/gemm_only.py $GEMM_SHAPE > gpurun_out/ncu_gemm1.log 2>&1
      echo "== ncu gemm -> exit $?"
      ncu -i gpurun_out/prof_gemm1.ncu-rep --page source --csv > gpurun_out/prof_gemm1_source.csv 2>/dev/null
      ncu -i gpurun_out/prof_gemm1.ncu-rep --page raw --csv > gpurun_out/prof_gemm1_raw.csv 2>/dev/null ;;
    shapes)
      timeout 600 python scripts/bench_kernels.py > gpurun_out/shapes.log 2>&1
      echo "== shapes -> exit $?"; tail -n 12 gpurun_out/shapes.log ;;
    profile)
      bash scripts/gpu_profile.sh ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
      echo "== smoke -> exit $?"; tail -n 5 gpurun_out/smoke.log ;;
  esac
done

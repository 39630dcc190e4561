# A/B of compile-time variants (each argument is a set of nvcc flags, "" = as shipped): build, bench (per-kernel times
# from its frame-after-frame pass), parity tests.  Report: python profiles/ab_report.py
i=0
for v in "$@"; do
  i=$((i+1)); tag=v$i
  echo "$tag: $v" >> gpurun_out/ab_tags.txt
  EVX_EXTRA_NVCC="$v" python -m cairo_b200.build --force > gpurun_out/ab_build_$tag.log 2>&1
  python bench.py --steps 24 --warmup 4 > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_configs.py tests/test_gpu_bins.py -m gpu -x -q 2>&1 | tail -2 > gpurun_out/ab_test_$tag.log
done

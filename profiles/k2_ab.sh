for v in "" "-DEVX_K2_BOTH" "-DEVX_K2_BOTH -DEVX_K2_SUBBOTH" "-DEVX_K2_SUBBOTH"; do
  tag=$(echo "$v" | tr -d ' -' ); tag=${tag:-base}
  EVX_EXTRA_NVCC="$v" python -m cairo_b200.build --force > gpurun_out/ab_build_$tag.log 2>&1
  python bench.py --steps 24 --warmup 4 --streams 1 > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -2 > gpurun_out/ab_test_$tag.log
done

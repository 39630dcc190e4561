#!/bin/bash
# A/B of prebuilt library variants (profiles/_variants/libevxgpu_<tag>.so, built here with EVX_EXTRA_NVCC): the frame
# pipeline alone at several slot counts, one frame after the other, and 16 streams.  bash profiles/variants_pipe.sh tag...
cp cairo_b200/libevxgpu.so /tmp/libevxgpu_orig.so
for tag in "$@"; do
  cp profiles/_variants/libevxgpu_$tag.so cairo_b200/libevxgpu.so
  echo "== $tag"
  for s in ${SLOTS:-1 6 8 12}; do EVXGPU_FRAME_SLOTS=$s python profiles/pipe_bench.py 240 2>&1 | tail -1; done
  if [ -n "$KTIMES" ]; then python profiles/kernel_times.py 2 2>&1 | tail -1; python profiles/kernel_times.py 4 12 2>&1 | tail -1; fi
  if [ -n "$STREAMS" ]; then EVXGPU_FRAME_SLOTS=1 python profiles/pipe_bench.py 60 2 1920 1080 $STREAMS 2>&1 | tail -1; fi
done
cp /tmp/libevxgpu_orig.so cairo_b200/libevxgpu.so

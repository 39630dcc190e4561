# three frame slots against two: parity tests, stress, bench at a few band sizes / lookaheads
(time python -m pytest tests -m gpu -x -q) > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?"
EVXGPU_FRAME_SLOTS=3 timeout 250 python profiles/stress_overlap.py > gpurun_out/s3_stress.log 2>&1; echo "stress rc=$?"
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 56 > gpurun_out/s3_$tag.json 2> gpurun_out/s3_$tag.err; echo "$tag rc=$?"; }
run s2 EVXGPU_FRAME_SLOTS=2
run s3 EVXGPU_FRAME_SLOTS=3 EVX_BENCH_LOOKAHEAD=6
run s3b4 EVXGPU_FRAME_SLOTS=3 EVXGPU_BAND_ROWS=4 EVX_BENCH_LOOKAHEAD=6
run s3b3 EVXGPU_FRAME_SLOTS=3 EVXGPU_BAND_ROWS=3 EVX_BENCH_LOOKAHEAD=6
run s3l4 EVXGPU_FRAME_SLOTS=3 EVX_BENCH_LOOKAHEAD=4

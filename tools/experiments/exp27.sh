# compute-sanitizer over the boundary-kernel tests (new classifier / warp-specialised layout)
set -u
mkdir -p gpurun_out
SEL='segment_sequence_matches or segment_edge or audio_scan_near or batched_streams or pattern_separation'
timeout 200 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_segmentation.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/racecheck_seg.log 2>&1; echo "racecheck rc $?"; grep -E "RACECHECK SUMMARY|passed|failed|Race reported|ERROR SUMMARY" gpurun_out/racecheck_seg.log | tail -5
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_segmentation.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > gpurun_out/memcheck_seg.log 2>&1; echo "memcheck rc $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_seg.log | tail -4

export QB200_FUSED=0
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or windowed_matches or stage1 or packed or known or hirschberg_splits or text" 2>&1 | tail -15 > gpurun_out/r2_memcheck.txt
echo "exit $?" >> gpurun_out/r2_memcheck.txt
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or stage1_compact or packed" 2>&1 | tail -15 > gpurun_out/r2_racecheck.txt
tail -5 gpurun_out/r2_memcheck.txt gpurun_out/r2_racecheck.txt

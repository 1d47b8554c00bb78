mkdir -p gpurun_out
(echo "fuzz_parity full (2448x2048, random parameter sets):"; timeout 900 python tools/fuzz_parity.py 8 101 full 2>&1 | tail -1
echo "fuzz_parity default:"; timeout 600 python tools/fuzz_parity.py 400 102 2>&1 | tail -1
echo "fuzz_bm:"; timeout 400 python tools/fuzz_bm.py 300 7 2>&1 | tail -1) | tee gpurun_out/r2_fuzz3.txt

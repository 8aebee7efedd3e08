#!/bin/bash
# Round-end evidence in one GPU call (not a test):  tests/final_capture.sh <out_dir>
#   full GPU test suite, smoke(), the bench line, the ncu launch list of a bench step and one `--set full` capture of the tensor-core kernels.
out=${1:-gpurun_out/final}
mkdir -p "$out"
python -m pytest tests -m gpu -q > "$out/pytest.log" 2>&1; tail -3 "$out/pytest.log"
python -c "import __graft_entry__ as g; g.smoke()" > "$out/smoke.log" 2>&1; tail -2 "$out/smoke.log"
python bench.py > "$out/bench.json" 2> "$out/bench.err"; cut -c1-400 "$out/bench.json"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/launches.csv" python bench.py --ncu --steps 2 --warmup 1 > "$out/ncu_l.log" 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sc2_fused|tc_gemm2" -s 11 -c 11 -f -o "$out/prof_tc" python bench.py --ncu --steps 2 --warmup 2 > "$out/ncu_full.log" 2>&1
ls -la "$out"

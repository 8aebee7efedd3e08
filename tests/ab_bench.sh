#!/bin/bash
# Developer A/B harness (not a test): per-kernel-category times of bench.py for several builds of the library, alternating.
#   tests/ab_bench.sh out_dir lib1.so lib2.so ...      ("-" = the main build)
out=$1; shift
mkdir -p "$out"
for rep in 1 2; do
  for lib in "$@"; do
    tag=$(basename "$lib" .so)
    if [ "$lib" = "-" ]; then unset CMF_LIB; tag=main; else export CMF_LIB=$PWD/$lib; fi
    python bench.py --no-cpu-baseline --no-extra-legs --steps 20 --warmup 5 > "$out/bench_${tag}_$rep.json" 2> "$out/bench_${tag}_$rep.err"
    python - "$out/bench_${tag}_$rep.json" "$tag" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[2], round(d["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items() if v["ms_per_step"] > 0})
PY
  done
done

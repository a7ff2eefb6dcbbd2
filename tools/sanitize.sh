#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over one small launch per kernel family.
# Usage (GPU box): bash tools/sanitize.sh [outdir]   -> <outdir>/{memcheck,racecheck,synccheck}.log + summary.txt
OUT=${1:-gpurun_out/sanitizer_r02}
mkdir -p "$OUT"
: > "$OUT/summary.txt"
python tools/sanitizer_cases.py > "$OUT/plain.log" 2>&1 || { echo "plain run failed" >> "$OUT/summary.txt"; tail -5 "$OUT/plain.log" >> "$OUT/summary.txt"; exit 1; }
for tool in memcheck synccheck racecheck; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 40 --log-file "$OUT/$tool.log" \
      python tools/sanitizer_cases.py > "$OUT/$tool.stdout.log" 2>&1
  rc=$?
  echo "== $tool: exit $rc; $(grep -c 'cases done' "$OUT/$tool.stdout.log") completed run(s)" >> "$OUT/summary.txt"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$OUT/$tool.log" >> "$OUT/summary.txt"
done
cat "$OUT/summary.txt"

# Round-2 evidence batch (GPU box): sanitizer runs of a small end-to-end stamp, the other BASELINE configurations on one GPU,
# and a --set full capture of the two-warpgroup flash kernel at the level-0 shape.
set -x
cd $GRAFT_REPO_ROOT
export_rep() {
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass > gpurun_out/$1_source.csv 2>/dev/null
  gzip -f gpurun_out/$1_source.csv
  rm -f gpurun_out/$1.ncu-rep
}
R=64 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python profiles/sanitize_stamp.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
R=64 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python profiles/sanitize_stamp.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
R=64 timeout 900 compute-sanitizer --tool synccheck --error-exitcode 3 python profiles/sanitize_stamp.py > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/sanitizer_synccheck.log
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_c3_1gpu.json 2> gpurun_out/bench_r2_c3_1gpu.err; echo "c3 rc=$?"
timeout 600 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/bench_r2_c4.json 2> gpurun_out/bench_r2_c4.err; echo "c4 rc=$?"
timeout 900 python bench.py --config c5 > gpurun_out/bench_r2_c5.json 2> gpurun_out/bench_r2_c5.err; echo "c5 rc=$?"
timeout 600 python bench.py --resolution 256 --denoise-steps 20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_256.json 2> gpurun_out/bench_r2_256.err; echo "256 rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:flash_attn2 -s 3 -c 1 -o gpurun_out/prof_r2_flash -f python profiles/flash_bench.py 5 > gpurun_out/ncu_flash.log 2>&1; echo rc=$?
export_rep prof_r2_flash
du -sh gpurun_out; ls gpurun_out

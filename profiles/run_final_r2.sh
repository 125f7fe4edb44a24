# Final refresh of the round-2 evidence after the last kernel change (GPU box): launch list + DRAM bytes of one C2 stamp,
# sanitizer runs of the small end-to-end stamp, the default bench line.
set -x
cd $GRAFT_REPO_ROOT
NCU="ncu --profile-from-start off --clock-control none"
timeout 1500 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/launches_r2.csv python profiles/profile_stamp.py --no-op-profile --no-graph --profiler-range > gpurun_out/ncu_list_r2.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches_r2.csv; gzip -f gpurun_out/launches_r2.csv
for t in memcheck racecheck synccheck; do
  R=64 timeout 900 compute-sanitizer --tool $t --error-exitcode 3 python profiles/sanitize_stamp.py > gpurun_out/sanitizer_$t.log 2>&1; echo "$t rc=$?"; tail -3 gpurun_out/sanitizer_$t.log
done
for t in memcheck racecheck; do
  FULL=1 R=64 timeout 900 compute-sanitizer --tool $t --error-exitcode 3 python profiles/sanitize_stamp.py > gpurun_out/sanitizer_full_$t.log 2>&1; echo "full $t rc=$?"; tail -3 gpurun_out/sanitizer_full_$t.log
done
python bench.py > gpurun_out/bench_final_r2.json 2> gpurun_out/bench_final_r2.err; echo "bench rc=$?"
python bench.py --resolution 256 --denoise-steps 20 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_256.json 2> gpurun_out/bench_r2_256.err; echo "256 rc=$?"
python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_c3_1gpu.json 2> gpurun_out/bench_r2_c3_1gpu.err; echo "c3 rc=$?"
python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_c4.json 2> gpurun_out/bench_r2_c4.err; echo "c4 rc=$?"

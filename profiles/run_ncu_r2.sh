# Round-2 profiling batch (run on the GPU box through gpurun): launch lists + --set full captures, exported to CSV on the box
# (the .ncu-rep files are large; gpurun merges at most 64 MiB back).
set -x
cd $GRAFT_REPO_ROOT
NCU="ncu --profile-from-start off --clock-control none"
export_rep() {  # $1 = report stem
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass > gpurun_out/$1_source.csv 2>/dev/null
  gzip -f gpurun_out/$1_source.csv
  rm -f gpurun_out/$1.ncu-rep
}
# 1. launch list of one warm C2 stamp: time + DRAM bytes per launch
timeout 1500 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/launches_r2.csv python profiles/profile_stamp.py --no-op-profile --no-graph --profiler-range > gpurun_out/ncu_list_r2.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches_r2.csv
gzip -f gpurun_out/launches_r2.csv
# 2. --set full captures inside a 512x512 stamp
timeout 600 $NCU --set full --import-source on -k regex:gemm_tc_kernel -s 1200 -c 8 -o gpurun_out/prof_r2_gemm512 -f python profiles/profile_stamp.py --no-op-profile --no-graph --profiler-range > gpurun_out/ncu_full1.log 2>&1; echo rc=$?
export_rep prof_r2_gemm512
timeout 600 $NCU --set full --import-source on -k regex:"flash_attn2|gn_fused|gn_group2|cross_attn" -s 40 -c 10 -o gpurun_out/prof_r2_misc512 -f python profiles/profile_stamp.py --no-op-profile --no-graph --profiler-range > gpurun_out/ncu_full2.log 2>&1; echo rc=$?
export_rep prof_r2_misc512
# 3. the server's operating point: 256x256, B=1
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_r2_256.csv python profiles/profile_stamp.py --resolution 256 --no-op-profile --no-graph --profiler-range > gpurun_out/ncu_list_r2_256.log 2>&1; echo rc=$?; gzip -f gpurun_out/launches_r2_256.csv
timeout 600 $NCU --set full --import-source on -k regex:"gemm_tc_kernel|flash_attn" -s 150 -c 8 -o gpurun_out/prof_r2_256 -f python profiles/profile_stamp.py --resolution 256 --no-op-profile --no-graph --profiler-range > gpurun_out/ncu_full3.log 2>&1; echo rc=$?
export_rep prof_r2_256
ls -la gpurun_out | tail -20; du -sh gpurun_out

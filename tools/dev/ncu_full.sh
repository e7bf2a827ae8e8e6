set -e
python tools/dev/prof_edge.py f16 > gpurun_out/prof_plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:_tc_kernel -c 12 -o gpurun_out/prof_r01b -f python tools/dev/prof_edge.py f16 > gpurun_out/prof_ncu.log 2>&1
tail -2 gpurun_out/prof_ncu.log

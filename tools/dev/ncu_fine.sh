set -e
python tools/dev/prof_edge.py f16 > gpurun_out/prof_plain.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on --warp-sampling-interval 0 -k regex:edge_tc_kernel -s 9 -c 3 -o gpurun_out/prof_fine -f python tools/dev/prof_edge.py f16 > gpurun_out/prof_ncu.log 2>&1
tail -2 gpurun_out/prof_ncu.log

# ncu --set full captures of the kernels beside the fused edge step; leaves small CSV summaries (the .ncu-rep stays on the box)
set -e
ncu --set full --clock-control none -k regex:"k_segment_reduce|k_gather_rows|k_knn_radius|k_knn_merge" \
    --launch-skip 0 -c 14 -o /tmp/r2_other python profiles/other_kernels_once.py > gpurun_out/r2_other_ncu.log 2>&1
ncu -i /tmp/r2_other.ncu-rep --page raw --csv > gpurun_out/r2_other_raw.csv 2>/dev/null
python profiles/ncu_pick.py gpurun_out/r2_other_raw.csv > gpurun_out/r2_other_summary2.txt
rm -f gpurun_out/r2_other_raw.csv

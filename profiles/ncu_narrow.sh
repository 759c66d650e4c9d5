# ncu --set full of the fan-in <= 8 first-layer kernels (profiles/other_kernels_once.py), condensed
set -e
ncu --set full --clock-control none -k regex:"k_narrow_in" --launch-skip 0 -c 6 -o /tmp/r2_narrow python profiles/other_kernels_once.py > gpurun_out/r2_narrow_ncu.log 2>&1
ncu -i /tmp/r2_narrow.ncu-rep --page raw --csv > gpurun_out/r2_narrow_raw.csv 2>/dev/null
python profiles/ncu_pick.py gpurun_out/r2_narrow_raw.csv > gpurun_out/r2_narrow_summary.txt
rm -f gpurun_out/r2_narrow_raw.csv

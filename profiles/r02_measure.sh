# Round-2 evidence on one B200: bench line, ncu launch list of the same command, ncu --set full of the edge-step kernels,
# microbenchmark sweep. Leaves text / json in gpurun_out/ (the .ncu-rep files stay on the box).
set -x
python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --no-e2e --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r02_ncu_launch.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches.md
ncu --set full --clock-control none -k regex:"k_tc_edge_fwd|k_tc_edge_bwd2|k_tc_wgrad|k_img_segment_reduce|k_tc_row_fwd|k_ordered_colsum|k_agg_fixup|k_rows_to_bf16" \
    --launch-skip 36 -c 12 -o /tmp/r02_edge python profiles/edge_step_once.py 1000000 2 full > gpurun_out/r02_ncu_edge.log 2>&1
ncu -i /tmp/r02_edge.ncu-rep --page raw --csv > /tmp/r02_edge_raw.csv 2>/dev/null
python profiles/ncu_pick.py /tmp/r02_edge_raw.csv > gpurun_out/r02_ncu_edge_kernels.txt
python profiles/microbench.py > gpurun_out/r02_microbench.md 2> gpurun_out/r02_microbench.err

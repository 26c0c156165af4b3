#!/bin/bash
# Round-2 evidence run on one B200 (gpurun): the default bench line, the ncu launch list of the bench command on the
# configs[2] cube and ncu --set full captures of the hot kernels.  The .ncu-rep files are summarised on the box (text) and
# removed: gpurun copies back at most 64 MiB.
mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_final_1gpu.json 2> gpurun_out/r2_bench_final_1gpu.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"smooth|csr_stream|sell_spmv|spmv_vector|cg_|dot_kernel|coarse_gemv|gather_kernel|scatter_kernel|restrict_fused" -c 2000 \
    --csv --log-file gpurun_out/r2_launches_raw.csv python bench.py --cube 118 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary \
    > gpurun_out/r2_bench_under_ncu.json 2> gpurun_out/r2_bench_under_ncu.err; echo ncu1 rc=$?
python tools/launch_summary.py gpurun_out/r2_launches_raw.csv > gpurun_out/r2_launch_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"smooth_|sell_spmv|spmv_vector|csr_stream|restrict_fused|coarse_gemv" -s 40 -c 30 \
    -o gpurun_out/r2_prof_full_118 python tools/profile_solve.py --cube 118 --maxiters 4 > gpurun_out/r2_prof_full_118.log 2>&1; echo ncu2 rc=$?
python tools/ncu_summary.py gpurun_out/r2_prof_full_118.ncu-rep l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum > gpurun_out/r2_ncu_full_118.txt 2>&1
python tools/ncu_lsu.py gpurun_out/r2_prof_full_118.ncu-rep smooth_ell_kernel 0 > gpurun_out/r2_smoother_source_118.txt 2>&1
python tools/make_traffic_json.py cube118=gpurun_out/r2_prof_full_118.ncu-rep > /dev/null 2>&1; cp profiles/r2_traffic.json gpurun_out/r2_traffic_118.json
ncu --set full --clock-control none -k regex:"smooth_ell|sell_spmv_kernel" -s 11 -c 14 \
    -o gpurun_out/r2_prof_full_255 python tools/profile_solve.py --cube 255 --maxiters 3 > gpurun_out/r2_prof_full_255.log 2>&1; echo ncu3 rc=$?
python tools/ncu_summary.py gpurun_out/r2_prof_full_255.ncu-rep l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed > gpurun_out/r2_ncu_full_255.txt 2>&1
python tools/make_traffic_json.py cube118=gpurun_out/r2_prof_full_118.ncu-rep cube255=gpurun_out/r2_prof_full_255.ncu-rep > /dev/null 2>&1; cp profiles/r2_traffic.json gpurun_out/r2_traffic.json
rm -f gpurun_out/r2_prof_full_118.ncu-rep gpurun_out/r2_prof_full_255.ncu-rep
du -sh gpurun_out

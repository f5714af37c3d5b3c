set -x
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU --kernel-name regex:assemble_mapped -c 1 -s 1 -o gpurun_out/r02_k2_mapped python tools/prof_k2.py c2_fluxonium_t200 8192 > gpurun_out/p1.log 2>&1
$NCU --kernel-name regex:assemble_entries -c 1 -s 1 -o gpurun_out/r02_k2_entries python tools/prof_k2.py c3_csfq4jj_t10 4096 > gpurun_out/p2.log 2>&1
$NCU --kernel-name regex:blk_her2k_dmma -c 1 -s 20 -o gpurun_out/r02_s2_her2k python tools/prof_run.py c3_csfqfloat_t7_full 3 10 1 0 2 1 > gpurun_out/p3.log 2>&1
$NCU --kernel-name regex:blk_reflect_hemv -c 1 -s 1000 -o gpurun_out/r02_s2_hemv python tools/prof_run.py c3_csfqfloat_t7_full 3 10 1 0 2 1 > gpurun_out/p4.log 2>&1
$NCU --kernel-name regex:mf_cheb_stencil -c 1 -s 2 -o gpurun_out/r02_mf_stencil1 python tools/prof_run.py c3_csfqfloat_t7_full 256 10 1 0 0 1 > gpurun_out/p5.log 2>&1
$NCU --kernel-name regex:s1l_tridiag -c 1 -s 3 -o gpurun_out/r02_s1l_stage2 python tools/prof_run.py c2_fluxonium_res_10k 0 10 1 1 0 2 > gpurun_out/p6.log 2>&1
python bench.py --steps 2 --warmup 1 --no-workloads --no-parallel-baseline --ref-budget 1 > gpurun_out/b_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-workloads --no-parallel-baseline --ref-budget 1 > gpurun_out/b_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail

#!/bin/bash
# Regenerates the round's evidence on a B200 box (run through gpurun): GPU tests, smoke, bench lines,
# ncu launch list of the bench command, one `ncu --set full` capture per hot kernel, per-op microbench.
set -x
python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -12
python bench.py > gpurun_out/bench_r1_lres.json 2> gpurun_out/bench_r1_lres.err; tail -2 gpurun_out/bench_r1_lres.err; cut -c1-200 gpurun_out/bench_r1_lres.json
python bench.py --workload sres --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r1_sres.json 2> gpurun_out/bench_r1_sres.err; cut -c1-200 gpurun_out/bench_r1_sres.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1_bench_lres.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_r1_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:bias_act_vec -s 1 -c 1 -o gpurun_out/r1_bias_fwd python tools/profile_kernels.py bias_act_fwd > /dev/null 2>&1
$NCU -k regex:bias_act_vec -s 2 -c 1 -o gpurun_out/r1_bias_bwd python tools/profile_kernels.py bias_act_bwd > /dev/null 2>&1
$NCU -k regex:upfirdn2d_stream -s 1 -c 1 -o gpurun_out/r1_up2 python tools/profile_kernels.py upfirdn_up2 > /dev/null 2>&1
$NCU -k regex:upfirdn2d_stream -s 1 -c 1 -o gpurun_out/r1_down2 python tools/profile_kernels.py upfirdn_down2 > /dev/null 2>&1
$NCU -k regex:filtered_lrelu_kernel -s 1 -c 1 -o gpurun_out/r1_flrelu python tools/profile_kernels.py flrelu_u2d2 > /dev/null 2>&1
$NCU -k regex:conv_fprop_tc -s 1 -c 1 -o gpurun_out/r1_conv python tools/profile_kernels.py conv_l8 > /dev/null 2>&1
$NCU -k regex:conv_wgrad -s 1 -c 1 -o gpurun_out/r1_wgrad python tools/profile_kernels.py conv_wgrad_l8 > /dev/null 2>&1
python tools/microbench.py > gpurun_out/microbench_r1.txt 2>&1
ls -la gpurun_out

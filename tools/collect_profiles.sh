#!/bin/bash
# Turns what tools/profile_round.sh left in gpurun_out/ into the tracked summaries under profiles/ (run in the repo root).
set -e
R=${1:-r01}
for n in bias_fwd:bias_act_fwd bias_bwd:bias_act_bwd_fused_db up2:upfirdn2d_stream_up2 down2:upfirdn2d_stream_down2 flrelu:filtered_lrelu conv:conv_fprop_tc wgrad:conv_wgrad_tc; do
  src=${n%%:*}; dst=${n##*:}
  [ -f gpurun_out/r1_$src.ncu-rep ] && python tools/ncu_summary.py gpurun_out/r1_$src.ncu-rep > profiles/${R}_ncu_$dst.md
done
cp gpurun_out/bench_r1_lres.json profiles/${R}_bench_lres.json
[ -f gpurun_out/bench_r1_sres.json ] && cp gpurun_out/bench_r1_sres.json profiles/${R}_bench_sres.json
cp gpurun_out/microbench_r1.txt profiles/${R}_microbench.txt
python tools/summarize_launches.py gpurun_out/launches_r1_bench_lres.csv > profiles/${R}_launches_bench_lres.md
gzip -c gpurun_out/launches_r1_bench_lres.csv > profiles/${R}_launches_bench_lres.csv.gz
if [ -f gpurun_out/launches_r1_bench_sres.csv ]; then
  python tools/summarize_launches.py gpurun_out/launches_r1_bench_sres.csv > profiles/${R}_launches_bench_sres.md
fi

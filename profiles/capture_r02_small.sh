#!/bin/bash
# Small-batch part of profiles/capture_r02.sh on its own (after a change to conv3x3_sb.cu only): the same commands,
# outputs into gpurun_out/, summaries made with profiles/summarize.py.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
K='regex:conv3x3_v2_kernel|conv3x3_sb_kernel|encode_packed_nhwc|head_fc_cluster|head_conv_kernel|planes_to_nhwc|widen_outputs'
python profiles/run_small_batch.py || { echo "small-batch runner failed"; exit 1; }
$NCU --set full -k "$K" --launch-skip 87 --launch-count 4 -o gpurun_out/r02_small_n1 -f python profiles/run_small_batch.py > /dev/null 2> gpurun_out/r02_ncu5.err
$NCU --set full -k "$K" --launch-skip 259 --launch-count 4 -o gpurun_out/r02_small_n16 -f python profiles/run_small_batch.py > /dev/null 2> gpurun_out/r02_ncu6.err
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_small_launches_raw.csv python profiles/run_small_batch.py > /dev/null 2> gpurun_out/r02_ncu7.err
ls -la gpurun_out/r02_small* | awk '{print $5, $9}'

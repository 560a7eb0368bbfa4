#!/bin/bash
# Round-2 ncu captures (run on the GPU box under gpurun, from the repo root).  Every command is first run WITHOUT ncu
# and must exit 0; numbers printed under ncu are never bench values.  Outputs go to gpurun_out/ (scratch); the
# summaries committed under profiles/ are made from them with profiles/summarize.py.
set -u
mkdir -p gpurun_out
BENCH="python bench.py --no-selfplay --no-cpu-baseline --no-extras --steps 2 --warmup 3"
NCU="ncu --clock-control none"
K='regex:conv3x3_v2_kernel|conv3x3_sb_kernel|encode_packed_nhwc|head_fc_cluster|head_conv_kernel|planes_to_nhwc|widen_outputs'
$BENCH > gpurun_out/r02_prof_bench_plain.json 2> gpurun_out/r02_prof_bench_plain.err || { echo "bench failed"; exit 1; }
python profiles/run_small_batch.py || { echo "small-batch runner failed"; exit 1; }
# (1) launch list of the bench command: every kernel launch with its duration
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_bench_launches_raw.csv $BENCH > /dev/null 2> gpurun_out/r02_ncu1.err
# (the range calibration at checkpoint load launches encode_packed_nhwc once: a step of 43 matching kernels starts at 1 + 43 s)
# (2) full sets: head of a timed step (encoder, stem, first/second conv of blocks 0 and 1), its tail (first conv of the
#     last block, last conv with the fused head convolutions, head FC)
$NCU --set full -k "$K" --launch-skip 130 --launch-count 6 -o gpurun_out/r02_step_head -f $BENCH > /dev/null 2> gpurun_out/r02_ncu2.err
$NCU --set full -k "$K" --launch-skip 170 --launch-count 3 -o gpurun_out/r02_step_tail -f $BENCH > /dev/null 2> gpurun_out/r02_ncu3.err
# (3) one tower conv with source correlation (the kernel VERDICT names): second conv of block 0
$NCU --set full --import-source on -k regex:conv3x3_v2_kernel --launch-skip 125 --launch-count 1 -o gpurun_out/r02_conv2_source -f $BENCH > /dev/null 2> gpurun_out/r02_ncu4.err
# (4) small batches: n = 1 (3rd call: encoder, stem, conv1, conv2 on conv3x3_sb) and n = 16
$NCU --set full -k "$K" --launch-skip 87 --launch-count 4 -o gpurun_out/r02_small_n1 -f python profiles/run_small_batch.py > /dev/null 2> gpurun_out/r02_ncu5.err
$NCU --set full -k "$K" --launch-skip 259 --launch-count 4 -o gpurun_out/r02_small_n16 -f python profiles/run_small_batch.py > /dev/null 2> gpurun_out/r02_ncu6.err
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02_small_launches_raw.csv python profiles/run_small_batch.py > /dev/null 2> gpurun_out/r02_ncu7.err
ls -la gpurun_out/r02_* | awk '{print $5, $9}'

mkdir -p gpurun_out
CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --multi-streams 0"
$CMD > gpurun_out/plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"lk_stereo|lk_temporal|select_tracked|gftt_select|mask_and_max|pyr_fused|response_nms|bucket_scan|pad_levels" -s 60 -c 14 -o gpurun_out/prof_all_r1g -f $CMD > gpurun_out/ncu10.log 2>&1
tail -2 gpurun_out/ncu10.log

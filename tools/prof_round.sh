set -x
mkdir -p gpurun_out/r01d
O=gpurun_out/r01d
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py > $O/bench_fp32.json 2> $O/bench_fp32.err
python bench.py --dtype tf32 > $O/bench_tf32.json 2> $O/bench_tf32.err
python bench.py --dtype bf16 > $O/bench_bf16.json 2> $O/bench_bf16.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/fp32_launches_raw.csv python bench.py --steps 2 --warmup 1 > $O/ncu_fp32.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bf16_launches_raw.csv python bench.py --dtype bf16 --steps 2 --warmup 1 > $O/ncu_bf16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm|attn_core|roi_align_fwd|fc_ln|nms_lazy|topk_bitonic' -c 14 -o $O/fp32_full python tools/prof_targets.py 1 fp32 > $O/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm|attn_core|roi_align_fwd|fc_ln' -c 10 -o $O/bf16_full python tools/prof_targets.py 1 bf16 > $O/ncu_full_bf16.log 2>&1
python tools/train_step_bench.py > $O/train_step.txt 2>&1
ls -la $O

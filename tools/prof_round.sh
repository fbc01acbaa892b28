set -x
R=${1:-r02}
O=gpurun_out/$R
mkdir -p $O
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py --steps 20 --warmup 5 > $O/bench_fp32.json 2> $O/bench_fp32.err
python tools/nms_bench.py > $O/nms_bench.json 2>&1
python tools/roi_bench.py > $O/roi_bench.json 2>&1
python bench.py --dtype tf32 --no-cpu-baseline > $O/bench_tf32.json 2> $O/bench_tf32.err
python bench.py --dtype bf16 --no-cpu-baseline > $O/bench_bf16.json 2> $O/bench_bf16.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/fp32_launches_raw.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_fp32.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/bf16_launches_raw.csv python bench.py --dtype bf16 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_bf16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm|attn_core|roi_align_fwd|fc_ln|nms_|topk_bucket' -c 22 -o $O/fp32_full python tools/prof_targets.py 1 fp32 > $O/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm|attn_core|roi_align_fwd|fc_ln' -c 12 -o $O/bf16_full python tools/prof_targets.py 1 bf16 > $O/ncu_full_bf16.log 2>&1
python tools/ncu_select.py $O/fp32_full.ncu-rep > $O/fp32_ncu_full_selected.csv
python tools/ncu_select.py $O/bf16_full.ncu-rep > $O/bf16_ncu_full_selected.csv
rm -f $O/fp32_full.ncu-rep $O/bf16_full.ncu-rep      # gpurun_out/ travels back only below 64 MiB
python tools/train_step_bench.py > $O/train_step.txt 2>&1
python tools/head_train_bench.py 16 128 5 --graph > $O/head_train_step.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/head_train_launches_raw.csv python tools/head_train_bench.py 16 128 1 > $O/ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'roi_align_bwd|wgrad|attn_bwd|sk_combine' -s 60 -c 14 -o $O/train_full python tools/head_train_bench.py 16 128 1 > $O/ncu_train_full.log 2>&1
python tools/ncu_select.py $O/train_full.ncu-rep > $O/train_ncu_full_selected.csv
ls -la $O; du -sh gpurun_out

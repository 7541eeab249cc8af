#!/bin/bash
# Round artefacts on ONE GPU (run under gpurun): bench lines, ncu launch list + full capture summaries, in-graph timeline,
# hash variants, packer bench, sanitizer record.  Everything lands in gpurun_out/p/ (copied into profiles/ afterwards).
cd "$(dirname "$0")/.."
O=gpurun_out/p; mkdir -p $O
R=${ROUND:-r2}
python bench.py --steps 20 --warmup 5 > $O/${R}_bench_c3_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference_arm.json 2> $O/bench_ref.err
python bench.py --workload C3x10 --steps 10 --warmup 3 --no-pack-leg --no-cpu-baseline > $O/${R}_bench_c3x10_n1.json 2> $O/bench_x10.err
# ncu launch list of the bench command itself (durations only; cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches_c3.csv python bench.py --steps 2 --warmup 3 --no-pack-leg --no-cpu-baseline > /dev/null 2> $O/ncu_launch.err
# full captures: the kernels of the pass (second serialised pass of tools/prof_pass.py), then the kernels of an incremental epoch
ncu --set full --clock-control none --import-source on -k regex:'k_hash3|k_match2|k_decide2|k_build_tables|k_clear' -s 7 -c 7 -f -o /tmp/${R}_full python tools/prof_pass.py C3 > $O/prof_pass.log 2>&1
ncu -i /tmp/${R}_full.ncu-rep --page raw --csv > /tmp/${R}_full_raw.csv 2>> $O/prof_pass.log
python tools/ncu_summary.py /tmp/${R}_full_raw.csv $O/${R}_ncu_full_c3.json >> $O/prof_pass.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_inc_|k_decide2' -s 2 -c 8 -f -o /tmp/${R}_inc python tools/prof_inc.py C3 > $O/prof_inc.log 2>&1
ncu -i /tmp/${R}_inc.ncu-rep --page raw --csv > /tmp/${R}_inc_raw.csv 2>> $O/prof_inc.log
python tools/ncu_summary.py /tmp/${R}_inc_raw.csv $O/${R}_ncu_full_incremental.json >> $O/prof_inc.log 2>&1
EPOCHS=4 python tools/prof_inc.py C3 > $O/${R}_incremental_epoch_kernels.txt 2>&1
# in-graph timeline of one replay
KR_ENGINE_LIB=tools/libkrengine_tl.so python tools/timeline.py C3 > $O/${R}_timeline_c3.txt 2>&1
# hash kernel variants, packer bench
tools/hash_bench > $O/${R}_hash_variants.txt 2>&1
tools/pack_bench 10000 100 20 3 > $O/${R}_pack_bench_c3.json 2> $O/pack_bench.err
# sanitizer (small passes through every kernel family, incremental epochs included)
for t in memcheck racecheck synccheck initcheck; do
  echo "== compute-sanitizer --tool $t" >> $O/${R}_sanitizer.txt
  timeout 600 compute-sanitizer --tool $t python tools/sanitize_run.py 2>&1 | grep -E "ok:|ERROR SUMMARY|hash_batch|Error|error" | head -40 >> $O/${R}_sanitizer.txt
done
ls -la $O; tail -2 $O/${R}_sanitizer.txt; head -c 600 $O/${R}_bench_c3_n1.json

#!/bin/bash
# The measurement batch whose outputs are committed under profiles/ (one B200): smoke, the GPU test suite, ncu --set full of
# the four step kernels at three catalog sizes, the default bench line, the reference arm and the two launch lists.
#   gpurun --timeout 2400 -- 'bash tools/final_batch.sh r04'
tag=${1:-r04}
out=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
python -m pytest tests -m gpu -x -q > $out/${tag}_gputests.log 2>&1; tail -2 $out/${tag}_gputests.log
for n in 1000000 125000 20000; do
  ncu --set full --clock-control none --import-source on -k regex:"k_factor|k_predict_tile|k_refactor|k_update_tile" -s 8 -c 4 \
      -o $out/${tag}_final_$n -f python tools/prof_step.py $n tile 3 > $out/${tag}_ncu_$n.log 2>&1
  ncu -i $out/${tag}_final_$n.ncu-rep --page raw --csv > $out/${tag}_raw_$n.csv 2>/dev/null
done
python tools/ncu_summary.py 1000000 $out/${tag}_raw_1000000.csv 125000 $out/${tag}_raw_125000.csv 20000 $out/${tag}_raw_20000.csv > profiles/${tag}_ncu_kernels.json
cp profiles/${tag}_ncu_kernels.json $out/${tag}_ncu_kernels.json
python bench.py > $out/${tag}_bench_c4_n1.json 2> $out/${tag}_bench_c4_n1.err; tail -2 $out/${tag}_bench_c4_n1.err
python bench.py --impl reference --steps 20 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_c4.csv \
    python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > $out/${tag}_ncu_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/${tag}_launches_c3.csv \
    python bench.py --workload c3 --steps 3 --warmup 3 --no-extra --no-cpu-baseline > $out/${tag}_ncu_c3.log 2>&1
# the episodic (C3) step's kernels
ncu --set full --clock-control none --import-source on -k regex:"k_env_begin|k_factor|k_predict_tile|k_refactor|k_update_tile|ssa_det_kernel|ssa_env_reduce_kernel|k_env_refresh|k_obs_to_f32" \
    -s 54 -c 18 -o $out/${tag}_c3 -f python bench.py --workload c3 --steps 4 --warmup 3 --no-extra --no-cpu-baseline --obs-dtype float32 > $out/${tag}_ncu_c3full.log 2>&1
ncu -i $out/${tag}_c3.ncu-rep --page raw --csv > $out/${tag}_raw_c3.csv 2>/dev/null
python tools/ncu_summary.py 40960 $out/${tag}_raw_c3.csv > $out/${tag}_ncu_c3_kernels.json
echo batch done

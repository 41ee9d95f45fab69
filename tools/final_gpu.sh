#!/bin/bash
# The round's GPU evidence in one gpurun call (1 GPU):  gpurun --timeout 1500 -- 'bash tools/final_gpu.sh r02'
# Writes everything under gpurun_out/; tools/make_profiles.py and a few cp's turn it into the tracked files under profiles/.
tag=${1:-r02}
mkdir -p gpurun_out
L=ctransformers_b200/lib
if ls $L/libctransformers_*.so >/dev/null 2>&1; then
  echo "== A/B of the build variants (tools/ab.py: device-timed greedy decode, 64 steps, best of 3)"
  specs="dflt=$L/libctransformers.so"
  for f in $L/libctransformers_*.so; do n=$(basename $f .so); specs="$specs ${n#libctransformers_}=$f"; done
  timeout -k 5 420 python tools/ab.py $specs 2>&1 | grep -v "^ *$" | tee gpurun_out/${tag}_ab.txt | grep "tok/s"
  best=$(grep "tok/s" gpurun_out/${tag}_ab.txt | sed 's/: / /' | sort -k2 -n -r | head -1 | cut -d" " -f1)
  echo "best variant: $best" | tee -a gpurun_out/${tag}_ab.txt
  if [ -n "$best" ] && [ "$best" != "dflt" ]; then cp $L/libctransformers_$best.so $L/libctransformers.so; echo "(the rest of this run uses $best)" | tee -a gpurun_out/${tag}_ab.txt; fi
fi
echo "== tests"; timeout -k 5 700 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu_full.txt 2>&1; rc=$?; tail -4 gpurun_out/${tag}_pytest_gpu_full.txt | tee gpurun_out/${tag}_pytest_gpu.txt
if [ $rc -ne 0 ] && [ -f $L/libctransformers_base.so ]; then
  echo "!! tests failed on the chosen build: falling back to the base variant for the rest of the run" | tee -a gpurun_out/${tag}_ab.txt
  grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/${tag}_pytest_gpu_full.txt | head -20
  cp $L/libctransformers_base.so $L/libctransformers.so
  timeout -k 5 300 python -m pytest tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${tag}_pytest_gpu_base.txt
fi
echo "== bench (default: 192 steps)"; timeout -k 5 400 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; cut -c1-300 gpurun_out/${tag}_bench_n1.json
echo "== bench (driver's flags)"; timeout -k 5 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_n1_s20.json 2>/dev/null; cut -c1-200 gpurun_out/${tag}_bench_n1_s20.json
echo "== reference arm"; timeout -k 5 400 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${tag}_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/${tag}_bench_reference.json
echo "== prefill2048"; timeout -k 5 300 python bench.py --workload prefill2048 --steps 3 > gpurun_out/${tag}_bench_prefill2048.json 2>/dev/null; cut -c1-200 gpurun_out/${tag}_bench_prefill2048.json
echo "== trace"; timeout -k 5 200 python tools/trace_step.py gpurun_out/${tag}_trace_step.json > gpurun_out/${tag}_trace_step.txt 2>&1; tail -12 gpurun_out/${tag}_trace_step.txt
echo "== ncu launch list (per-launch time, DRAM bytes, instructions; cold caches, serialised)"
CTB_NO_SPEC=1 timeout -k 5 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
  -k regex:'k_step|k_pstep|k_sample|k_argmax|k_advance' -s 4 -c 16 --csv --log-file gpurun_out/${tag}_launches_step.csv python tools/prof_decode.py 40 8 2>&1 | tail -2
echo "== ncu --set full of one prefill launch and two decode launches"
CTB_NO_SPEC=1 timeout -k 5 500 ncu --set full --clock-control none --import-source on -k regex:'k_step|k_pstep' -s 1 -c 2 -f -o gpurun_out/${tag}_full python tools/prof_decode.py 40 3 2>&1 | tail -3
ls -la gpurun_out | tail -15

#!/usr/bin/env bash
# Where one ./vectorAdd process (the unit of the reference's 5000-process bash loop) spends its time.
# Usage: bash tools/startup_trace.sh <outdir>
set -u
OUT=${1:-gpurun_out/startup}; mkdir -p "$OUT"
cd k8s-gpu-hpa_b200
for i in 1 2 3 4 5; do
  T0=$(date +%s.%N); B200VA_TRACE_STARTUP=1 ./vectorAdd > /dev/null 2> "../$OUT/trace_$i.txt"; T1=$(date +%s.%N)
  echo "[wall] whole process $(python3 -c "print('%.1f' % (($T1-$T0)*1e3))") ms" >> "../$OUT/trace_$i.txt"
done
for i in 1 2 3; do
  T0=$(date +%s.%N); CUDA_MODULE_LOADING=EAGER B200VA_TRACE_STARTUP=1 ./vectorAdd > /dev/null 2> "../$OUT/trace_eager_$i.txt"; T1=$(date +%s.%N)
  echo "[wall] whole process $(python3 -c "print('%.1f' % (($T1-$T0)*1e3))") ms (CUDA_MODULE_LOADING=EAGER)" >> "../$OUT/trace_eager_$i.txt"
done
T0=$(date +%s.%N); bash -c "for (( c=1; c<=10; c++ )); do ./vectorAdd; done" > /dev/null 2>&1; T1=$(date +%s.%N)
python3 -c "print('{\"bash_loop_iterations\": 10, \"wall_s\": %.3f, \"s_per_process\": %.4f}' % ($T1-$T0, ($T1-$T0)/10))" > "../$OUT/bash_loop.json"
cd ..; tail -n 12 "$OUT/trace_3.txt"; cat "$OUT/trace_eager_2.txt" | tail -4; cat "$OUT/bash_loop.json"; nvidia-smi -q | grep -i "persistence" | head -2

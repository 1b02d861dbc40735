#!/usr/bin/env bash
# One batched GPU session (gpurun boxes are expensive to get): smoke, sweep, tests, bench, ncu.
# Usage (from the repo root on the GPU box):  bash tools/gpu_round.sh <tag> [stages...]
# Everything lands in gpurun_out/<tag>/ ; each stage is wrapped in `timeout`.
set -u
TAG=${1:-r01}; shift || true
STAGES=${*:-"info smoke sweep tests bench ncu"}
OUT=gpurun_out/$TAG
PKG=k8s-gpu-hpa_b200
mkdir -p "$OUT"
has() { [[ " $STAGES " == *" $1 "* ]]; }

if has info; then
  nvidia-smi > "$OUT/nvidia-smi.txt" 2>&1
  nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,memory.total --format=csv >> "$OUT/nvidia-smi.txt" 2>&1
  nproc > "$OUT/host.txt"; lscpu | head -25 >> "$OUT/host.txt"; free -g >> "$OUT/host.txt"
  { echo "cgroup cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>&1)"; echo "cfs quota: $(cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us 2>&1) / $(cat /sys/fs/cgroup/cpu/cpu.cfs_period_us 2>&1)";
    lscpu | grep -i numa; nvidia-smi topo -m; for d in /sys/bus/pci/devices/*; do [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && echo "$d numa_node=$(cat $d/numa_node)"; done; } >> "$OUT/host.txt" 2>&1
fi

if has smoke; then
  timeout 600 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; echo "smoke exit=$?" | tee -a "$OUT/status.txt"
fi

if has sweep; then
  python tools/gen_geometries.py all > "$OUT/geometries.txt"
  timeout 900 $PKG/b200va_tune --n $((1<<28)) --reps 20 --warmup 3 < "$OUT/geometries.txt" > "$OUT/tune_2p28.jsonl" 2> "$OUT/tune_2p28.err"
  echo "sweep 2^28 exit=$?" | tee -a "$OUT/status.txt"
  timeout 600 $PKG/b200va_tune --n $((1<<24)) --reps 50 --warmup 5 < "$OUT/geometries.txt" > "$OUT/tune_2p24.jsonl" 2> "$OUT/tune_2p24.err"
  echo "sweep 2^24 exit=$?" | tee -a "$OUT/status.txt"
fi

if has ab; then
  python tools/gen_ab.py ${AB_LIST:-round2} > "$OUT/ab_geometries.txt"
  for lg in ${AB_SIZES:-28 26 25 24}; do
    reps=20; [ $lg -le 26 ] && reps=80; [ $lg -le 24 ] && reps=200; [ $lg -le 20 ] && reps=1000
    timeout 900 $PKG/b200va_tune --n $((1<<lg)) --reps $reps --warmup 3 --rounds 7 < "$OUT/ab_geometries.txt" > "$OUT/ab_2p$lg.jsonl" 2> "$OUT/ab_2p$lg.err"
    echo "ab 2^$lg exit=$?" | tee -a "$OUT/status.txt"
  done
  for lg in ${AB_NOPDL_SIZES:-}; do
    reps=20; [ $lg -le 24 ] && reps=200
    B200VA_NO_PDL=1 timeout 900 $PKG/b200va_tune --n $((1<<lg)) --reps $reps --warmup 3 --rounds 7 < "$OUT/ab_geometries.txt" > "$OUT/ab_nopdl_2p$lg.jsonl" 2>> "$OUT/ab_2p$lg.err"
  done
fi

if has nsweep; then
  timeout 900 $PKG/b200va_sweep --lo 16 --hi 30 > "$OUT/sweep_n.jsonl" 2> "$OUT/sweep_n.err"; echo "nsweep exit=$?" | tee -a "$OUT/status.txt"
fi

if has e2e; then
  timeout 900 python tools/e2e_sweep.py > "$OUT/e2e_sweep.jsonl" 2> "$OUT/e2e_sweep.err"; echo "e2e sweep exit=$?" | tee -a "$OUT/status.txt"
fi

if has sanitizer; then
  for k in auto k0 k1 k2 k3; do
    timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 $PKG/vectorAdd --n 300007 --iters 2 --kernel $k > "$OUT/sanitizer_memcheck_$k.log" 2>&1
    echo "memcheck $k exit=$?" | tee -a "$OUT/status.txt"
  done
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 $PKG/vectorAdd --n 300007 --iters 2 --kernel k2 > "$OUT/sanitizer_racecheck_k2.log" 2>&1
  echo "racecheck k2 exit=$?" | tee -a "$OUT/status.txt"
  timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 $PKG/vectorAdd --n 300007 --iters 2 --kernel k2 > "$OUT/sanitizer_initcheck_k2.log" 2>&1
  echo "initcheck k2 exit=$?" | tee -a "$OUT/status.txt"
  timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 $PKG/vectorAdd --n 300007 --iters 2 --kernel k2 > "$OUT/sanitizer_synccheck_k2.log" 2>&1
  echo "synccheck k2 exit=$?" | tee -a "$OUT/status.txt"
fi

if has sanitizer2; then
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "misalignment and (auto or k2)" > "$OUT/sanitizer2_parity.log" 2>&1
  echo "memcheck pytest misalignment exit=$?" | tee -a "$OUT/status.txt"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stream.py -q -x -k "misaligned" > "$OUT/sanitizer2_stream.log" 2>&1
  echo "memcheck pytest stream exit=$?" | tee -a "$OUT/status.txt"
fi

if has tests; then
  timeout 1500 python -m pytest tests -q -m gpu -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest gpu exit=$?" | tee -a "$OUT/status.txt"
  tail -5 "$OUT/pytest_gpu.log"
fi

if has bench; then
  nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
      --format=csv -lms 200 > "$OUT/clocks.csv" &
  SMI=$!
  timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; echo "bench reference exit=$?" | tee -a "$OUT/status.txt"
  timeout 900 python bench.py > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit=$?" | tee -a "$OUT/status.txt"
  for k in k0 k1 k2 k3; do
    timeout 600 python bench.py --kernel $k --steps 100 --no-e2e --no-cpu-baseline > "$OUT/bench_$k.json" 2>> "$OUT/bench.err"
  done
  timeout 600 python bench.py --zero-copy --steps 20 --no-cpu-baseline > "$OUT/bench_zero_copy.json" 2>> "$OUT/bench.err"
  kill $SMI
  cat "$OUT/bench.json"
fi

if has cli; then
  ( cd $PKG
    timeout 300 ./vectorAdd > "../$OUT/cli_default.log" 2>&1; echo "cli default exit=$?" | tee -a "../$OUT/status.txt"
    # the reference's own shape (cuda-test-deployment.yaml:19), 20 iterations instead of 5000: wall time per process
    T0=$(date +%s.%N); bash -c "for (( c=1; c<=20; c++ )); do ./vectorAdd; done" > /dev/null 2>&1; T1=$(date +%s.%N)
    python -c "print('{\"bash_loop_iterations\": 20, \"wall_s\": %.3f, \"s_per_process\": %.4f}' % ($T1-$T0, ($T1-$T0)/20))" > "../$OUT/cli_bash_loop.json"
    timeout 600 ./vectorAdd --mode resident --n 2^28 --iters 100 --cpu-baseline --json "../$OUT/cli_2p28.json" > /dev/null 2>> "../$OUT/cli.err"
    timeout 600 ./vectorAdd --mode resident --n 2^24 --iters 5000 --nvml --duration 20 --json "../$OUT/cli_hpa_replay.json" > /dev/null 2>> "../$OUT/cli.err"
    timeout 600 ./vectorAdd --mode resident --n 2^24 --iters 5000 --graph 100 --json "../$OUT/cli_loop_graph.json" > /dev/null 2>> "../$OUT/cli.err"
    timeout 600 ./vectorAdd --mode staged --n 2^28 --iters 5 --json "../$OUT/cli_staged.json" > /dev/null 2>> "../$OUT/cli.err"
  )
fi

if has e2ex; then
  for G in ${GPUS:-1 4}; do
    i=0
    for opt in "" "--wc-inputs" "--zero-copy" "--zero-copy --wc-inputs" "--chunk-elems 16777216 --depth 2"; do
      i=$((i+1))
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29520+i)) \
          bench.py --gpus $G --steps 20 --warmup 3 --e2e-steps 6 --no-cpu-baseline $opt 2>> "$OUT/e2ex.err" | tail -1 | \
          python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'gpus': $G, 'opt': '$opt', 'e2e': d['e2e']}))" >> "$OUT/e2ex.jsonl"
    done
  done
  cat "$OUT/e2ex.jsonl"
fi

if has hpa; then
  timeout 900 python tools/hpa_trigger_replay.py 20 > "$OUT/hpa_trigger_replay.jsonl" 2> "$OUT/hpa_trigger_replay.err"; echo "hpa replay exit=$?" | tee -a "$OUT/status.txt"
fi

if has stream; then
  timeout 900 python tools/stream_bench.py > "$OUT/stream_bench.jsonl" 2> "$OUT/stream_bench.err"; echo "stream bench exit=$?" | tee -a "$OUT/status.txt"
fi

if has mgpu; then
  nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
  for G in ${GPUS:-2}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $G --steps 100 --warmup 5 > "$OUT/bench_${G}gpu.json" 2> "$OUT/bench_${G}gpu.err"; echo "bench ${G}gpu exit=$?" | tee -a "$OUT/status.txt"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --impl reference --gpus $G --steps 5 --warmup 1 > "$OUT/bench_reference_${G}gpu.json" 2>> "$OUT/bench_${G}gpu.err"
  ( cd $PKG
    timeout 600 ./vectorAdd --gpus $G --n 2^30 --iters 50 --json "../$OUT/cli_2p30_${G}gpu.json" > /dev/null 2>> "../$OUT/cli.err"; echo "cli 2^30 ${G}gpu exit=$?" | tee -a "../$OUT/status.txt"
    timeout 600 ./vectorAdd --gpus $G --n 2^28 --iters 200 --json "../$OUT/cli_2p28_${G}gpu.json" > /dev/null 2>> "../$OUT/cli.err"
    timeout 600 ./vectorAdd --gpus $G --n 1000000007 --iters 20 --json "../$OUT/cli_ragged_${G}gpu.json" > /dev/null 2>> "../$OUT/cli.err"
    timeout 600 ./vectorAdd --gpus $G --mode staged --n 2^30 --iters 3 --json "../$OUT/cli_staged_2p30_${G}gpu.json" > /dev/null 2>> "../$OUT/cli.err"
  )
  tail -1 "$OUT/bench_${G}gpu.json" | cut -c1-400
  done
fi

if has ncu; then
  # launch list of the bench command (cold-cache, serialised: compare shares, not absolutes)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
      python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1
  echo "ncu launches exit=$?" | tee -a "$OUT/status.txt"
  # full capture of the production kernel and of the variants
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:vadd_ -s 3 -c 2 -f -o "$OUT/prof_auto" \
      python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
  echo "ncu full exit=$?" | tee -a "$OUT/status.txt"
  for k in ${NCU_KERNELS:-k0 k1 k2}; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:vadd_ -s 3 -c 1 -f -o "$OUT/prof_$k" \
        $PKG/vectorAdd --mode resident --n 2^28 --iters 3 --kernel $k --verify none >> "$OUT/ncu_full.log" 2>&1
  done
fi
ls -la "$OUT" | head -50
cat "$OUT/status.txt"

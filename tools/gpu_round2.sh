#!/usr/bin/env bash
# Round-2 GPU sessions (one gpurun box per call; everything lands in gpurun_out/<tag>/).
# Usage:  bash tools/gpu_round2.sh <tag> [stages...]
#   smoke tests bench cold headline skew nsweep ncu ncu24 ncuskew pageable mgpu probe cli
set -u
TAG=${1:-r02a}; shift || true
STAGES=${*:-"info smoke tests bench"}
OUT=gpurun_out/$TAG
PKG=k8s-gpu-hpa_b200
mkdir -p "$OUT"
has() { [[ " $STAGES " == *" $1 "* ]]; }
note() { echo "$*" | tee -a "$OUT/status.txt"; }

if has info; then
  nvidia-smi > "$OUT/nvidia-smi.txt" 2>&1
  { nproc; lscpu | head -25; free -g; echo "cgroup cpu.max: $(cat /sys/fs/cgroup/cpu.max 2>&1)"; lscpu | grep -i numa; nvidia-smi topo -m;
    for d in /sys/bus/pci/devices/*; do [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && echo "$d numa_node=$(cat $d/numa_node)"; done; ulimit -l; } > "$OUT/host.txt" 2>&1
  make -C $PKG -q all; note "make -q (0 = the snapshot's binaries are up to date): $?"
fi

if has smoke; then
  timeout 600 python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; note "smoke exit=$?"
fi

if has tests; then
  timeout 1500 python -m pytest tests -q -m gpu ${PYTEST_X:-} > "$OUT/pytest_gpu.log" 2>&1; note "pytest gpu exit=$?"
  tail -5 "$OUT/pytest_gpu.log"
fi

if has bench; then
  timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"; note "bench reference exit=$?"
  timeout 900 python bench.py --steps 20 --warmup 5 > "$OUT/bench_s20.json" 2> "$OUT/bench.err"; note "bench (driver's flags) exit=$?"
  timeout 900 python bench.py > "$OUT/bench.json" 2>> "$OUT/bench.err"; note "bench exit=$?"
  timeout 600 python bench.py --chain --no-e2e --no-cpu-baseline > "$OUT/bench_chain.json" 2>> "$OUT/bench.err"; note "bench --chain exit=$?"
  for k in k0 k2 k3; do
    timeout 600 python bench.py --kernel $k --steps 100 --no-e2e --no-extras --no-cpu-baseline > "$OUT/bench_$k.json" 2>> "$OUT/bench.err"
  done
  cut -c1-600 "$OUT/bench.json"
fi

if has cold; then
  python tools/gen_ab.py cold > "$OUT/cold_geometries.txt"
  for lg in ${COLD_SIZES:-21 22 23 24}; do
    reps=400; [ $lg -ge 23 ] && reps=200
    timeout 900 $PKG/b200va_tune --n $((1<<lg)) --reps $reps --warmup 20 --rounds 7 --cold < "$OUT/cold_geometries.txt" > "$OUT/cold_2p$lg.jsonl" 2> "$OUT/cold_2p$lg.err"
    note "cold A/B 2^$lg exit=$?"
  done
fi

if has hotab; then
  # the a1 loop's shape: the same buffers relaunched (L2-assisted at these sizes), with and without early loads
  python tools/gen_ab.py cold > "$OUT/cold_geometries.txt"
  for lg in ${HOT_SIZES:-22 23 24}; do
    timeout 600 $PKG/b200va_tune --n $((1<<lg)) --reps 200 --warmup 20 --rounds 7 < "$OUT/cold_geometries.txt" > "$OUT/hot_2p$lg.jsonl" 2> "$OUT/hot_2p$lg.err"
    note "hot A/B 2^$lg exit=$?"
  done
fi

if has headline; then
  python tools/gen_ab.py headline > "$OUT/headline_geometries.txt"
  timeout 900 $PKG/b200va_tune --n $((1<<28)) --reps 20 --warmup 3 --rounds 7 --probes < "$OUT/headline_geometries.txt" > "$OUT/headline_2p28.jsonl" 2> "$OUT/headline_2p28.err"
  note "headline A/B exit=$?"
  timeout 900 $PKG/b200va_tune --n $((1<<27)) --reps 40 --warmup 3 --rounds 7 --probes < "$OUT/headline_geometries.txt" > "$OUT/headline_2p27.jsonl" 2> "$OUT/headline_2p27.err"
fi

if has skew; then
  python tools/gen_ab.py skew > "$OUT/skew_geometries.txt"
  : > "$OUT/channel_skew.jsonl"
  for d in 0 256 4096 65536 1048576 $((33554432+4096)) 1024 16384 2097152 $((1048576+256)); do
    timeout 300 $PKG/b200va_tune --n $((1<<28)) --reps 20 --warmup 3 --rounds 5 --skew $d --probes < "$OUT/skew_geometries.txt" >> "$OUT/channel_skew.jsonl" 2>> "$OUT/channel_skew.err"
  done
  note "channel skew sweep exit=$?"
fi

if has sanitizer; then
  # the round-2 kernels under compute-sanitizer: early-load form, K1c (mbarrier + try_cancel response slot), K2c after the edge fix
  cat > "$OUT/san_geos.txt" <<'GEO'
2 512 1 0 0 1 0 0 0 1 0
2 512 1 0 0 1 0 0 0 2 0
2 512 2 0 0 0 0 0 0 2 0
4 256 2 0 0 1 0 0 0 2 1
2 256 4 0 0 0 0 0 0 1 0
2 256 2 0 0 1 0 0 0 0 1
2 128 4 0 0 1 0 0 0 1 1
4 256 2 0 0 1 0 0 0 1 1
3 128 0 1 0 1 3 8192 2
GEO
  for tool in memcheck racecheck synccheck initcheck; do
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 $PKG/b200va_tune --n 300007 --reps 3 --warmup 1 --rounds 1 --cold < "$OUT/san_geos.txt" > "$OUT/sanitizer_$tool.log" 2>&1
    note "compute-sanitizer $tool exit=$? ($(grep -c 'ERROR SUMMARY: 0 errors' "$OUT/sanitizer_$tool.log") clean summaries)"
  done
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 $PKG/vectorAdd --mode resident --n 300007 --iters 40 --graph 8 > "$OUT/sanitizer_memcheck_loop.log" 2>&1
  note "compute-sanitizer memcheck launch loop exit=$?"
fi

if has nsweep; then
  timeout 900 $PKG/b200va_sweep --lo 16 --hi 30 > "$OUT/sweep_n.jsonl" 2> "$OUT/sweep_n.err"; note "nsweep exit=$?"
fi

if has ncu; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches_bench.csv" \
      python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > "$OUT/ncu_launches.log" 2>&1
  note "ncu launches exit=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:vadd_ -s 3 -c 2 -f -o "$OUT/prof_auto_2p28" \
      python bench.py --steps 5 --warmup 3 --no-e2e --no-extras --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
  note "ncu full 2^28 exit=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:vadd_ -s 3 -c 2 -f -o "$OUT/prof_chain_2p28" \
      python bench.py --chain --steps 5 --warmup 3 --no-e2e --no-extras --no-cpu-baseline >> "$OUT/ncu_full.log" 2>&1
  note "ncu full 2^28 --chain exit=$?"
fi

if has ncu24; then
  # the HPA-trigger loop (configs[4]): hot loop over the same 192 MiB -> DRAM bytes next to algorithmic bytes
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:vadd_ -s 60 -c 3 -f -o "$OUT/prof_loop_2p24" \
      $PKG/vectorAdd --mode resident --n 2^24 --iters 200 --verify none >> "$OUT/ncu_full.log" 2>&1
  note "ncu 2^24 loop exit=$?"
  timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:vadd_ -s 60 -c 3 -f -o "$OUT/prof_loop_2p24_hot" \
      $PKG/vectorAdd --mode resident --n 2^24 --iters 200 --verify none >> "$OUT/ncu_full.log" 2>&1
  note "ncu 2^24 loop (no cache flush between replays) exit=$?"
fi

if has ncucold; then
  # DRAM bytes and L2 hit rate of the three ahead-of-the-wait forms on cold 2^23 buffers (ncu serialises launches, so
  # the overlap itself is not visible here -- the point is that the prefetch adds no DRAM traffic)
  printf "2 512 2 0 0 0 0 0 0 0 0\n2 512 2 0 0 0 0 0 0 2 0\n2 512 2 0 0 0 0 0 0 1 0\n" > "$OUT/cold_trio.txt"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:vadd_vec -c 12 -f -o "$OUT/prof_cold_trio_2p23" \
      $PKG/b200va_tune --n $((1<<23)) --reps 2 --warmup 0 --rounds 1 --cold < "$OUT/cold_trio.txt" > "$OUT/ncu_cold.log" 2>&1
  note "ncu cold trio 2^23 exit=$?"
fi

if has stream; then
  timeout 900 python tools/stream_bench.py > "$OUT/stream_bench.jsonl" 2> "$OUT/stream_bench.err"; note "stream bench exit=$?"
fi

if has ncuskew; then
  M=dram__cycles_active.max.pct_of_peak_sustained_elapsed,dram__cycles_active.min.pct_of_peak_sustained_elapsed,dram__cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__read_throughput.avg.pct_of_peak_sustained_elapsed,dram__write_throughput.avg.pct_of_peak_sustained_elapsed
  echo "1 256 1 0 0 0 0 0 0" > "$OUT/one_geo.txt"; python tools/gen_ab.py skew | sed -n 2p >> "$OUT/one_geo.txt"
  for d in ${NCU_SKEWS:-0 4096 1048576 $((33554432+4096))}; do
    timeout 300 ncu --metrics $M --clock-control none -k regex:"vadd_vec|probe_" -c 12 --csv --log-file "$OUT/ncu_skew_$d.csv" \
        $PKG/b200va_tune --n $((1<<28)) --reps 1 --warmup 0 --rounds 1 --skew $d --probes < "$OUT/one_geo.txt" > /dev/null 2>> "$OUT/ncu_skew.err"
  done
  note "ncu per-channel skew exit=$?"
fi

if has pageable; then
  timeout 600 python tools/host_path_variants.py > "$OUT/host_path_variants.jsonl" 2> "$OUT/host_path_variants.err"; note "host path variants exit=$?"
  ( cd $PKG
    timeout 600 ./vectorAdd --mode staged --host-mem pageable --n 2^28 --iters 8 --json "../$OUT/cli_staged_pageable.json" > /dev/null 2>> "../$OUT/cli.err"
    timeout 600 ./vectorAdd --mode staged --n 2^28 --iters 8 --json "../$OUT/cli_staged_pinned.json" > /dev/null 2>> "../$OUT/cli.err" )
fi

if has cli; then
  ( cd $PKG
    timeout 600 ./vectorAdd --mode resident --n 2^24 --iters 5000 --graph 50 --json "../$OUT/cli_loop_graph.json" > /dev/null 2>> "../$OUT/cli.err"
    timeout 600 ./vectorAdd --mode resident --n 2^28 --iters 100 --cpu-baseline --json "../$OUT/cli_2p28.json" > /dev/null 2>> "../$OUT/cli.err" )
  timeout 600 python tools/hpa_trigger_replay.py ${HPA_S:-20} > "$OUT/hpa_trigger_replay.jsonl" 2> "$OUT/hpa_trigger_replay.err"; note "hpa replay exit=$?"
fi

if has probe; then
  timeout 600 python tools/pcie_probe_mgpu.py > "$OUT/pcie_probe_mgpu.jsonl" 2> "$OUT/pcie_probe_mgpu.err"; note "pcie probe exit=$?"
  timeout 300 python tools/pcie_probe_mgpu.py --sets "0,1,2,3;all" --placement remote > "$OUT/pcie_probe_mgpu_remote.jsonl" 2>> "$OUT/pcie_probe_mgpu.err"
fi

if has mgpu; then
  nvidia-smi topo -m > "$OUT/topo.txt" 2>&1
  for G in ${GPUS:-1 2 4 8}; do
    if [ "$G" = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29510+G))"; fi
    timeout 900 $L bench.py --gpus $G --steps 20 --warmup 5 > "$OUT/bench_${G}gpu.json" 2> "$OUT/bench_${G}gpu.err"; note "bench ${G}gpu exit=$?"
    tail -1 "$OUT/bench_${G}gpu.json" | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print(json.dumps({'n':d['n_gpus'],'value':d['value'],'e2e_ms':d['e2e']['ms_per_step'],'e2e_frac':d['e2e']['roofline']['frac'],'probe_ms':d['e2e']['roofline']['probe_ms'],'pageable_ms':d['e2e_pageable']['ms_per_step'],'first':d['e2e_pageable']['first_step_wall_ms'],'strong':d['strong_2p30']['value'],'strong_ok':d['strong_2p30']['digest_ok'],'loop_us':d['loop_2p24']['us_per_iter'],'nodes':d['e2e']['device_numaGpu_numaA_B_C_per_rank']}))" | tee -a "$OUT/mgpu_summary.jsonl"
  done
  if [ "${IDENTITY:-1}" = 1 ]; then
    for G in 2 4; do
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29530+G)) \
          bench.py --gpus $G --steps 20 --warmup 5 --identity-mapping --no-extras > "$OUT/bench_${G}gpu_identity.json" 2>> "$OUT/bench_${G}gpu.err"
      note "bench ${G}gpu identity mapping exit=$?"
    done
  fi
  ( cd $PKG
    timeout 600 ./vectorAdd --gpus 8 --n 2^30 --iters 50 --json "../$OUT/cli_2p30_8gpu.json" > /dev/null 2>> "../$OUT/cli.err"; echo "cli 2^30 8gpu exit=$?" | tee -a "../$OUT/status.txt" )
  timeout 900 python -m pytest tests/test_gpu_host_path.py -q -m gpu -k "shards_across or restores" > "$OUT/pytest_mgpu.log" 2>&1; note "pytest multi-GPU subset exit=$?"
fi
ls -la "$OUT" | head -60
cat "$OUT/status.txt"

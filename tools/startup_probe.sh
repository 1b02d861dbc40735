#!/usr/bin/env bash
# Per-process cost of the reference's launch-loop shape: a do-nothing CUDA process
# (context create + destroy) vs ./vectorAdd, 10 runs each, wall seconds per process.
set -u
OUT=${1:-gpurun_out/startup}
mkdir -p "$OUT"
cat > "$OUT/ctx_only.cu" <<'CU'
#include <cuda_runtime.h>
int main() { void* p; if (cudaMalloc(&p, 4) != cudaSuccess) return 1; cudaFree(p); cudaDeviceReset(); return 0; }
CU
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o "$OUT/ctx_only" "$OUT/ctx_only.cu" || exit 1
t() { local T0=$(date +%s.%N); for i in $(seq 10); do "$@" > /dev/null 2>&1; done; local T1=$(date +%s.%N); python -c "print('%.4f' % (($T1-$T0)/10))"; }
A=$(t "$OUT/ctx_only"); B=$(t k8s-gpu-hpa_b200/vectorAdd); C=$(t k8s-gpu-hpa_b200/vectorAdd --kernel k0)
echo "{\"context_only_s\": $A, \"vectorAdd_s\": $B, \"vectorAdd_k0_s\": $C, \"gpus_visible\": $(nvidia-smi -L | wc -l)}" | tee "$OUT/startup.json"

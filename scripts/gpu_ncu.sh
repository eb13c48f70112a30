#!/bin/bash
# round-2 ncu call (1 GPU): launch list of the bench command + full captures of the kernels this round touched
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode-sweep"
$BENCH > $O/r2_ncu_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/r2_launches.csv $BENCH > $O/r2_ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l $O/r2_launches.csv
for W in gemm_qkv_rowmajor gemm_qkv_headmajor attn_vit attn_long noise; do
  case $W in gemm*) KR=gemm_bf16;; attn_vit) KR=attn_vit_kernel;; attn_long) KR=attn_long_kernel;; noise) KR=noise_patchify;; esac
  python scripts/ncu_kernels2.py $W > $O/r2_ncu_${W}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$KR -s 2 -c 1 -f -o $O/r2_ncu_$W python scripts/ncu_kernels2.py $W > $O/r2_ncu_$W.log 2>&1
  echo "$W rc=$?"
done
ls -la $O/*.ncu-rep

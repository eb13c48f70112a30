#!/bin/bash
# round-2 final ncu call (1 GPU): launch list of the bench command on the final tree, DRAM traffic of the GEMM shapes, full
# captures of the two ViT attention kernels
O=gpurun_out
mkdir -p $O
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode-sweep"
$BENCH > $O/r2f_ncu_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/r2f_launches.csv $BENCH > $O/r2f_ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l $O/r2f_launches.csv
python scripts/gemm_traffic.py > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r2f_gemm_traffic.csv python scripts/gemm_traffic.py > /dev/null 2>&1
echo "traffic rc=$?"
for W in attn_vit attn_long; do
  python scripts/ncu_kernels2.py $W > $O/r2f_ncu_${W}_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:${W}_kernel -s 2 -c 1 -f -o $O/r2f_ncu_$W python scripts/ncu_kernels2.py $W > $O/r2f_ncu_$W.log 2>&1
  echo "$W rc=$?"
done
ls -la $O/r2f*.ncu-rep

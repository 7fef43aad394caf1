# Concurrent host-memory bandwidth of 1, 2, 4, 8 GPUs of one box (tools/microbench/host_read_probe.cu), one process per GPU.
# usage (under gpurun --gpus 8): bash tools/host_bw_probe.sh > gpurun_out/host_bw_probe.log
P=tools/microbench/host_read_probe
NG=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m
for G in 1 2 4 8; do
  [ $G -le $NG ] || continue
  echo "== $G process(es) at the same time"
  T=$(python -c "import time; print(time.time() + 6.0)")
  for i in $(seq 0 $((G-1))); do CUDA_VISIBLE_DEVICES=$i $P $T gpu$i & done
  wait
done

tag=r02
mkdir -p gpurun_out
for kv in bwd2_dst:edge_bwd2 eblk_h1:edge_fwd3 nodefwd:mlp3_fwd2 bwd:mlp3_bwd csr:segment_sum lin_p:node_gemm wgrad:wgrad_tc; do
  k=${kv%%:*}; rx=${kv##*:}
  MGN_PROF_ONLY=$k timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -f -o gpurun_out/${tag}_full_$k python tools/prof_kernels.py 1000 1000 1 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log
done
